#!/bin/bash
# End-of-round verification (run under gpurun): GPU parity tests, the contract bench line, the per-config table.
OUT=gpurun_out
date +%s > $OUT/t0
timeout 420 python -m pytest tests -m gpu -x -q > $OUT/final_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/final_pytest.log
date +%s > $OUT/t1
timeout 240 python bench.py > $OUT/final_bench.log 2>&1; echo "bench rc=$?" >> $OUT/final_bench.log
date +%s > $OUT/t2
timeout 200 python tools/bench_configs.py --configs 1,4,5,3 --modes bf16 --steps 10 > $OUT/final_configs.log 2>&1; echo "configs rc=$?" >> $OUT/final_configs.log
date +%s > $OUT/t3
tail -3 $OUT/final_pytest.log; tail -2 $OUT/final_bench.log | cut -c1-400; tail -5 $OUT/final_configs.log | cut -c1-300
