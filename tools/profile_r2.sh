#!/bin/bash
# Round-2 profiling recipe (run under gpurun): plain run first, then the ncu launch list and --set full captures of the
# default (fp32-faithful, split bf16 pair) inference path at config 2.
OUT=gpurun_out
python tools/run_cfg.py 2 fp32 > $OUT/plain_cfg2_x3.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/launches_r2_cfg2_x3.csv python tools/run_cfg.py 2 fp32 > $OUT/ncu_l.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_nt_tc_kernel -s 1 -c 3 \
    -o $OUT/prof_r2_gemm_x3 -f python tools/run_cfg.py 2 fp32 > $OUT/ncu_f1.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm3_boundary_kernel -s 1 -c 2 \
    -o $OUT/prof_r2_gemm3_boundary_x3 -f python tools/run_cfg.py 2 fp32 > $OUT/ncu_f2.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:flow_boundary_kernel -s 3 -c 2 \
    -o $OUT/prof_r2_flow_boundary_x3 -f python tools/run_cfg.py 2 fp32 > $OUT/ncu_f3.log 2>&1
python tools/run_cfg.py 2 bf16 > $OUT/plain_cfg2_bf16.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_nt_tc_kernel -s 1 -c 1 \
    -o $OUT/prof_r2_gemm_bf16 -f python tools/run_cfg.py 2 bf16 > $OUT/ncu_f4.log 2>&1
ls -la $OUT/*.ncu-rep
