#!/bin/bash
# End-of-round ncu evidence for the CURRENT kernels (run under gpurun after the plain runs exited 0):
# launch lists (eager forward+inverse, configs 2 and 4) and --set full of the dominant GEMM and the row-band boundary.
OUT=gpurun_out
for cfg in 2 4; do
  NFDPM_GRAPHS=0 timeout 170 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file $OUT/launches_final_cfg$cfg.csv python tools/run_cfg.py $cfg bf16 > $OUT/ncu_cfg$cfg.log 2>&1
done
# L0 GEMM2 (M=32768, N=K=512): bench_gemms launches the K=64 shape 338 times first (2 eager + 21 graph replays x 16)
timeout 170 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_tc_kernel -s 338 -c 2 \
    -o $OUT/prof_final_gemm2 -f python tools/bench_gemms.py > $OUT/ncu_full_gemm2.log 2>&1
NFDPM_GRAPHS=0 timeout 170 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:flow_boundary_tiled -c 3 -o $OUT/prof_final_rowband -f python tools/run_cfg.py 4 bf16 > $OUT/ncu_full_rowband.log 2>&1
ls -la $OUT/*.ncu-rep $OUT/launches_final_*.csv
