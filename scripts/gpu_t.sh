#!/bin/bash
# ncu --set full of trsm_rows_kernel with tail blocks at the C5 / 8-GPU per-rank shape
mkdir -p gpurun_out
timeout 800 ncu --set full --clock-control none --import-source on -k regex:trsm_rows -c 1 -o gpurun_out/r2_trsm_rows_c5_tail python scripts/one_trsm_rows.py 32768 65536 > gpurun_out/t_ncu_trsm.log 2>&1
echo "ncu trsm rc=$?"; tail -2 gpurun_out/t_ncu_trsm.log
