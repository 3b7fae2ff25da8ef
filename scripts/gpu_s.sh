#!/bin/bash
# trsm_rows tail blocks: correctness, then old vs new library at the C5 (8 GPU) and other shapes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "trsm" 2>&1 | tail -4
for shape in "32768 65536" "8192 20480" "4096 4096" "4096 102400"; do
  for lib in old new; do
    if [ $lib = old ]; then export GPAR_B200_LIB=$PWD/gpar_b200/libgpar_b200_old.so; else unset GPAR_B200_LIB; fi
    echo -n "$lib: "; timeout 300 python scripts/one_trsm_rows.py $shape 2>&1 | tail -1
  done
done | tee gpurun_out/s_trsm_rows.txt
unset GPAR_B200_LIB
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
