#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "potrf" 2>&1 | tail -3
echo "--- variants"; timeout 120 python scripts/bench_potrf_variants.py 1024 2048 4096 8424 16384 2>&1 | tail -1 | tee gpurun_out/m_variants.jsonl
echo "--- full suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
GPAR_B200_LIB=$PWD/gpar_b200/libgpar_b200_prof.so timeout 200 python scripts/prof_budget.py 4096 8424 2>&1 | grep -E "n = |mainloop:|flag waits|idle|tile_solve|column pace" -A0 | cut -c1-200
GPAR_B200_LIB=$PWD/gpar_b200/libgpar_b200_prof.so timeout 200 python scripts/prof_budget.py 4096 2>&1 | grep -A1 "column pace" | tail -1 | cut -c1-300
