#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=$PWD/gpar_b200
for v in "" _nopf; do
  GPAR_B200_LIB=$L/libgpar_b200$v.so timeout 200 python scripts/bench_potrf_variants.py 1024 2048 4096 8424 16384 2>&1 | tail -1 | tee -a gpurun_out/c_variants.jsonl
done
GPAR_B200_LIB=$L/libgpar_b200_prof.so timeout 300 python scripts/prof_budget.py 4096 8424 > gpurun_out/c_budget.txt 2>&1
grep -A3 "column pace" gpurun_out/c_budget.txt | cut -c1-1500
grep "n = \|idle\|flag waits\|ticket\|epilogue\|tile_solve\|wait L_jj" gpurun_out/c_budget.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/c_pytest.txt; tail -3 gpurun_out/c_pytest.txt
GPAR_B200_LIB=$L/libgpar_b200_san.so timeout 900 compute-sanitizer --tool racecheck --print-limit 30 python scripts/sanitize_small.py 400 > gpurun_out/c_racecheck_san.txt 2>&1; tail -3 gpurun_out/c_racecheck_san.txt
GPAR_B200_LIB=$L/libgpar_b200_san.so timeout 600 compute-sanitizer --tool synccheck --print-limit 10 python scripts/sanitize_small.py 400 > gpurun_out/c_synccheck_san.txt 2>&1; tail -2 gpurun_out/c_synccheck_san.txt
