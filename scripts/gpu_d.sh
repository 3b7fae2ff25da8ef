#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=$PWD/gpar_b200
timeout 100 python scripts/prof_latency.py > gpurun_out/d_latency.txt 2>&1; cat gpurun_out/d_latency.txt
timeout 100 python scripts/prof_diag.py > gpurun_out/d_diag.txt 2>&1; cat gpurun_out/d_diag.txt
timeout 200 python scripts/bench_potrf_variants.py 1024 4096 8424 2>&1 | tail -1 | tee gpurun_out/d_variants.jsonl
GPAR_B200_LIB=$L/libgpar_b200_san.so timeout 900 compute-sanitizer --tool racecheck --print-limit 30 python scripts/sanitize_small.py 400 > gpurun_out/d_racecheck_san.txt 2>&1; tail -3 gpurun_out/d_racecheck_san.txt
GPAR_B200_LIB=$L/libgpar_b200_san.so timeout 600 compute-sanitizer --tool synccheck --print-limit 10 python scripts/sanitize_small.py 400 > gpurun_out/d_synccheck_san.txt 2>&1; tail -2 gpurun_out/d_synccheck_san.txt
timeout 600 compute-sanitizer --tool racecheck --print-limit 30 python scripts/sanitize_small.py 400 > gpurun_out/d_racecheck.txt 2>&1; tail -2 gpurun_out/d_racecheck.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 30 python scripts/sanitize_small.py 700 > gpurun_out/d_memcheck.txt 2>&1; tail -2 gpurun_out/d_memcheck.txt
