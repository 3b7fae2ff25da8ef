#!/bin/bash
# GPU call B (1 GPU): new tests, cycle budget of the dataflow kernel, sanitizer logs, ncu launch list + full capture.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/b_pytest.txt; tail -3 gpurun_out/b_pytest.txt
GPAR_B200_LIB=$PWD/gpar_b200/libgpar_b200_prof.so timeout 300 python scripts/prof_budget.py 2048 4096 8424 16384 > gpurun_out/b_budget.txt 2>&1
head -c 6000 gpurun_out/b_budget.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py 700 > gpurun_out/b_memcheck.txt 2>&1; tail -4 gpurun_out/b_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_small.py 400 > gpurun_out/b_racecheck.txt 2>&1; tail -4 gpurun_out/b_racecheck.txt
timeout 400 compute-sanitizer --tool synccheck --print-limit 20 python scripts/sanitize_small.py 400 > gpurun_out/b_synccheck.txt 2>&1; tail -3 gpurun_out/b_synccheck.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/b_launches_c3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-anchor > gpurun_out/b_ncu_bench.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/b_launches_c3.csv)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:potrf_dataflow -s 2 -c 1 -o gpurun_out/r2_potrf_n8424 python scripts/one_potrf.py 8424 > gpurun_out/b_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep | tail -3
