#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python scripts/bench_potrf_variants.py 1024 4096 8424 16384 2>&1 | tail -1 | tee gpurun_out/h_variants.jsonl
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/h_pytest.txt; tail -3 gpurun_out/h_pytest.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:potrf_dataflow -s 2 -c 1 -o gpurun_out/r2_potrf_n8424_b python scripts/one_potrf.py 8424 > gpurun_out/h_ncu_full.log 2>&1
echo "ncu potrf rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:trsm_rows -c 1 -o gpurun_out/r2_trsm_rows_c5 python scripts/one_trsm_rows.py 32768 65536 > gpurun_out/h_ncu_trsm.log 2>&1
echo "ncu trsm rc=$?"; tail -2 gpurun_out/h_ncu_trsm.log
timeout 300 python scripts/one_trsm_rows.py 32768 65536 | tail -1
