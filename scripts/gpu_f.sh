#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_grad.py tests/test_gpu_sparse.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/f_pytest.txt; tail -4 gpurun_out/f_pytest.txt
echo "--- potri: tile dataflow"; timeout 300 python scripts/prof_grad.py 1024 2048 4096 7424 2>&1 | tee gpurun_out/f_prof_grad_dataflow.txt
echo "--- potri: round-1 row sweep"; GPAR_TRTRI_ROWS=1 timeout 300 python scripts/prof_grad.py 2048 4096 7424 2>&1 | tee gpurun_out/f_prof_grad_rows.txt
