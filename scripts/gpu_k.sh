#!/bin/bash
# 2-GPU call: compute-sanitizer over the sharded Cholesky (CUDA IPC peer memory, in-kernel NVLink stores).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 700 compute-sanitizer --target-processes all --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_dist.py -k "potrf_sharded_world2" -x -q > gpurun_out/k_memcheck_multi.txt 2>&1
echo "memcheck multi rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Error" gpurun_out/k_memcheck_multi.txt | head -12
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -3
