#!/bin/bash
# 2-GPU call: NCCL / NVLink tests (skipped on 1-GPU boxes) and the N = 2 bench path on the small C5 stand-in.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -4
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/g_pytest_dist.txt; tail -6 gpurun_out/g_pytest_dist.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --config c5s > gpurun_out/g_bench_c5s_n2.json 2> gpurun_out/g_bench_c5s_n2.err
echo "bench c5s N=2 rc=$?"; tail -c 600 gpurun_out/g_bench_c5s_n2.err; head -c 1800 gpurun_out/g_bench_c5s_n2.json; echo
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --config c5s --no-cpu-baseline --no-anchor > gpurun_out/g_bench_c5s_n1.json 2> gpurun_out/g_bench_c5s_n1.err
echo "bench c5s N=1 rc=$?"; tail -c 300 gpurun_out/g_bench_c5s_n1.err; head -c 600 gpurun_out/g_bench_c5s_n1.json; echo
