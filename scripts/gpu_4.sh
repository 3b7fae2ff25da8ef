#!/bin/bash
# 4-GPU call: world > 2 paths of the sharded Cholesky (REST peer pushes) with the final kernel.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 scripts/bench_potrf_multi.py 8192 32768 2> gpurun_out/n4_potrf_multi.err | tee gpurun_out/n4_potrf_multi.jsonl | cut -c1-200
echo "potrf_multi rc=$?"
timeout 120 python scripts/bench_potrf_multi.py 8192 2>/dev/null | cut -c1-200
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 3 --warmup 3 --config c5s > gpurun_out/n4_bench_c5s.json 2> gpurun_out/n4_bench_c5s.err
echo "bench c5s N=4 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/n4_bench_c5s.json').read().strip().splitlines()[-1]); print(d['selfcheck'], d['ms_per_step'])"
