#!/bin/bash
# 8-GPU call: sharded Cholesky alone (with NVLink byte counters around it) and the N = 8 bench line (C5, strong-scaled).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi nvlink -gt d -i 0 > gpurun_out/n8_nvlink_before.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/bench_potrf_multi.py 32768 > gpurun_out/n8_potrf_multi.jsonl 2> gpurun_out/n8_potrf_multi.err
echo "potrf_multi rc=$?"; tail -2 gpurun_out/n8_potrf_multi.jsonl | cut -c1-300
nvidia-smi nvlink -gt d -i 0 > gpurun_out/n8_nvlink_after.txt 2>&1
python - <<'PY'
import re
def tot(path):
    tx = rx = 0
    for ln in open(path):
        m = re.search(r"Data (Tx|Rx): (\d+) KiB", ln)
        if m:
            if m.group(1) == "Tx": tx += int(m.group(2))
            else: rx += int(m.group(2))
    return tx, rx
try:
    b, a = tot("gpurun_out/n8_nvlink_before.txt"), tot("gpurun_out/n8_nvlink_after.txt")
    print("GPU0 NVLink during 4 sharded factorisations of n=32768: tx %.2f GB, rx %.2f GB" % ((a[0]-b[0])*1024/1e9, (a[1]-b[1])*1024/1e9))
except Exception as e:
    print("nvlink parse failed", e)
PY
GPAR_BENCH_BUDGET_S=${BUDGET:-60} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/n8_bench_c5.json 2> gpurun_out/n8_bench_c5.err
echo "bench rc=$?"; tail -c 500 gpurun_out/n8_bench_c5.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/n8_bench_c5.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "steps", "warmup", "n_gpus", "scaling", "selfcheck", "gpu_launches", "achieved_tflops_end_to_end", "frac_of_fp64_roofline_end_to_end"): print(k, d.get(k))
    print("e2e", d.get("e2e")); print("roofline", {k: d["roofline"].get(k) for k in ("kernel", "achieved", "frac", "launch_ms")})
except Exception as e:
    print("parse failed", e)
PY
