#!/bin/bash
# GPU call A (1 GPU): full GPU test suite, then the N = 1 bench line (C3 + baselines + C5 anchor).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
nproc > gpurun_out/a_nproc.txt; free -g >> gpurun_out/a_nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/a_pytest.txt
tail -5 gpurun_out/a_pytest.txt
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/a_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/a_bench.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "steps", "warmup"): print(k, d.get(k))
    print("e2e", d.get("e2e"))
    print("roofline", {k: d["roofline"].get(k) for k in ("achieved", "frac", "launch_ms", "cusolver_potrf_ms")})
    print("cusolver", d.get("cusolver_baseline"))
    cb = d.get("cpu_baseline", {}); print("cpu", {k: cb.get(k) for k in ("value", "cores", "t_logpdf_s", "t_condition_s", "chains_run", "t_chain_mean_s")})
    print("anchor", d.get("scale_anchor"))
except Exception as e:
    print("parse failed", e)
PY
