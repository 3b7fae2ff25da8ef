#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err ) 2> gpurun_out/i_bench.time
echo "bench rc=$?"; tail -3 gpurun_out/i_bench.time; tail -c 400 gpurun_out/i_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/i_bench.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "steps", "warmup", "frac_of_fp64_roofline_end_to_end"): print(k, d.get(k))
    print("e2e", d.get("e2e"))
    print("roofline", {k: d["roofline"].get(k) for k in ("achieved", "frac", "launch_ms", "cusolver_potrf_ms", "traffic")})
    print("cusolver", {k: d["cusolver_baseline"].get(k) for k in ("value", "ms_per_step", "chains_run")})
    cb = d.get("cpu_baseline", {}); print("cpu", {k: cb.get(k) for k in ("value", "cores", "t_logpdf_s", "t_condition_s", "chains_run", "t_chain_mean_s", "extrapolated")})
    print("parity", d.get("parity_full_size"))
    print("anchor", {k: d["scale_anchor"].get(k) for k in ("value", "ms_per_step", "achieved_tflops_end_to_end", "error")})
except Exception as e:
    print("parse failed", e)
PY
( time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/i_bench_ref.json 2> gpurun_out/i_bench_ref.err ) 2> gpurun_out/i_bench_ref.time
echo "reference rc=$?"; tail -3 gpurun_out/i_bench_ref.time; head -c 700 gpurun_out/i_bench_ref.json
