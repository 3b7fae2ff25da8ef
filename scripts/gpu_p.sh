#!/bin/bash
# Final evidence call (1 GPU): bench line, launch list, ncu full capture, other configs, sanitizers, cycle budget.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=$PWD/gpar_b200
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/p_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "frac_e2e", d["frac_of_fp64_roofline_end_to_end"])
print("roofline", {k: d["roofline"].get(k) for k in ("achieved", "frac", "launch_ms", "cusolver_potrf_ms")})
print("parity", {k: v for k, v in d["parity_full_size"].items() if k != "what"})
print("anchor", d["scale_anchor"].get("ms_per_step"))
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/p_launches_c3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-anchor > /dev/null 2>&1; echo "launch list rc=$? lines=$(wc -l < gpurun_out/p_launches_c3.csv)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:potrf_dataflow -s 2 -c 1 -o gpurun_out/r2_potrf_n8424_final python scripts/one_potrf.py 8424 > /dev/null 2>&1; echo "ncu full rc=$?"
timeout 300 python scripts/run_configs.py c2 c4 2>&1 | tail -2 | tee gpurun_out/p_configs.jsonl | cut -c1-330
GPAR_B200_LIB=$L/libgpar_b200_prof.so timeout 200 python scripts/prof_budget.py 4096 8424 > gpurun_out/p_budget.txt 2>&1; grep "n = \|mainloop:" gpurun_out/p_budget.txt
GPAR_B200_LIB=$L/libgpar_b200_prof.so timeout 100 python scripts/prof_chain.py 4096 > gpurun_out/p_chain.txt 2>&1; tail -3 gpurun_out/p_chain.txt | cut -c1-300
timeout 100 python scripts/prof_diag.py > gpurun_out/p_diag.txt 2>&1; tail -1 gpurun_out/p_diag.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py 700 > gpurun_out/p_memcheck.txt 2>&1; tail -1 gpurun_out/p_memcheck.txt
GPAR_B200_LIB=$L/libgpar_b200_san.so timeout 900 compute-sanitizer --tool racecheck --print-limit 30 python scripts/sanitize_small.py 400 > gpurun_out/p_racecheck_san.txt 2>&1; tail -2 gpurun_out/p_racecheck_san.txt
timeout 600 compute-sanitizer --tool racecheck --print-limit 30 python scripts/sanitize_small.py 400 > gpurun_out/p_racecheck.txt 2>&1; tail -1 gpurun_out/p_racecheck.txt
