#!/usr/bin/env python
"""Headline benchmark of the GPAR hot path: (logpdf + predict) calls per second, fp64.

    python bench.py --gpus 1 --steps 5 --warmup 3            # C3 on one B200 (the metric's configuration)
    torchrun ... bench.py --gpus N ...                       # N > 1: C5, strong-scaled over the N GPUs
    python bench.py --impl reference --steps 2 --warmup 1    # the reference's op sequence on the host cores

* N = 1 runs BASELINE.json configs[2] (C3: n=8192, m=4, p=8, EQ+linear, markov=2, replace+impute, 10 %
  missing, n*=1024, S=100).  With replace=True all chains share their inputs, so this workload does not shard.
* N > 1 runs BASELINE.json configs[4] (C5: n=32768, m=2, p=8, replace=False, impute=False, n*=2048, S=256) --
  the configuration north_star shards: the S chains are partitioned over the ranks (gpar_b200.dist) and every
  factorisation of the training rows is spread over the ranks by the in-kernel NVLink Cholesky
  (gpar_potrf_multi through Engine(group=...)).  Total work is fixed: "scaling": "strong".  The N = 1 line
  carries the single-GPU time of the same C5 workload as `scale_anchor` (measured in a child process).

One JSON line on stdout (rank 0).  DESIGN.md section 4 explains every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "logpdf+predict calls/sec (GPAR hot path)"

CONFIGS = {
    # name: (data kwargs, regressor kwargs)
    "c3": (dict(n=8192, m=4, p=8, ns=1024, S=100, missing=0.1),
           dict(scale=0.25, noise=0.1, linear=True, linear_scale=10.0, nonlinear=True, nonlinear_scale=1.0,
                markov=2, replace=True, impute=True, normalise_y=True)),
    "c2": (dict(n=4096, m=2, p=4, ns=1024, S=100, missing=0.0),
           dict(scale=0.25, noise=0.1, linear=False, nonlinear=True, nonlinear_scale=1.0, replace=False,
                impute=False, normalise_y=True)),
    "c5": (dict(n=32768, m=2, p=8, ns=2048, S=256, missing=0.0),
           dict(scale=0.25, noise=0.1, linear=True, linear_scale=10.0, nonlinear=True, nonlinear_scale=1.0,
                replace=False, impute=False, normalise_y=True)),
    # small stand-ins of the same shapes for tests / dry runs (never a bench line)
    "c5s": (dict(n=4608, m=2, p=3, ns=256, S=8, missing=0.0),
            dict(scale=0.25, noise=0.1, linear=True, linear_scale=10.0, nonlinear=True, nonlinear_scale=1.0,
                 replace=False, impute=False, normalise_y=True)),
    "tiny": (dict(n=512, m=2, p=3, ns=128, S=8, missing=0.1),
             dict(scale=0.25, noise=0.1, linear=True, linear_scale=10.0, nonlinear=True, nonlinear_scale=1.0,
                  markov=2, replace=True, impute=True, normalise_y=True)),
}


def make_data(n, m, p, ns, S, missing=0.0, seed=0):
    """Synthetic inputs of SURVEY.md 8(d): seeds data 0, missingness 1, test inputs 2, normals 3."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, (n, m))
    f = np.zeros((n, p))
    f[:, 0] = sum(np.sin(2 * np.pi * (k + 1) * x[:, k]) / (k + 1) for k in range(m))
    for j in range(1, p):
        f[:, j] = np.cos(f[:, j - 1]) ** 2 + np.sin(3 * x[:, j % m]) + 0.5 * f[:, j - 1]
    y = f + 0.1 * rng.standard_normal((n, p))
    if missing > 0 and p > 1:
        mask = np.random.default_rng(seed + 1).uniform(size=(n, p - 1)) < missing
        y[:, 1:][mask] = np.nan
    xs = np.random.default_rng(seed + 2).uniform(0, 1, (ns, m))
    r3 = np.random.default_rng(seed + 3)
    Z = r3.standard_normal((S, p, ns))
    Z2 = r3.standard_normal((S, p, ns))
    return dict(x=x, y=y, xs=xs, Z=Z, Z2=Z2)


def algorithmic_flops(y, ns, S, replace):
    """F_logpdf + F_predict of SURVEY.md 8(d) (LAPACK counts, minimum work, U_i per 8(d))."""
    n, p = y.shape
    avail = ~np.isnan(y)
    F = 0.0
    for i in range(p):
        na = float(avail[:, i].sum())
        F += na ** 3 / 3 + na ** 2  # logpdf
        U = 1 if (i == 0 or replace) else S
        F += na ** 3 / 3 + 2 * na ** 2 + U * (na ** 2 * ns + na * ns ** 2 + ns ** 3 / 3 + 2 * na * ns)
    F += S * p * ns ** 2
    return F


def config_dict(name, data_kw, reg_kw, world):
    """`config` of the JSON line -- identical in both arms (the driver compares them)."""
    if reg_kw.get("replace", False):
        par = f"replicas x{world} (path does not shard at replace=True)" if world > 1 else "single GPU"
    else:
        par = (f"chains sharded x{world} + Cholesky sharded x{world} (in-kernel NVLink pushes)" if world > 1
               else "single GPU")
    n = data_kw["n"] + data_kw["ns"]
    return {"workload": f"{name}: " + json.dumps(data_kw, sort_keys=True), **reg_kw, "parallelism": par,
            "l2": f"inputs_larger_than_l2 (Gram/Cholesky matrix {8.0 * n * n / 1e6:.0f} MB per layer)"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU reference arm: the reference's op sequence (oracle/torch_ref.py: torch CPU fp64 = what lab -> torch
# dispatches to; MKL potrf / trsm / gemm) on the box's host cores.
# ---------------------------------------------------------------------------------------------
def host_threads():
    """All the host threads this process may use.  torchrun exports OMP_NUM_THREADS=1; the reference arm
    overrides it (torch.set_num_threads) so that rank 0 uses the whole box like at N = 1."""
    import torch

    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(n, 1))
    return {"cores": int(torch.get_num_threads()), "os_cpu_count": os.cpu_count(), "affinity": n,
            "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")}


class _OverBudget(Exception):
    pass


def _cpu_dgemm_rate():
    """fp64 GEMM rate of the host (flop/s) from a 2048^3 torch.matmul: sizes the reference run."""
    import torch

    a = torch.rand(2048, 2048, dtype=torch.float64)
    a @ a
    t0 = time.perf_counter()
    a @ a
    return 2.0 * 2048 ** 3 / max(time.perf_counter() - t0, 1e-6)


def _cpu_entry_time():
    """Seconds per matrix entry of one elementwise fp64 pass with an `exp` (best of three on 4096^2): the Gram
    build of a reference layer costs about a dozen such passes (distances, exp, products, sums, the noise add)."""
    import torch

    a = torch.rand(4096, 4096, dtype=torch.float64)
    best = 1.0
    for _ in range(4):
        t0 = time.perf_counter()
        torch.exp(a)
        best = min(best, time.perf_counter() - t0)
    return best / a.numel()


def reference_step(data, data_kw, reg_kw, budget_s):
    """One (condition, logpdf, predict) pass of the reference op sequence at the FULL configuration:
    full-n logpdf, full-n conditioning, then the S chains one after the other (every chain at full n, n*)
    until all ran or the wall budget is used up.  Returns (seconds of the full workload, detail dict);
    when the budget cut the chains short the remaining chains are charged at the mean measured chain
    (BASELINE.md section 3 truncation rule: scaled in S only) and the result is flagged.

    Configurations whose full logpdf + conditioning cannot fit the budget or the host memory (C5: p factors of
    8.6 GB, ~10 minutes of potrf on 16 cores) run the logpdf layer by layer up to 45 % of the budget, charge
    the remaining layers and the conditioning at the measured per-layer cost and the chains from one
    measured chain-layer of a one-output model -- flagged `extrapolated` with what was measured."""
    import psutil

    from oracle.torch_ref import TorchNormals, TorchRegressor, timed_step

    S, n, p, ns = data_kw["S"], data_kw["n"], data_kw["p"], data_kw["ns"]
    rate = _cpu_dgemm_rate()
    replace = bool(reg_kw.get("replace", False))
    pred_layer = (1.4 * (n ** 3 / 3.0 + (float(n) ** 3 if replace else 0.0)) / rate
                  + 12.0 * _cpu_entry_time() * float(n) ** 2)
    pred_chain = 1.4 * p * 3.0 * float(n) ** 2 * ns / rate
    need = 1.7 * p * 8.0 * n * n  # p cached factors of the conditioned model + Gram temporaries
    avail = float(psutil.virtual_memory().available)
    detail = {"host_mem_available_gb": avail / 1e9, "host_dgemm_gflops": rate / 1e9,
              "predicted_fixed_s": 2 * p * pred_layer, "predicted_chain_s": pred_chain}
    # the full path needs the factors in host memory and its fixed part plus one chain inside the budget (the
    # chains then run until the budget is used: C3 is ~340 s on 16 cores and runs in full).  C5 (a step is ~9
    # hours of host time, its fixed part ~7 minutes) takes the bounded path below.
    if need < 0.7 * avail and 2 * p * pred_layer + pred_chain < budget_s:
        r = timed_step(reg_kw, data, S, device="cpu", budget_s=budget_s)
        t_fixed = r["t_logpdf"] + r["t_condition"]
        full = r["chains"] == S
        t_full = t_fixed + (r["t_chains"] if full else r["t_chain_mean"] * S)
        detail.update(t_logpdf_s=r["t_logpdf"], t_condition_s=r["t_condition"], chains_run=r["chains"],
                      t_chain_mean_s=r["t_chain_mean"], t_measured_s=r["t_total"], extrapolated=not full,
                      logpdf=r["logpdf"], _mean=r["mean"])
        return t_full, detail
    budget_s = min(budget_s, float(os.environ.get("GPAR_REF_BOUNDED_BUDGET_S", 300.0)))
    detail["bounded_budget_s"] = budget_s
    t_start = time.perf_counter()
    done = []

    def tick(phase, layer):
        done.append(time.perf_counter() - t_start)
        if done[-1] > 0.45 * budget_s and layer + 1 < p:
            raise _OverBudget()

    reg = TorchRegressor(device="cpu", tick=tick, **reg_kw)
    reg.condition(data["x"], data["y"])
    lp = None
    t_start = time.perf_counter()
    try:
        lp = float(reg.logpdf(data["x"], data["y"]))
    except _OverBudget:
        pass
    layers = len(done)
    t_lp = done[-1] * p / layers  # remaining layers at the mean measured layer (same n, one more input column each)
    t_cl = None
    if time.perf_counter() - t_start + 2.5 * t_lp / p < 0.9 * budget_s:
        one = TorchRegressor(device="cpu", **reg_kw)
        one.condition(data["x"], data["y"][:, :1])
        g = one.conditioned()
        t0 = time.perf_counter()
        one.sample_chain(g, data["xs"], normals=TorchNormals(queue=[data["Z"][0, 0]]))
        t_cl = time.perf_counter() - t0
    else:
        t_cl = pred_chain / p
    detail.update(t_logpdf_s=t_lp, t_condition_s=t_lp, logpdf_layers_measured=layers, chains_run=0, t_chain_layer_s=t_cl,
                  t_chain_mean_s=p * t_cl, extrapolated=True, logpdf=lp,
                  note=f"budget/memory-bounded: {layers} of {p} logpdf layers measured at full n; conditioning charged at "
                       "the logpdf cost (same op sequence), chains at p x S x one chain-layer")
    return 2 * t_lp + p * S * t_cl, detail


def cpu_baseline(data, data_kw, reg_kw, budget_s, want_mean=False):
    th = host_threads()
    t_full, detail = reference_step(data, data_kw, reg_kw, budget_s)
    S = data_kw["S"]
    what = (f"reference op sequence (oracle/torch_ref.py: torch CPU fp64 -> MKL potrf/trsm/gemm; fresh TRSM per "
            f"posterior mean/kernel call, S x p recomputation) at the FULL configuration n={data_kw['n']}, "
            f"n*={data_kw['ns']}: full logpdf {detail['t_logpdf_s']:.1f} s + full conditioning "
            f"{detail['t_condition_s']:.1f} s + {detail['chains_run']} of S={S} chains")
    if "logpdf_layers_measured" in detail:
        what = (f"reference op sequence (oracle/torch_ref.py, torch CPU fp64) at the FULL n={data_kw['n']}, n*={data_kw['ns']}, "
                f"bounded by wall budget / host memory: {detail['logpdf_layers_measured']} of {data_kw['p']} logpdf layers "
                f"measured ({detail['t_logpdf_s']:.0f} s for all {data_kw['p']} at that rate), conditioning charged at the same "
                f"cost, chains at p x S x one chain-layer ({detail['t_chain_layer_s']:.1f} s): EXTRAPOLATED")
    elif detail["extrapolated"]:
        what += " (remaining chains charged at the mean measured chain: extrapolated in S only)"
    mean = detail.pop("_mean", None)
    out = {"value": 1.0 / t_full, "unit": "calls/s", "cores": th["cores"], "kind": "port", "sample": what,
           "seconds_full_workload": t_full, "threads": th, **detail}
    return (out, mean) if want_mean else out


def run_reference(args, name, data_kw, reg_kw, world):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    data = make_data(**data_kw)
    # the driver allows 1800 s at N = 1 and 870 s per N in the scaling run
    budget = float(os.environ.get("GPAR_REF_BUDGET_S", 600.0 if world == 1 else 420.0))
    # ONE full pass of the workload (a C3 pass is minutes of CPU time; K + W passes would not fit the
    # driver's limit): steps / warmup report what was actually run, *_requested what was asked for.
    t_wall0 = time.perf_counter()
    base = cpu_baseline(data, data_kw, reg_kw, budget)
    wall = time.perf_counter() - t_wall0
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "calls/s", "n_gpus": args.gpus,
        "steps": 1, "warmup": 0, "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": 1e3 / base["value"], "measured_wall_s": wall, "extrapolated": bool(base["extrapolated"]),
        "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(name, data_kw, reg_kw, world), "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "calls/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# Our arm
# ---------------------------------------------------------------------------------------------
def _time_events(fn, reps):
    import torch

    fn()
    torch.cuda.synchronize()
    best, tot = float("inf"), 0.0
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = min(best, ms)
        tot += ms
    return best / 1e3, tot / reps / 1e3


def measure_fp64_peaks(eng):
    """MEASURED_PEAKS.json carries no fp64 figure, so the denominators are measured live:
    cuBLAS DGEMM 8192^3 (burst, best of 5) plus our raw DMMA / DFMA issue-rate probes."""
    import torch

    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=eng.device)
    b = torch.randn(n, n, dtype=torch.float64, device=eng.device)
    c = torch.empty_like(a)
    best, avg = _time_events(lambda: torch.matmul(a, b, out=c), 5)
    out = {"dgemm_tflops": 2.0 * n ** 3 / best / 1e12, "dgemm_tflops_avg": 2.0 * n ** 3 / avg / 1e12,
           "how": "torch.matmul fp64 8192^3 (cuBLAS), CUDA events, best of 5 / mean of 5"}
    del a, b, c
    for mode, name in ((0, "dmma_probe_tflops"), (1, "dfma_probe_tflops")):
        fl = [0.0]

        def run():
            fl[0] = eng.fp64_probe(mode, 4000)

        best, _ = _time_events(run, 3)
        out[name] = fl[0] / best / 1e12
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        out["hbm_gbs_measured_peaks_json"] = json.load(open(path)).get("hbm_gbs")
    return out


def _ncu_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of a kernel, from the committed
    `ncu --set full` capture summary (profiles/ncu_traffic.json; written by scripts/ncu_summary.py)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        ent = json.load(open(path)).get(kernel_key)
        if ent:
            return float(ent["dram_bytes"]), ent.get("source", "profiles/ncu_traffic.json")
    return None, None


def measure_potrf_kernel(eng, peaks):
    """Roofline of the dominant kernel of C3, potrf_dataflow_kernel (persistent tile-dataflow Cholesky, fp64
    DMMA): one launch factors the joint [training; test] matrix of a C3 layer (n = 8424 rows, one appended
    right-hand-side row).  Algorithmic flops per launch = n^3 / 3 + n^2; duration = CUDA
    events around the launch on the launching stream (mean of 5, after warm-up).  cuSOLVER's potrf
    (torch.linalg.cholesky) on the same matrix is timed beside it."""
    import torch

    from gpar_b200.spec import lower_terms

    n = 8424
    ld = n
    spec = lower_terms([dict(type="eq", variance=1.0, cols=[0, 1, 2, 3], scales=[0.25] * 4)])
    X = torch.rand(n * 4, dtype=torch.float64, device=eng.device)
    d = torch.full((n,), 0.1, dtype=torch.float64, device=eng.device)
    J = eng.empty(n * ld)
    u = eng.zeros(ld)
    times = []
    for it in range(7):
        eng.gram(spec, X, 4, n, J, ld, diag=d, lower_only=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.potrf(J, ld, n, B=u, ldb=ld, nb=1)
        b.record()
        torch.cuda.synchronize()
        if it >= 2:
            times.append(a.elapsed_time(b) / 1e3)
    avg = float(np.mean(times))
    flops = n ** 3 / 3.0 + float(n) ** 2
    ach = flops / avg / 1e12
    # cuSOLVER on the same (symmetrised) matrix
    eng.gram(spec, X, 4, n, J, ld, diag=d, lower_only=False)
    Jm = J.reshape(n, ld)
    _, t_cus = _time_events(lambda: torch.linalg.cholesky(Jm), 3)
    alg_bytes = 2 * 8.0 * n * (n + 1) / 2  # lower triangle read once, written once (left-looking)
    traffic, tsrc = _ncu_traffic("potrf_dataflow_kernel")
    return {"kernel": "potrf_dataflow_kernel (persistent tile-dataflow Cholesky, DMMA m8n8k4)", "bound": "tensor",
            "achieved": ach, "peak": peaks["dgemm_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["dgemm_tflops"],
            "traffic": traffic, "traffic_source": tsrc, "launch_ms": avg * 1e3, "algorithmic_flops": flops,
            "algorithmic_bytes": alg_bytes, "cusolver_potrf_ms": t_cus * 1e3,
            "peak_source": "cuBLAS DGEMM 8192^3 measured live in this run (MEASURED_PEAKS.json has no fp64 entry)",
            "shape": {"n": n, "appended_rows": 1}}


def measure_gram_kernel(eng):
    """Secondary roofline (north_star: "the Gram build's byte roofline"): gram_kernel building the lower
    triangle of the C3 joint matrix (n = 8424, 4 input columns, EQ + diag) -- HBM-bound on its output.
    Algorithmic bytes = 8 * n (n + 1) / 2 written + 8 * n * d read; peak = hbm_gbs of MEASURED_PEAKS.json."""
    import torch

    from gpar_b200.spec import lower_terms

    n, d = 8424, 4
    spec = lower_terms([dict(type="eq", variance=1.0, cols=list(range(d)), scales=[0.25] * d)])
    X = torch.rand(n * d, dtype=torch.float64, device=eng.device)
    dv = torch.full((n,), 0.1, dtype=torch.float64, device=eng.device)
    J = eng.empty(n * n)
    _, avg = _time_events(lambda: eng.gram(spec, X, d, n, J, n, diag=dv, lower_only=True), 5)
    alg_bytes = 8.0 * n * (n + 1) / 2 + 8.0 * n * d
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    if os.path.exists(path):
        peak, src = float(json.load(open(path)).get("hbm_gbs", peak)), "MEASURED_PEAKS.json hbm_gbs"
    ach = alg_bytes / avg / 1e9
    return {"kernel": "gram_kernel (fused EQ Gram + diag, lower tiles)", "bound": "hbm", "achieved": ach,
            "peak": peak, "unit": "GB/s", "frac": ach / peak, "launch_ms": avg * 1e3, "algorithmic_bytes": alg_bytes,
            "peak_source": src, "shape": {"n": n, "d": d},
            "note": "one fp64 exp per entry: the fp64 pipe (not HBM) is the nearer bound, see DESIGN.md section 3"}


def measure_trsm_rows_kernel(eng, peaks, n, nb):
    """Roofline of the dominant kernel of C5, trsm_rows_kernel: W = K(x*_s, X_a) L^-T for the rows of a pass of
    diverged chains (nb rows against an n x n factor): nb * n^2 flops (triangular solve, LAPACK count)."""
    import torch

    from gpar_b200.spec import lower_terms

    spec = lower_terms([dict(type="eq", variance=1.0, cols=[0, 1], scales=[0.25] * 2)])
    X = torch.rand(n * 2, dtype=torch.float64, device=eng.device)
    d = torch.full((n,), 0.1, dtype=torch.float64, device=eng.device)
    ld = n + (n & 1)
    J = eng.empty(n * ld)
    eng.gram(spec, X, 2, n, J, ld, diag=d, lower_only=True)
    ws, _ = eng.potrf(J, ld, n)
    E = torch.rand(nb * ld, dtype=torch.float64, device=eng.device)
    _, avg = _time_events(lambda: eng.trsm_rows(J, ld, n, ws, E, ld, nb), 2)
    flops = float(nb) * float(n) ** 2
    ach = flops / avg / 1e12
    traffic, tsrc = _ncu_traffic("trsm_rows_kernel")
    return {"kernel": "trsm_rows_kernel (B <- B L^-T, DMMA m8n8k4, one row block of B per CTA: whole waves of 128 rows + a tail wave of 32 / 64 / 96-row blocks)", "bound": "tensor",
            "achieved": ach, "peak": peaks["dgemm_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["dgemm_tflops"],
            "traffic": traffic, "traffic_source": tsrc, "launch_ms": avg * 1e3, "algorithmic_flops": flops,
            "algorithmic_bytes": 8.0 * (n * (n + 1) / 2 + 2.0 * nb * n),
            "peak_source": "cuBLAS DGEMM 8192^3 measured live in this run (MEASURED_PEAKS.json has no fp64 entry)",
            "shape": {"n": n, "rows": nb}}


def cusolver_baseline(data, data_kw, reg_kw):
    """Secondary bar: the reference op sequence on torch CUDA tensors -- cuSOLVER
    potrf, cuBLAS trsm / gemm, elementwise exp -- on the same B200, full configuration, all S chains."""
    import torch

    from oracle.torch_ref import timed_step

    r = timed_step(reg_kw, data, data_kw["S"], device="cuda", sync=torch.cuda.synchronize,
                   budget_s=float(os.environ.get("GPAR_CUSOLVER_BUDGET_S", 60.0)))
    full = r["chains"] == data_kw["S"]
    t_full = r["t_logpdf"] + r["t_condition"] + (r["t_chains"] if full else r["t_chain_mean"] * data_kw["S"])
    return r["mean"], {"value": 1.0 / t_full, "unit": "calls/s", "ms_per_step": 1e3 * t_full, "kind": "torch-CUDA restatement "
            "of the reference op sequence (oracle/torch_ref.py on device='cuda': torch.linalg.cholesky -> cuSOLVER, "
            "solve_triangular / matmul -> cuBLAS); library code, timed beside the product, never on its path",
            "t_logpdf_ms": 1e3 * r["t_logpdf"], "t_condition_ms": 1e3 * r["t_condition"],
            "t_chain_mean_ms": 1e3 * r["t_chain_mean"], "chains_run": r["chains"], "extrapolated_in_S": not full,
            "logpdf": r["logpdf"]}


def run_ours(args, name, data_kw, reg_kw):
    import torch

    from gpar_b200 import GPARRegressor
    from gpar_b200.dist import chain_slice, predict_sharded
    from gpar_b200.engine import Engine
    from gpar_b200.model import DevMat
    from gpar_b200.regression import _construct_gpar

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    replace = bool(reg_kw.get("replace", False))
    sharded = world > 1 and not replace  # chains + Cholesky shard; otherwise N replicas
    S = data_kw["S"]
    data = make_data(seed=0 if sharded else 10 * rank, **data_kw)
    eng = Engine(group=dist.group.WORLD if sharded else None)
    reg = GPARRegressor(engine=eng, **reg_kw)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def step_api():
        """The calls a user makes: numpy in, numpy out (host<->device copies inside)."""
        reg.condition(data["x"], data["y"])
        lp = reg.logpdf(data["x"], data["y"])
        if sharded:
            mean = predict_sharded(reg, data["xs"], num_samples=S, group=dist.group.WORLD)
        else:
            mean = reg.predict(data["xs"], num_samples=S)
        return lp, mean

    # ---- warm-up; heavy workloads (C5: tens of seconds per step) clamp K to a wall budget -----------
    W = max(args.warmup, 3)
    K = args.steps
    budget = float(os.environ.get("GPAR_BENCH_BUDGET_S", 150.0))
    step_api()  # first call: module load, allocator growth, peer-buffer exchange
    sync_all()
    t0 = time.perf_counter()
    step_api()
    sync_all()
    t_probe = time.perf_counter() - t0
    if t_probe > 5.0:
        W = 3
        K = int(min(K, max(2, budget // t_probe)))
        if dist is not None:
            kk = torch.tensor([K, W], device="cuda")
            dist.broadcast(kk, src=0)
            K, W = int(kk[0]), int(kk[1])
    for _ in range(W - 2):
        step_api()
    sync_all()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region: K steps, device time (CUDA events) and wall time ----------------
    l0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(K):
        lp, mean = step_api()
    ev1.record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t0
    t_dev = ev0.elapsed_time(ev1) / 1e3
    launches = eng.launches - l0
    sync_all()
    clocks = sampler.stop() if rank == 0 else None

    # ---- device-resident variant: same step with x / xs already in HBM and no result read-back ------
    reg.condition(data["x"], data["y"])
    xdev = DevMat.from_host(eng, reg.x, spare=reg.p + 1)
    xsdev = DevMat.from_host(eng, data["xs"], spare=reg.p + 1)
    ones_s = np.ones((data_kw["ns"], reg.p))
    c0, c1 = chain_slice(S, rank, world) if sharded else (0, S)

    def step_resident():
        reg._release_sharded()
        gp = _construct_gpar(reg, reg.vs, reg.m, reg.p)
        lpv = gp.logpdf(xdev, reg.y, reg.w)
        reg._release_sharded()
        gp2 = _construct_gpar(reg, reg.vs, reg.m, reg.p)
        out = eng.zeros(max(data_kw["ns"] * reg.p, 1))
        if c1 > c0:
            smp = gp2.sample(xsdev, ones_s, num_samples=c1 - c0, train=(xdev, reg.y, reg.w), return_device=True)
            eng.sum_axis0_add(smp.reshape(-1), c1 - c0, smp.shape[1] * smp.shape[2], out)
        if sharded:
            dist.all_reduce(out)
        return lpv, out

    # (a heavy step -- C5: tens of seconds -- runs the very same kernels as the API steps above: no extra warm-up)
    for _ in range(2 if t_probe <= 5.0 else (1 if t_probe <= 20.0 else 0)):
        step_resident()
    sync_all()
    Kr = K if t_probe <= 5.0 else 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(Kr):
        step_resident()
    e1.record()
    torch.cuda.synchronize()
    t_res = e0.elapsed_time(e1) / 1e3 * (K / Kr)
    sync_all()

    t_e2e, t_val, t_devmax = t_wall, t_res, t_dev
    if dist is not None:
        tt = torch.tensor([t_wall, t_res, t_dev], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e, t_val, t_devmax = float(tt[0]), float(tt[1]), float(tt[2])

    # ---- self-check of the sharded path: the same logpdf on ONE GPU (rank 0, unsharded engine) ------
    selfcheck = None
    if sharded:
        if rank == 0:
            solo = GPARRegressor(engine=Engine(), **reg_kw)
            solo.condition(data["x"], data["y"])
            lp1 = float(solo.logpdf(data["x"], data["y"]))
            selfcheck = {"logpdf_sharded": float(lp), "logpdf_single_gpu": lp1,
                         "rel_diff": abs(float(lp) - lp1) / max(abs(lp1), 1e-300)}
            del solo
        sync_all()

    units = (world if not sharded else 1) * K  # replicas: N independent calls per step; sharded: one call
    if rank == 0:
        F = algorithmic_flops(data["y"], data_kw["ns"], S, replace)
        peaks = measure_fp64_peaks(eng)
        h2d = 8 * (data["x"].size + 2 * data["y"].size + data["xs"].size)  # x, y (twice: logpdf + condition), xs
        d2h = 8 * (1 + data_kw["ns"] * data_kw["p"])
        line = {
            "metric": METRIC, "value": units / t_val, "unit": "calls/s", "n_gpus": world, "steps": K, "warmup": W,
            "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": 1e3 * t_val / K, "higher_is_better": True, "scaling": "strong" if sharded else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(name, data_kw, reg_kw, world),
            "e2e": {"value": units / t_e2e, "unit": "calls/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * t_e2e / K,
                    "device_ms_per_step": 1e3 * t_devmax / K},
            "gpu_launches": int(launches), "clocks": clocks,
            "algorithmic_flops_per_step": F,
            "achieved_tflops_end_to_end": F * K / t_e2e / 1e12,
            "frac_of_fp64_roofline_end_to_end": (F * K / t_e2e / 1e12) / (peaks["dgemm_tflops"] * (world if sharded else 1)),
            "fp64_peaks": peaks, "logpdf": float(lp),
        }
        if selfcheck:
            line["selfcheck"] = selfcheck
        if replace or data_kw["n"] < 16384:
            line["roofline"] = measure_potrf_kernel(eng, peaks)
            line["roofline_gram"] = measure_gram_kernel(eng)
        else:
            nb = min(37 * data_kw["ns"], (c1 - c0) * data_kw["ns"])
            line["roofline"] = measure_trsm_rows_kernel(eng, peaks, data_kw["n"], nb)
        if world == 1 and not args.no_cpu_baseline:
            del xdev, xsdev
            torch.cuda.empty_cache()
            parity = {"what": "full-size parity of this run: our logpdf / predictive means (same injected normals, "
                              "reference draw order) against the reference op sequence on torch-CUDA (all S chains) "
                              "and torch-CPU (the chains its budget allowed)"}
            try:
                mean_ref, line["cusolver_baseline"] = cusolver_baseline(data, data_kw, reg_kw)
                n_ch = line["cusolver_baseline"]["chains_run"]
                ours = reg.predict(data["xs"], num_samples=n_ch, normals={"Z": data["Z"][:n_ch]})
                parity["logpdf_rel_vs_torch_cuda"] = abs(float(lp) - line["cusolver_baseline"]["logpdf"]) / abs(float(lp))
                parity["predict_mean_rel_vs_torch_cuda"] = float(np.max(np.abs(ours - mean_ref)) / np.max(np.abs(mean_ref)))
                parity["chains_vs_torch_cuda"] = int(n_ch)
            except Exception as e:  # library path out of memory etc.: report, do not lose the line
                line["cusolver_baseline"] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
            line["cpu_baseline"], mean_cpu = cpu_baseline(data, data_kw, reg_kw,
                                                          float(os.environ.get("GPAR_CPU_BASELINE_BUDGET_S", 45.0)),
                                                          want_mean=True)
            if mean_cpu is not None and line["cpu_baseline"].get("chains_run", 0) > 0:
                n_ch = line["cpu_baseline"]["chains_run"]
                ours = reg.predict(data["xs"], num_samples=n_ch, normals={"Z": data["Z"][:n_ch]})
                parity["logpdf_rel_vs_torch_cpu"] = abs(float(lp) - line["cpu_baseline"]["logpdf"]) / abs(float(lp))
                parity["predict_mean_rel_vs_torch_cpu"] = float(np.max(np.abs(ours - mean_cpu)) / np.max(np.abs(mean_cpu)))
                parity["chains_vs_torch_cpu"] = int(n_ch)
            line["parity_full_size"] = parity
        if world == 1 and name == "c3" and not args.no_anchor and not os.environ.get("GPAR_BENCH_NO_ANCHOR"):
            line["scale_anchor"] = run_anchor()
        print(json.dumps(line))
    if dist is not None:
        sync_all()
        eng.close_peer_buffers()
        dist.destroy_process_group()


def run_anchor():
    """Single-GPU time of the workload the N > 1 runs strong-scale (C5), measured in a child process so
    that a failure there cannot take the headline line down.  1 warm-up-free timed step after one small
    warm-up pass (a C5 step is minutes on one GPU)."""
    import torch

    torch.cuda.empty_cache()
    cmd = [sys.executable, os.path.abspath(__file__), "--config", "c5", "--anchor-child", "--gpus", "1"]
    try:
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                             timeout=float(os.environ.get("GPAR_ANCHOR_TIMEOUT_S", 420.0)))
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": (out.stderr or out.stdout)[-400:]}
    except Exception as e:
        return {"error": repr(e)[:300]}


def anchor_child(name, data_kw, reg_kw):
    import torch

    from gpar_b200 import GPARRegressor
    from gpar_b200.engine import Engine

    torch.cuda.set_device(0)
    data = make_data(**data_kw)
    eng = Engine()
    reg = GPARRegressor(engine=eng, **reg_kw)
    small = make_data(**{**data_kw, "n": 2048, "ns": 256, "S": 4})
    reg.condition(small["x"], small["y"])
    reg.logpdf(small["x"], small["y"])
    reg.predict(small["xs"], num_samples=4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reg.condition(data["x"], data["y"])
    lp = reg.logpdf(data["x"], data["y"])
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    mean = reg.predict(data["xs"], num_samples=data_kw["S"])
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    F = algorithmic_flops(data["y"], data_kw["ns"], data_kw["S"], False)
    print(json.dumps({"workload": config_dict(name, data_kw, reg_kw, 1)["workload"], "n_gpus": 1, "steps": 1,
                      "warmup": "one small pass (n=2048) for allocator / module load; a C5 step is minutes",
                      "value": 1.0 / (t2 - t0), "unit": "calls/s", "ms_per_step": 1e3 * (t2 - t0),
                      "logpdf_ms": 1e3 * (t1 - t0), "predict_ms": 1e3 * (t2 - t1), "logpdf": float(lp),
                      "mean_abs_max": float(np.abs(mean).max()), "algorithmic_flops_per_step": F,
                      "achieved_tflops_end_to_end": F / (t2 - t0) / 1e12}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-anchor", action="store_true")
    ap.add_argument("--anchor-child", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    name = args.config or ("c3" if world == 1 else "c5")
    data_kw, reg_kw = CONFIGS[name]
    if args.anchor_child:
        return anchor_child(name, data_kw, reg_kw)
    if args.impl == "reference":
        return run_reference(args, name, data_kw, reg_kw, world)
    return run_ours(args, name, data_kw, reg_kw)


if __name__ == "__main__":
    main()
