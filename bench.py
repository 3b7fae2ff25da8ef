#!/usr/bin/env python
"""Headline benchmark: (logpdf + predict) calls per second of the GPAR hot path at
BASELINE.json configs[2] (C3: n=8192, m=4, p=8, EQ+linear, markov=2, replace+impute, 10 %
missing, n*=1024, S=100), fp64, on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps 5 --warmup 3            # our arm
    python bench.py --impl reference --steps 2 --warmup 1    # CPU reference arm (oracle port)

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (data kwargs, regressor kwargs)
    "c3": (dict(n=8192, m=4, p=8, ns=1024, S=100, missing=0.1),
           dict(scale=0.25, noise=0.1, linear=True, linear_scale=10.0, nonlinear=True, nonlinear_scale=1.0,
                markov=2, replace=True, impute=True, normalise_y=True)),
    "c2": (dict(n=4096, m=2, p=4, ns=1024, S=100, missing=0.0),
           dict(scale=0.25, noise=0.1, linear=False, nonlinear=True, nonlinear_scale=1.0, replace=False,
                impute=False, normalise_y=True)),
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


def cpu_port_step(data, reg_kw, n_cpu, ns_cpu, S_cpu):
    """One (logpdf + predict) pass of the oracle port on a bounded sample of the workload."""
    from oracle.gpar_oracle import Normals, OracleRegressor

    x, y, xs = data["x"][:n_cpu], data["y"][:n_cpu], data["xs"][:ns_cpu]
    p = y.shape[1]
    ora = OracleRegressor(**reg_kw)
    t0 = time.perf_counter()
    ora.condition(x, y)
    lp = ora.logpdf(x, y)
    queue = [data["Z"][s, i, :ns_cpu] for s in range(S_cpu) for i in range(p)]
    mean = ora.predict(xs, num_samples=S_cpu, normals=Normals(queue=queue))
    dt = time.perf_counter() - t0
    return dt, lp, mean


def cpu_baseline(data, data_kw, reg_kw, steps=1, warmup=0):
    """Times the oracle port ("port": the reference's stheno/lab stack is not installable here) on
    the host cores with all the threads MKL/OpenBLAS will use, on a bounded sample, and scales the
    rate to the full workload by the algorithmic flop ratio of SURVEY 8(d)."""
    n_cpu = min(data_kw["n"], 2048)
    ns_cpu = min(data_kw["ns"], 256)
    S_cpu = min(data_kw["S"], 2)
    for _ in range(warmup):
        cpu_port_step(data, reg_kw, n_cpu, ns_cpu, S_cpu)
    dts = [cpu_port_step(data, reg_kw, n_cpu, ns_cpu, S_cpu)[0] for _ in range(max(steps, 1))]
    dt = float(np.median(dts))
    F_sample = algorithmic_flops(data["y"][:n_cpu], ns_cpu, S_cpu, reg_kw.get("replace", False))
    F_full = algorithmic_flops(data["y"], data_kw["ns"], data_kw["S"], reg_kw.get("replace", False))
    try:
        from threadpoolctl import threadpool_info

        cores = max([d.get("num_threads", 1) for d in threadpool_info()] + [1])
    except Exception:
        cores = os.cpu_count()
    return {
        "value": (1.0 / dt) * (F_sample / F_full),
        "unit": "calls/s",
        "cores": int(cores),
        "kind": "port",
        "sample": (f"oracle port (numpy/scipy fp64, reference-faithful op order) on the first n={n_cpu} rows, "
                   f"n*={ns_cpu}, S={S_cpu} chains of the same seeded data: {dt:.3f} s per logpdf+predict; "
                   f"rate scaled to the full workload by the SURVEY 8(d) flop ratio {F_sample / F_full:.3e}"),
        "sample_seconds": dt,
    }, dt


def run_reference(args, data_kw, reg_kw):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    data = make_data(**data_kw)
    base, dt = cpu_baseline(data, data_kw, reg_kw, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "logpdf+predict calls/sec (GPAR hot path)", "value": base["value"],
        "unit": "calls/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 / base["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.config}: " + json.dumps(data_kw, sort_keys=True), **reg_kw},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "calls/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    data_kw, reg_kw = CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference(args, data_kw, reg_kw)

    import torch

    from gpar_b200 import GPARRegressor
    from gpar_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = max(args.warmup, 3)
    K = args.steps

    # The path does not shard at replace=True (all chains share their inputs):
    # N GPUs run N independent replicas (different seeded data sets), no data-path collective.
    data = make_data(seed=10 * rank, **data_kw)
    eng = Engine()
    reg = GPARRegressor(engine=eng, **reg_kw)
    S = data_kw["S"]

    def step_api():
        """The call a user makes: numpy in, numpy out (host<->device copies inside)."""
        reg.condition(data["x"], data["y"])
        lp = reg.logpdf(data["x"], data["y"])
        mean = reg.predict(data["xs"], num_samples=S)
        return lp, mean

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(W):
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

    # device-resident variant: same step with x / xs already in HBM and no result read-back
    from gpar_b200.model import DevMat
    from gpar_b200.regression import _construct_gpar

    reg.condition(data["x"], data["y"])
    xdev = DevMat.from_host(eng, reg.x, spare=reg.p + 1)
    xsdev = DevMat.from_host(eng, data["xs"], spare=reg.p + 1)
    ones_s = np.ones((data_kw["ns"], reg.p))

    def step_resident():
        gp = _construct_gpar(reg, reg.vs, reg.m, reg.p)
        lpv = gp.logpdf(xdev, reg.y, reg.w)
        gp2 = _construct_gpar(reg, reg.vs, reg.m, reg.p)
        smp = gp2.sample(xsdev, ones_s, num_samples=S, train=(xdev, reg.y, reg.w), return_device=True)
        out = eng.empty(smp.shape[1] * smp.shape[2])
        eng.mean_axis0(smp.reshape(-1), S, smp.shape[1] * smp.shape[2], out)
        return lpv, out

    for _ in range(2):
        step_resident()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step_resident()
    e1.record()
    torch.cuda.synchronize()
    t_res = e0.elapsed_time(e1) / 1e3
    sync_all()

    t_e2e, t_val = t_wall, t_res
    if dist is not None:
        tt = torch.tensor([t_wall, t_res, float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e, t_val = float(tt[0]), float(tt[1])

    if rank == 0:
        F = algorithmic_flops(data["y"], data_kw["ns"], S, reg_kw.get("replace", False))
        peaks = measure_fp64_peaks(eng)
        roof = measure_dominant_kernel(eng, peaks)
        h2d = 8 * (data["x"].size + 2 * data["y"].size + data["xs"].size)  # x, y (twice: logpdf + condition), xs
        d2h = 8 * (1 + data_kw["ns"] * data_kw["p"])
        line = {
            "metric": "logpdf+predict calls/sec (GPAR hot path)",
            "value": world * K / t_val, "unit": "calls/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * t_val / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.config}: " + json.dumps(data_kw, sort_keys=True), **reg_kw,
                       "parallelism": f"replicas x{world} (path does not shard at replace=True)",
                       "l2": "inputs_larger_than_l2 (joint Gram/Cholesky matrix 568 MB per layer)"},
            "e2e": {"value": world * K / t_e2e, "unit": "calls/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * t_e2e / K,
                    "device_ms_per_step": 1e3 * t_dev / K},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "algorithmic_flops_per_step": F,
            "achieved_tflops_end_to_end": F * K / t_e2e / 1e12,
            "frac_of_fp64_roofline_end_to_end": (F * K / t_e2e / 1e12) / peaks["dgemm_tflops"],
            "fp64_peaks": peaks,
            "roofline": roof,
            "roofline_gram": measure_gram_kernel(eng),
            "logpdf": float(lp),
        }
        if not args.no_cpu_baseline:
            line["cpu_baseline"], _ = cpu_baseline(data, data_kw, reg_kw)
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


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


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel at this shape, from the
# committed `ncu --set full` capture (profiles/r1_ncu_summary.md): 4.332 GB read + 0.289 GB written.
NCU_DRAM_BYTES_PER_LAUNCH = 4.332022e9 + 0.289156e9
NCU_DRAM_SOURCE = "ncu --set full capture of potrf_dataflow_kernel at n=8424 (profiles/r1_ncu_summary.md), per launch"


def measure_dominant_kernel(eng, peaks):
    """Roofline of the dominant kernel, potrf_dataflow_kernel (persistent tile-dataflow Cholesky,
    fp64 DMMA): one launch factors the joint [training; test] matrix of a C3 layer (n = 8424 rows,
    one appended right-hand-side row).  Algorithmic flops per launch = n^3 / 3 + n^2;
    duration = CUDA events around the launch on the launching stream (mean of 5, after warm-up)."""
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
    # HBM side of the same launch: the lower triangle is read once and written once (left-looking)
    alg_bytes = 2 * 8.0 * n * (n + 1) / 2
    return {"kernel": "potrf_dataflow_kernel (persistent tile-dataflow Cholesky, DMMA m8n8k4)", "bound": "tensor",
            "achieved": ach, "peak": peaks["dgemm_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["dgemm_tflops"],
            "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "traffic_source": NCU_DRAM_SOURCE, "launch_ms": avg * 1e3, "algorithmic_flops": flops, "algorithmic_bytes": alg_bytes,
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
    return {"kernel": "gram_kernel (fused EQ Gram + diag, lower tiles, 64 x 64 tiles)", "bound": "hbm", "achieved": ach,
            "peak": peak, "unit": "GB/s", "frac": ach / peak, "launch_ms": avg * 1e3, "algorithmic_bytes": alg_bytes,
            "peak_source": src, "shape": {"n": n, "d": d},
            "note": "one fp64 exp per entry: the fp64 pipe (not HBM) is the nearer bound, see DESIGN.md section 3"}


if __name__ == "__main__":
    main()
