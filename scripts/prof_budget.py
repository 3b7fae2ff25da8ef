"""Where the cycles of potrf_dataflow_kernel go: per-CTA accounting of a -DGPAR_DF_PROF build
(GPAR_B200_LIB=gpar_b200/libgpar_b200_prof.so python scripts/prof_budget.py 4096 8424 16384)."""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gpar_b200.engine import Engine  # noqa: E402
from gpar_b200.spec import lower_terms  # noqa: E402

CATS = ["ticket+decode", "mainloop", "  of which flag waits", "epilogue (count wait + RMW store + fence)",
        "wait L_jj ready", "tile_solve", "publish", "HEAD syrk", "HEAD wait PRE", "HEAD assemble", "HEAD diag factor",
        "D0", "#tasks", "#k-tiles"]


def main():
    eng = Engine()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    prof = torch.zeros(64 + 16 * 256 + 1024, dtype=torch.int64, device="cuda")
    eng.lib.gpar_debug_set_dataflow_prof(C.c_void_p(prof.data_ptr()))
    mhz = 1965.0
    out = {}
    for n in [int(a) for a in sys.argv[1:]] or [8424]:
        spec = lower_terms([dict(type="eq", variance=1.0, cols=[0, 1, 2, 3], scales=[0.25] * 4)])
        X = torch.rand(n * 4, dtype=torch.float64, device="cuda")
        d = torch.full((n,), 0.1, dtype=torch.float64, device="cuda")
        ld = n + (n & 1)
        J = eng.empty(n * ld)
        u = eng.zeros(ld)
        ms = []
        for it in range(4):
            eng.gram(spec, X, 4, n, J, ld, diag=d, lower_only=True)
            prof.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.potrf(J, ld, n, B=u, ldb=ld, nb=1)
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        tl = prof.cpu().numpy()[64 + 16 * 256:]
        pr = prof.cpu().numpy()[64:64 + 16 * 256].reshape(-1, 16)[:sms]
        t0, t1 = pr[:, 15].min(), pr[:, 14].max()
        span_us = (t1 - t0) / 1e3
        res = {"n": n, "launch_ms": ms[-1], "kernel_span_us": span_us}
        print(f"n = {n}: launch {ms[-1]:.3f} ms, kernel span {span_us:.0f} us, {sms} CTAs")
        tot = 0.0
        for c, name in enumerate(CATS):
            v = pr[:, c].astype(np.float64)
            if c >= 12:
                print(f"  {name:45s} total {v.sum():10.0f}  per CTA {v.mean():8.1f}")
                res[name] = float(v.sum())
                continue
            us = v / mhz
            res[name] = float(us.mean())
            if c != 2:
                tot += us.mean()
            print(f"  {name:45s} mean {us.mean():9.1f} us/CTA  ({100 * us.mean() / span_us:5.1f} % of span)  max {us.max():9.1f}")
        idle = span_us - tot
        res["idle/unaccounted"] = float(idle)
        print(f"  {'idle at the end / unaccounted':45s} mean {idle:9.1f} us/CTA  ({100 * idle / span_us:5.1f} % of span)")
        end = (pr[:, 14] - t0) / 1e3
        print(f"  CTA end times: min {end.min():.0f} us, median {np.median(end):.0f} us, max {end.max():.0f} us")
        ktiles = pr[:, 13].sum()
        ml = (pr[:, 1].astype(np.float64) / mhz).sum()
        print(f"  mainloop: {ml / ktiles:.2f} us per k-tile (incl. waits); {((pr[:, 1] - pr[:, 2]) / mhz).sum() / ktiles:.2f} without flag waits"
              f"  [ideal at DGEMM peak 35.4 TF: {2 * 128 ** 3 / (35.4e12 / sms) * 1e6:.2f}]")
        res["us_per_ktile"] = float(ml / ktiles)
        nt = (n + 127) // 128
        stamps = (tl[2:nt] - t0) / 1e3  # L_kk published, k = 2 .. nt - 1
        if len(stamps) > 2:
            dk = np.diff(stamps)
            print("  column pace (us between consecutive L_kk publications), k = 3 ..:")
            print("   " + " ".join(f"{v:.0f}" for v in dk))
            res["column_pace_us"] = [float(v) for v in dk]
        out[n] = res
    print(json.dumps(out))


if __name__ == "__main__":
    main()
