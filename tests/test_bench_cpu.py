"""bench.py's reference arm on CPU (tiny configuration): the JSON contract of the line, the full-workload path and
the budget-bounded path (what C5 takes on a real box)."""
import json
import os
import subprocess
import sys

import numpy as np

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "tiny",
                          "--steps", "20", "--warmup", "5"], stdout=subprocess.PIPE, text=True, check=True,
                         env={**os.environ, "GPAR_REF_BUDGET_S": "60"}).stdout
    line = json.loads(out.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "calls/s" and line["dtype"] == "f64"
    assert line["steps"] == 1 and line["steps_requested"] == 20 and line["warmup_requested"] == 5
    assert line["extrapolated"] is False and line["cpu_baseline"]["chains_run"] == 8
    # the run really took what it claims: one pass of ms_per_step inside the measured wall time
    assert line["ms_per_step"] / 1e3 <= line["measured_wall_s"] + 1e-6
    assert line["e2e"] == {"value": line["value"], "unit": "calls/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the config of the two arms is one function of (name, data, regressor kwargs, world)
    data_kw, reg_kw = bench.CONFIGS["tiny"]
    assert line["config"] == bench.config_dict("tiny", data_kw, reg_kw, 1)
    assert line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["kind"] == "port"


def test_reference_step_budget_bounded_path(monkeypatch):
    """A host too slow for the budget (forced through the calibration) takes the layer-by-layer path and says so."""
    data_kw, reg_kw = bench.CONFIGS["tiny"]
    data = bench.make_data(**data_kw)
    monkeypatch.setattr(bench, "_cpu_dgemm_rate", lambda: 1.0e3)  # 1 kflop/s: nothing fits
    t_full, detail = bench.reference_step(data, data_kw, reg_kw, budget_s=5.0)
    assert detail["extrapolated"] is True and 1 <= detail["logpdf_layers_measured"] <= data_kw["p"]
    assert t_full > 0 and detail["t_chain_layer_s"] > 0
    # against the full path on the same data: the extrapolation is within a small factor for this tiny case
    monkeypatch.undo()
    t_ref, d_ref = bench.reference_step(data, data_kw, reg_kw, budget_s=60.0)
    assert d_ref["extrapolated"] is False and np.isfinite(d_ref["logpdf"])
    assert 0.1 < t_full / t_ref < 10.0


def test_algorithmic_flops_counts_distinct_input_sets():
    """U_i = 1 with replace (chains share their inputs), S otherwise."""
    y = np.zeros((100, 3))
    a = bench.algorithmic_flops(y, 10, 7, replace=True)
    b = bench.algorithmic_flops(y, 10, 7, replace=False)
    per_chain = 100.0 ** 2 * 10 + 100 * 10 ** 2 + 10 ** 3 / 3 + 2 * 100 * 10
    assert np.isclose(b - a, 2 * 6 * per_chain)
