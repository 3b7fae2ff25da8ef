"""CPU tests of the product's host logic (bit-exact integer path against the reference's
golden vectors) and of the C-ABI shared library: it must load and export every symbol that
include/gpar_b200.h declares.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest

from gpar_b200 import _lib
from gpar_b200.model import last, merge, per_output
from gpar_b200.spec import LayerModel, Vars, determine_indices, lower_terms, model_terms, vector_from_init
from tests.test_oracle_golden import DETERMINE_GOLDEN, EXPECTED_KEEP_FALSE, EXPECTED_KEEP_TRUE, Y_GOLDEN
from oracle import gpar_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_merge_last_golden():
    original = np.array([1, 2, 3, 4])
    updates = np.array([5, 6])
    assert merge(original, updates, np.array([True, True, False, False])).tolist() == [5, 6, 3, 4]
    assert merge(original, updates, np.array([True, False, True, False])).tolist() == [5, 2, 6, 4]
    xs = [1, 2, 3, 4]
    assert list(last(xs)) == [(False, 1), (False, 2), (False, 3), (True, 4)]
    assert list(last(xs, [1, 2])) == [(False, 2), (False, 3)]
    assert list(last(xs, [0, 3])) == [(False, 1), (True, 4)]
    assert list(last([])) == []
    assert list(last([], [0, 1])) == []


@pytest.mark.parametrize("which", [0, 1])
def test_per_output_golden(which):
    def run(keep):
        out = []
        for yi, wi, mask in per_output(Y_GOLDEN, Y_GOLDEN, keep=keep):
            v = yi[:, 0] if which == 0 else wi
            out.append(([None if np.isnan(c) else c for c in v.tolist()], mask.tolist()))
        return out

    assert run(False) == EXPECTED_KEEP_FALSE
    assert run(True) == EXPECTED_KEEP_TRUE
    assert list(per_output({True: [2, 3], False: [4]}, None, keep=False)) == [4]


def test_per_output_matches_oracle_random():
    rng = np.random.default_rng(0)
    for _ in range(20):
        n, p = rng.integers(1, 40), rng.integers(1, 6)
        y = rng.standard_normal((n, p))
        y[rng.uniform(size=(n, p)) < 0.35] = np.nan
        w = rng.uniform(size=(n, p))
        for keep in (False, True):
            a = list(per_output(y, w, keep=keep))
            b = list(O.per_output(y, w, keep=keep))
            assert len(a) == len(b)
            for (ya, wa, ma), (yb, wb, mb) in zip(a, b):
                assert np.array_equal(ma, mb)
                assert np.array_equal(ya, yb, equal_nan=True) and np.array_equal(wa, wb)


@pytest.mark.parametrize("args,expected", DETERMINE_GOLDEN)
def test_determine_indices_golden(args, expected):
    assert determine_indices(*args) == expected


def test_vector_from_init():
    assert vector_from_init(2, 2).tolist() == [2, 2]
    assert vector_from_init(np.array([1, 2, 3]), 2).tolist() == [1, 2]
    with pytest.raises(ValueError):
        vector_from_init(np.random.randn(2, 2), 1)
    with pytest.raises(ValueError):
        vector_from_init(np.array([1, 2]), 3)


def test_model_terms_match_oracle_and_names():
    cfg = dict(scale=0.5, scale_tie=False, per=True, per_period=2.0, per_scale=1.5, per_decay=10.0, input_linear=True,
               input_linear_scale=3.0, linear=True, linear_scale=7.0, nonlinear=True, nonlinear_scale=0.3, rq=True,
               markov=2, noise=[0.1, 0.2, 0.3, 0.4])
    vs, vo = Vars(), O._Vars()
    for pi in range(4):
        ta, na = model_terms(vs, 2, pi, **cfg)
        tb, nb = O.model_terms(vo, 2, pi, **cfg)
        assert na == nb and len(ta) == len(tb)
        for a, b in zip(ta, tb):
            assert a["type"] == b["type"] and list(a.get("cols", [])) == list(b.get("cols", []))
    assert sorted(vs.names) == sorted(vo.values.keys())
    assert "3/output/nonlin/alpha" in vs and "2/input/per/pers" in vs and "1/input/lin/const" in vs
    vt = Vars()
    model_terms(vt, 2, 0, **{**cfg, "scale_tie": True})
    model_terms(vt, 2, 1, **{**cfg, "scale_tie": True})
    assert "0/input/scales" in vt and "1/input/scales" not in vt


def test_lower_terms_feature_form():
    terms = [dict(type="eq", variance=2.0, cols=[0, 1], scales=[0.5, 4.0]),
             dict(type="periodic", variance=1.5, cols=[0, 1], scales=[1, 2, 3, 4], periods=[2.0, 4.0], decays=[10, 20]),
             dict(type="linear", variance=1.0, cols=[], scales=[]),
             dict(type="const", variance=0.7),
             dict(type="rq", variance=1.0, cols=[3], scales=[2.0], alpha=0.01)]
    s = lower_terms(terms)
    assert s.n_terms == 4 and s.n_feats == 2 + 6 + 0 + 1
    assert [s.terms[i].type for i in range(4)] == [_lib.TERM_EQ, _lib.TERM_EQ, _lib.TERM_CONST, _lib.TERM_RQ]
    assert (s.terms[1].f_begin, s.terms[1].f_end) == (2, 8)
    assert [s.feat_op[i] for i in range(2, 8)] == [1, 1, 2, 2, 0, 0]
    assert s.feat_a[0] == 2.0 and s.feat_a[1] == 0.25 and s.feat_a[4] == 1 / 3 and s.feat_a[7] == 1 / 20
    assert abs(s.feat_b[2] - np.pi) < 1e-15 and s.feat_col[8] == 3
    f, noise = LayerModel(terms, 0.3)
    assert noise == 0.3 and f.spec.n_terms == 4


def test_vars_latent_roundtrip():
    vs = Vars()
    vs.bnd("a", np.array([0.5, 2.0]))
    vs.bnd("b", 1e-2, lower=1e-3, upper=1e3)
    vs.get("c", 1.0)
    names = ["a", "b", "c"]
    z = vs.get_latent_vector(names)
    vs2 = vs.copy()
    vs2.set_latent_vector(names, z)
    for nme in names:
        np.testing.assert_allclose(vs2[nme], vs[nme], rtol=1e-12)


def test_shared_library_exports_header_symbols():
    header = open(os.path.join(ROOT, "include", "gpar_b200.h")).read()
    declared = set(re.findall(r"\b(gpar_[a-z0-9_]+)\s*\(", header))
    declared -= {"gpar_kernel_spec_t", "gpar_term_t"}
    assert len(declared) >= 18
    assert os.path.exists(_lib.LIB_PATH), "run `python -m gpar_b200.build` (build() does this)"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert not any("debug" in n or "probe" in n for n in declared)  # diagnostics live in gpar_b200_debug.h
    dheader = open(os.path.join(ROOT, "include", "gpar_b200_debug.h")).read()
    ddecl = set(re.findall(r"\b(gpar_[a-z0-9_]+)\s*\(", dheader))
    assert ddecl == set(_lib.DEBUG_HOOKS) | set(_lib.DEBUG_PROBES), ddecl
    for name in _lib.DEBUG_HOOKS:
        assert hasattr(lib, name)
    dlib = ctypes.CDLL(_lib.DEBUG_LIB_PATH)
    for name in _lib.DEBUG_PROBES:
        assert hasattr(dlib, name) and not hasattr(lib, name)
    assert _lib.load().gpar_abi_version() == 1
    # struct layout must match the header (terms are 32 B, spec = 8 + 8*32 + 96*(4+4+8+8))
    assert ctypes.sizeof(_lib.Term) == 32
    assert ctypes.sizeof(_lib.KernelSpec) == 8 + 8 * 32 + 96 * 24


def test_engine_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from gpar_b200.engine import Engine

    with pytest.raises(_lib.GparError):
        Engine()


# ---- task list of the Cholesky dataflow kernel and its multi-GPU partition (host decode, no GPU) ----
@pytest.mark.parametrize("n,nb,batch,grid", [(100, 0, 1, 148), (128, 1, 1, 148), (300, 1, 1, 148), (1000, 130, 1, 148),
                                             (2049, 1, 1, 148), (640, 0, 3, 148), (5000, 1, 1, 148), (8424, 1, 1, 148)])
def test_dataflow_task_list_covers_every_tile_once_and_orders_dependencies(n, nb, batch, grid):
    import ctypes as C

    from gpar_b200 import _lib
    from gpar_b200.dist import tile_row_owner

    lib = _lib.load()
    nt, nbt = (n + 127) // 128, (nb + 127) // 128 if nb > 0 else 0
    total = lib.gpar_debug_decode_ticket(n, nb, batch, grid, -1, None)
    out = (C.c_int32 * 6)()
    D0, HEAD, PLAIN, PRE = 0, 1, 2, 3
    tasks = []
    for t in range(total):
        lib.gpar_debug_decode_ticket(n, nb, batch, grid, t, out)
        tasks.append(tuple(out))
    final = {}  # (matrix, what, i, j) -> ticket of the task that finishes it
    split_seen = False
    for t, (kind, b, i, j, part, nparts) in enumerate(tasks):
        assert 0 <= part < nparts <= 8
        split_seen |= nparts > 1
        if part > 0:  # K-parts of a tile carry consecutive tickets, in order
            assert tasks[t - 1] == (kind, b, i, j, part - 1, nparts)
        if part + 1 < nparts:
            continue
        if kind == D0:
            keys = [(b, "diag", 0, 0)]
        elif kind == HEAD:
            assert j == i - 1
            keys = [(b, "tile", i, j), (b, "diag", i, i)]
        elif kind == PLAIN:
            assert i > j
            keys = [(b, "tile", i, j)]
        else:
            assert kind == PRE and i == j and i >= 2
            keys = [(b, "pre", i, i)]
        for k in keys:
            assert k not in final, f"{k} scheduled twice"
            final[k] = t
    if n >= 5000 and batch == 1:
        assert split_seen  # the tail of a large sweep is split
    for b in range(batch):
        # coverage: every diagonal tile, every tile below the diagonal (incl. appended rows), PRE for k >= 2
        for k in range(nt):
            assert (b, "diag", k, k) in final
            if k >= 2:
                assert (b, "pre", k, k) in final
        for j in range(nt):
            for i in range(j + 1, nt + nbt):
                assert (b, "tile", i, j) in final
    # dependencies of every task (partial or final): the k-tiles it streams and, for final parts, diag j
    for t, (kind, b, i, j, part, nparts) in enumerate(tasks):
        if kind == D0:
            continue
        nk = i - 1 if kind == PRE else j
        k0, k1 = part * nk // nparts, (part + 1) * nk // nparts
        if nparts > 1:
            assert k1 - k0 >= 3  # no crumbs
        deps = [(b, "tile", i, l) for l in range(k0, k1)]
        if kind != PRE:
            deps += [(b, "tile", j, l) for l in range(k0, k1)]
            if part + 1 == nparts:
                deps.append((b, "diag", j, j))
        head = kind == HEAD
        for d in deps:
            if head and d == (b, "tile", i, i - 2):
                assert final[d] <= t + 2 * nparts  # the tile ticketed right behind the head (and its parts)
            else:
                assert final[d] < t, f"{tasks[t]} depends on later ticket {d}"
        if head and i >= 2 and part + 1 == nparts:
            assert final[(b, "pre", i, i)] <= t + 3 * nparts
    # multi-GPU partition: blocks of GPAR_ROW_BLOCK = 4 tile rows, dealt cyclically
    for world in (2, 3, 8):
        owners = [tile_row_owner(i, world) for i in range(nt + nbt)]
        assert owners[:4] == [0, 0, 0, 0][: len(owners[:4])]
        assert set(owners) == set(range(min(world, (nt + nbt + 3) // 4)))


def test_numpy_virtual_index_reproduces_np_percentile():
    """Host half of the device credible bounds: (j, gamma) + numpy's _lerp == np.percentile bit for bit."""
    from gpar_b200.engine import Engine

    rng = np.random.default_rng(0)
    for ns in (1, 2, 3, 10, 100, 101, 257):
        a = np.sort(rng.standard_normal(ns))
        for q in (2.5, 97.5, 50.0, 0.0, 100.0):
            j, g = Engine.numpy_virtual_index(ns, q)
            lo, hi = a[j], a[min(j + 1, ns - 1)]
            d = hi - lo
            got = hi - d * (1.0 - g) if g >= 0.5 else lo + d * g
            assert got == np.percentile(a, q), (ns, q)


def test_regressor_reports_kernel_spec_limits_at_condition():
    """Wide fully-connected models exceed the fixed-size device kernel spec (GPAR_MAX_FEATS): the regressor says
    so when the data arrives, and markov=k lifts it."""
    from gpar_b200 import GPARRegressor

    x = np.zeros((4, 2)); y = np.zeros((4, 60))
    reg = GPARRegressor(linear=True, nonlinear=True)
    with pytest.raises(ValueError, match="markov"):
        reg.condition(x, y)
    GPARRegressor(linear=True, nonlinear=True, markov=3).condition(x, y)


# ---- row-block plan of gpar_trsm_rows (host function shared with the launch; no GPU) ----
def test_trsm_row_plan_covers_rows_in_whole_waves_plus_one_tail_wave():
    import ctypes as C

    from gpar_b200 import _lib

    lib = _lib.load()
    out = (C.c_int64 * 3)()
    rng = np.random.default_rng(0)
    cases = [1, 4, 31, 32, 33, 100, 128, 129, 4096, 148 * 32, 148 * 32 + 1, 148 * 128 - 1, 148 * 128, 148 * 128 + 1,
             65536, 20480, 102400, 27904, 31761] + [int(v) for v in rng.integers(1, 400000, 200)]
    for sms in (148, 132, 7):
        for nb in cases:
            assert lib.gpar_debug_trsm_row_plan(nb, sms, out) == 0
            full, tail, h = int(out[0]), int(out[1]), int(out[2])
            assert full % sms == 0                                 # whole waves of 128-row blocks
            rest = nb - full * 128
            if tail == 0:                                          # they cover everything (last one may be ragged)
                assert -128 < rest <= 0
                continue
            assert rest > 0
            assert h in (32, 64, 96, 128) and 0 < tail <= sms      # one more wave
            assert (tail - 1) * h < rest <= tail * h               # covers the rest exactly, last block ragged
            if h > 32:                                             # the smallest height that fits one wave
                assert -(-rest // (h - 32)) > sms
    # the C5 per-rank pass on 8 GPUs: 512 row tiles -> 3 waves + 136 blocks of 64 rows
    lib.gpar_debug_trsm_row_plan(65536, 148, out)
    assert tuple(out) == (444, 136, 64)
    assert lib.gpar_debug_trsm_row_plan(0, 148, out) != 0


def test_wave_efficiency_model_of_the_chain_passes():
    from gpar_b200.engine import wave_efficiency

    assert wave_efficiency(148, 148) == 1.0 and wave_efficiency(296, 148) == 1.0
    assert wave_efficiency(512, 148) == pytest.approx(512 / (3.6 * 148))      # 3 waves + 64-row tail blocks
    assert wave_efficiency(149, 148) == pytest.approx(149 / (1.4 * 148))      # one block left: 32-row blocks
    assert wave_efficiency(148 + 120, 148) == pytest.approx(268 / (2.0 * 148))  # too many rows left: a full wave
    assert all(0.0 < wave_efficiency(t, 148) <= 1.0 + 1e-12 for t in range(1, 2000))
