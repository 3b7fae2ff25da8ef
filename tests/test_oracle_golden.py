"""Pins the oracle's integer/boolean path bit-exactly against the golden
vectors held by the reference's own tests (reference tests/test_model.py:30-38,
46-52, 55-105; tests/test_regression.py:43-83)."""
import numpy as np
import pytest

from oracle.gpar_oracle import determine_indices, last, merge, per_output, vector_from_init


def test_merge_golden():
    original = np.array([1, 2, 3, 4])
    updates = np.array([5, 6])
    assert merge(original, updates, np.array([True, True, False, False])).tolist() == [5, 6, 3, 4]
    assert merge(original, updates, np.array([True, False, True, False])).tolist() == [5, 2, 6, 4]


def test_last_golden():
    xs = [1, 2, 3, 4]
    assert list(last(xs)) == [(False, 1), (False, 2), (False, 3), (True, 4)]
    assert list(last(xs, [1, 2])) == [(False, 2), (False, 3)]
    assert list(last(xs, [0, 3])) == [(False, 1), (True, 4)]
    assert list(last([])) == []
    assert list(last([], [0, 1])) == []


Y_GOLDEN = np.array(
    [
        [1, 2, np.nan, np.nan],
        [3, np.nan, 4, np.nan],
        [5, 6, 7, np.nan],
        [8, np.nan, np.nan, np.nan],
        [9, 10, np.nan, np.nan],
        [11, np.nan, np.nan, 12],
    ]
)
EXPECTED_KEEP_FALSE = [
    ([1, 3, 5, 8, 9, 11], [True, True, True, True, True, True]),
    ([2, 6, 10], [True, False, True, False, True, False]),
    ([7], [False, True, False]),
    ([], [False]),
]
EXPECTED_KEEP_TRUE = [
    ([1, 3, 5, 8, 9, 11], [True, True, True, True, True, True]),
    ([2, None, 6, 10, None], [True, True, True, False, True, True]),
    ([4, 7, None], [False, True, True, False, True]),
    ([12], [False, False, True]),
]


@pytest.mark.parametrize("which", [0, 1])
def test_per_output_golden(which):
    def run(keep):
        out = []
        for yi, wi, mask in per_output(Y_GOLDEN, Y_GOLDEN, keep=keep):
            assert yi.ndim == 2 and wi.ndim == 1
            v = yi[:, 0] if which == 0 else wi
            out.append(([None if np.isnan(c) else c for c in v.tolist()], mask.tolist()))
        return out

    assert run(False) == EXPECTED_KEEP_FALSE
    assert run(True) == EXPECTED_KEEP_TRUE


def test_per_output_caching():
    assert list(per_output({True: [2, 3], False: [3, 4]}, None, keep=True)) == [2, 3]
    assert list(per_output({True: [2, 3], False: [4]}, None, keep=False)) == [4]


def test_vector_from_init():
    assert vector_from_init(2, 2).tolist() == [2, 2]
    assert vector_from_init(np.array([1, 2, 3]), 2).tolist() == [1, 2]
    with pytest.raises(ValueError):
        vector_from_init(np.random.randn(2, 2), 1)
    with pytest.raises(ValueError):
        vector_from_init(np.array([1, 2]), 3)


DETERMINE_GOLDEN = [
    ((1, 0, None), ([0], [], 0)), ((1, 1, None), ([0], [1], 1)), ((1, 2, None), ([0], [1, 2], 2)),
    ((2, 0, None), ([0, 1], [], 0)), ((2, 1, None), ([0, 1], [2], 1)), ((2, 2, None), ([0, 1], [2, 3], 2)),
    ((1, 0, 0), ([0], [], 0)), ((1, 1, 0), ([0], [], 0)), ((1, 2, 0), ([0], [], 0)),
    ((2, 0, 0), ([0, 1], [], 0)), ((2, 1, 0), ([0, 1], [], 0)), ((2, 2, 0), ([0, 1], [], 0)),
    ((1, 0, 1), ([0], [], 0)), ((1, 1, 1), ([0], [1], 1)), ((1, 2, 1), ([0], [2], 1)),
    ((2, 0, 1), ([0, 1], [], 0)), ((2, 1, 1), ([0, 1], [2], 1)), ((2, 2, 1), ([0, 1], [3], 1)),
    ((1, 0, 2), ([0], [], 0)), ((1, 1, 2), ([0], [1], 1)), ((1, 2, 2), ([0], [1, 2], 2)),
    ((2, 0, 2), ([0, 1], [], 0)), ((2, 1, 2), ([0, 1], [2], 1)), ((2, 2, 2), ([0, 1], [2, 3], 2)),
]


@pytest.mark.parametrize("args,expected", DETERMINE_GOLDEN)
def test_determine_indices_golden(args, expected):
    assert determine_indices(*args) == expected


# ---- committed oracle outputs on small seeded workloads (scripts/make_golden.py) -------------------
def test_oracle_matches_committed_fixture():
    import os
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "scripts"))
    import make_golden

    gold = np.load(os.path.join(root, "tests", "golden", "oracle_small.npz"))
    for name in make_golden.CASES:
        got = make_golden.run(name)
        for k, v in got.items():
            np.testing.assert_allclose(v, gold[k], rtol=1e-9, atol=1e-10, err_msg=k)
