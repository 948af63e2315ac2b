"""CPU: the Munkres restatement returns an OPTIMAL assignment (cost equal to scipy's
linear_sum_assignment) on random, tie-heavy and rectangular matrices, for both zero-search
start rules.  Which optimal pairing is chosen on ties is the restated munkres 1.1.x order
(parity unpinned, see oracle/munkres_ref.py)."""
import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment

from oracle.munkres_ref import Munkres


def _cost(m, pairs):
    return sum(m[r][c] for r, c in pairs)


@pytest.mark.parametrize("rule", ["previous", "origin"])
def test_optimal_cost_random(rule):
    rng = np.random.default_rng(0)
    for trial in range(120):
        r = int(rng.integers(1, 13))
        c = int(rng.integers(r, 14))
        if trial % 3 == 0:
            m = rng.integers(0, 4, size=(r, c)).astype(np.float64)      # tie heavy
        elif trial % 3 == 1:
            m = np.round(rng.random((r, c)) * 3) * 100 - rng.random((r, c))  # group.py:66 shape
        else:
            m = rng.standard_normal((r, c))
        pairs = Munkres(rule).compute(m.copy())
        assert len(pairs) == r
        assert len({p[0] for p in pairs}) == r and len({p[1] for p in pairs}) == r
        rr, cc = linear_sum_assignment(m)
        assert np.isclose(_cost(m, pairs), m[rr, cc].sum(), rtol=1e-12, atol=1e-9)


def test_padded_columns_like_group_py():
    """num_added > num_grouped: group.py:71-78 appends 1e10 columns."""
    rng = np.random.default_rng(1)
    m = np.round(rng.random((7, 3)) * 2) * 100 - rng.random((7, 3))
    sq = np.concatenate((m, np.zeros((7, 4)) + 1e10), axis=1)
    pairs = Munkres().compute(sq.copy())
    real = [(r, c) for r, c in pairs if c < 3]
    assert len(pairs) == 7 and len(real) == 3
    rr, cc = linear_sum_assignment(sq)
    assert np.isclose(_cost(sq, pairs), sq[rr, cc].sum())


def test_rows_are_ascending_and_input_shape_respected():
    m = np.array([[4.0, 1.0, 3.0], [2.0, 0.0, 5.0]])
    pairs = Munkres().compute(m.copy())
    assert [p[0] for p in pairs] == [0, 1]
    assert all(0 <= c < 3 for _, c in pairs)
