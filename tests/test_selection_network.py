"""CPU model of the list kernel's pruning network (polardecoding_b200/csrc/list_decode.cu, leaf()): the L-th and (L+1)-th
smallest of the 2L candidate path metrics from a bitonic MERGE network over index = 2*slot + {cheaper child, dearer child},
with mirror steps (partner slot = slot ^ (size-1), the partner's OTHER register), in-lane steps at distance 1 and the last
merge cut off after its mirror step.  Checked against a plain sort for every list size the kernel is built for, on random,
tie-heavy and +inf (list still filling) inputs.  This is a model of the algorithm, not of the product path."""
import numpy as np
import pytest


def network(cm, cM):
    """cm <= cM per slot (shape (L,)); returns (lo, hi) = (L-th, (L+1)-th smallest) exactly as the kernel's lanes compute them"""
    L = len(cm)
    k = np.arange(L)
    x, y = cm.copy(), cM.copy()
    size = 2
    while size <= L:
        low = (k & (size >> 1)) == 0
        ox, oy = x[k ^ (size - 1)], y[k ^ (size - 1)]
        x = np.where(low, np.minimum(x, oy), np.maximum(x, oy))
        y = np.where(low, np.minimum(y, ox), np.maximum(y, ox))
        if size < L:
            d = size >> 2
            while d > 0:
                lowd = (k & d) == 0
                px, py = x[k ^ d], y[k ^ d]
                x = np.where(lowd, np.minimum(x, px), np.maximum(x, px))
                y = np.where(lowd, np.minimum(y, py), np.maximum(y, py))
                d >>= 1
            x, y = np.minimum(x, y), np.maximum(x, y)
        size <<= 1
    lowh = (k & (L >> 1)) == 0
    v = np.where(lowh, np.maximum(x, y), np.minimum(x, y))
    d = 1
    while d < (L >> 1):
        pv = v[k ^ d]
        v = np.where(lowh, np.maximum(v, pv), np.minimum(v, pv))
        d <<= 1
    w = v[k ^ (L >> 1)]
    lo = np.where(lowh, v, w)
    hi = np.where(lowh, w, v)
    assert (lo == lo[0]).all() and (hi == hi[0]).all()      # every lane of the frame ends with the same pair
    return lo[0], hi[0]


@pytest.mark.parametrize("L", [2, 4, 8, 16, 32])
def test_network_gives_the_two_middle_order_statistics(L):
    rng = np.random.default_rng(L)
    for trial in range(400):
        kind = trial % 4
        pm = rng.random(L) * 20
        inc = rng.random(L) * 8
        if kind == 1:                       # tie-heavy: few distinct values
            pm = rng.integers(0, 3, L).astype(float)
            inc = rng.integers(0, 3, L).astype(float)
        if kind == 2:                       # list still filling: empty slots carry +inf
            pm[rng.integers(1, L + 1):] = np.inf
        if kind == 3:                       # a zero LLR: both children cost the same
            inc[rng.integers(0, L)] = 0.0
        cm, cM = pm, pm + inc
        lo, hi = network(cm, cM)
        ref = np.sort(np.concatenate([cm, cM]))
        assert lo == ref[L - 1] and hi == ref[L], (L, trial, lo, hi, ref[L - 1], ref[L])
