"""CPU model of the list kernel's frozen-prefix routine (polardecoding_b200/csrc/list_decode.cu, "the frozen prefix").
Before the first information bit every decision is 0, so the SC schedule is a fixed butterfly.  The kernel evaluates the
subtree over the first 2^D leaves level by level, in place, copies out the stage blocks that hold the first information
bit, and resumes its serial loop there.  This model checks, for every prefix length and several block lengths, that
  (1) the leaf LLRs of the butterfly equal those of the serial array-form schedule (the reference's order of operations), and
  (2) the copied-out blocks are exactly what the serial schedule holds in its stage arrays when it reaches that leaf group.
It is a model of the algorithm (numpy, fp64, same CHK formula), not of the product path."""
import numpy as np
import pytest

T = [(0.196, 0.65), (0.433, 0.55), (0.71, 0.45), (1.05, 0.35), (1.508, 0.25), (2.252, 0.15), (4.5, 0.05)]


def tbl(x):
    for t, v in T:
        if x < t:
            return v
    return 0.0


def chk(a, b):
    m = min(abs(a), abs(b))
    s = (1.0 if a >= 0 else -1.0) * (1.0 if b >= 0 else -1.0)
    return s * m + (tbl(abs(a + b)) - tbl(abs(a - b)))


def serial(llr, n, stop_leaf):
    """array-form SC schedule with all decisions 0 up to (not including) leaf `stop_leaf`: returns the leaf LLRs seen and the
    stage arrays (stage s: 2^s values) as they stand when the schedule is about to process leaf stop_leaf"""
    N = 1 << n
    st = {s: np.zeros(1 << s) for s in range(n)}
    lam = []
    for j in range(stop_leaf):
        t = n - 1 if j == 0 else (j & -j).bit_length() - 1
        for s in range(t, -1, -1):
            src = llr if s + 1 == n else st[s + 1]
            h = 1 << s
            if s == t and j != 0:
                st[s] = np.array([src[i + h] + src[i] for i in range(h)])          # g with partial sum 0
            else:
                st[s] = np.array([chk(src[i], src[i + h]) for i in range(h)])       # f
        lam.append(st[0][0])
    return np.array(lam), st


def butterfly(llr, n, P):
    """the kernel's routine for P leaf groups: returns leaf LLRs of the subtree and the copied-out blocks {stage: values}"""
    D = (4 * P - 1).bit_length()
    inside = 4 * P < (1 << D)
    buf = llr.copy()
    for s in range(n - 1, D - 1, -1):                       # f-layers down to stage D
        h = 1 << s
        buf = np.array([chk(buf[i], buf[i + h]) for i in range(h)])
    out = {}
    for s in range(D - 1, -1, -1):                          # in-place levels: blocks of 2^(s+1) -> f half | g half
        h = 1 << s
        nb = np.empty_like(buf)
        for b0 in range(0, 1 << D, 2 * h):
            for i in range(h):
                up, lo = buf[b0 + i], buf[b0 + h + i]
                nb[b0 + i] = chk(up, lo)
                nb[b0 + h + i] = lo + up
        buf = nb
        if inside and s >= 3:
            blk = ((4 * P) >> s) << s
            out[s] = buf[blk:blk + h].copy()
    return buf, out, D, inside


@pytest.mark.parametrize("n", [5, 6, 7, 8])
def test_prefix_butterfly_equals_the_serial_schedule(n):
    N = 1 << n
    rng = np.random.default_rng(n)
    llr = rng.standard_normal(N) * 3 + 1
    for P in range(2, N // 8 + 1):
        leaves, out, D, inside = butterfly(llr, n, P)
        lam, st = serial(llr, n, 4 * P)
        assert np.array_equal(leaves[:4 * P], lam), (n, P)
        for s, blk in out.items():
            if (4 * P) % (1 << s):                          # otherwise the resuming chain rewrites that stage first
                assert np.array_equal(blk, st[s]), (n, P, s)
        if inside:                                          # stages the resuming chain reads must all have been provided
            c = ((P & -P).bit_length() - 1) + 3             # it opens with g at stage ctz(P)+2, which reads stage ctz(P)+3
            assert c >= D or c in out or c < 3, (n, P, c, D)
