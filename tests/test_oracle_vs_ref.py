"""CPU, build container only: the oracle restatement against the reference's OWN decoders (oracle/_ref/libref_*.so,
compiled from the unmodified /root/reference sources) on fresh random frames.  Skipped where oracle/_ref is absent
(the GPU box gets the prebuilt files; a checkout without the reference skips)."""
import numpy as np
import pytest

from oracle_lib import Oracle, RefHarness, awgn_llr, have_ref, port

CASES = [("SC_128", 200, 2.0), ("SC_1024", 30, 2.0), ("SC_128_fag", 60, 1.0), ("SCL_128", 120, 1.0), ("SCL_128_fag", 60, 2.0),
         ("CASCL_128", 150, 1.5), ("SCL_1024", 12, 1.0), ("CASCL_1024_L8", 12, 1.5), ("CASCL_1024_sys", 10, 1.5),
         ("BP_128", 30, 2.0), ("BP_128_fag", 15, 1.0), ("BP_1024", 3, 2.0)]


@pytest.mark.ref
@pytest.mark.parametrize("prog,B,ebn0", CASES)
def test_decoder_matches_compiled_reference(prog, B, ebn0):
    if not have_ref(prog):
        pytest.skip("oracle/_ref not built (no reference tree)")
    o, h = Oracle(prog), RefHarness(prog)
    assert (o.I == h.I).all() and (o.inI == h.inI).all() and o.L == h.L and o.iters == h.iters
    u, llr = o.frames_ref_stream(ebn0, B, seed=4242)   # reference-style frames incl. CRC, reference noise generator
    assert (h.decode(llr) == o.decode(llr)[0]).all()


@pytest.mark.ref
def test_chk_phi_rng_match_compiled_reference():
    if not have_ref("CASCL_1024_L8"):
        pytest.skip("oracle/_ref not built")
    import ctypes as C
    h = RefHarness("CASCL_1024_L8")
    lib = port()
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.standard_normal(4000) * 3, [0.0, -0.0, 0.196, 0.433, 0.71, 1.05, 1.508, 2.252, 4.5, 999.0, -999.0]])
    b = np.concatenate([rng.standard_normal(4000) * 3, [0.0, 0.0, 0.196, -0.433, 0.0, 1.05, -1.508, 2.252, -4.5, 0.3, 7.0]])
    for x, y in zip(a, b):
        assert h.lib.ref_chk(x, y) == lib.po_chk(x, y)
        assert h.lib.ref_phi(x, 0) == lib.po_phi(x, 0) and h.lib.ref_phi(x, 1) == lib.po_phi(x, 1)
    from oracle_lib import PoRng
    g = PoRng()
    lib.po_rng_seed(C.byref(g), 1242)
    h.lib.ref_rng_restart(1242)
    p, q, r, s = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    for _ in range(2000):
        lib.po_rng_normal_pair(C.byref(g), 0.8, C.byref(p), C.byref(q))
        h.lib.ref_normal_pair(0.8, C.byref(r), C.byref(s))
        assert p.value == r.value and q.value == s.value
