"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck): a few frames each, results checked against the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from polardecoding_b200 import Engine
from oracle_lib import Oracle

for prog, B, real in (("CASCL_1024_L8", 9, "f32"), ("CASCL_1024_L8", 5, "f64"), ("SC_1024", 33, "f32"), ("CASCL_128", 13, "f64"), ("SCL_128", 7, "f32"),
                      ("BP_128", 5, "f64"), ("BP_1024", 2, "f32"), ("CASCL_1024_sys", 5, "f64")):
    eng = Engine(prog, real=real, seed=3, bp_early_stop=1 if prog.startswith("BP") else 0, iter_max=12 if prog.startswith("BP") else 0)
    llr, u = eng.channel(2.0, 5, B)
    got, flags = eng.decode_llr(llr)
    acc, fe = eng.simulate_batch(2.0, 5, B, want_frame_err=True)
    if real == "f64":
        want, _ = Oracle(prog).decode(llr, iters=12 if prog.startswith("BP") else None)
        assert (got == want).all(), prog
    print(prog, real, "ok", acc.as_dict())
    eng.close()
