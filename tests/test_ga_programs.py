"""CPU: the host ports of the reference's DE-GA analysis programs (polardecoding_b200/host/polar_ga.c) print exactly what
the compiled reference prints (tests/golden/ga_*.txt = stdout of the unmodified programs; md5 as in SURVEY.md K8)."""
import hashlib
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "polardecoding_b200", "host", "bin")
MD5 = {"BPDEGA_128": "27665e52", "BPRGA_128": "9fe8b4a5", "BPRGA_1024": "a4e486f3", "BPRGA_128_allbit": "03a810c1"}


@pytest.mark.parametrize("prog", sorted(MD5))
def test_ga_program_reproduces_reference_table(prog):
    exe = os.path.join(BIN, prog)
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "polardecoding_b200", "host"), "bin/" + prog], stdout=subprocess.DEVNULL)
    out = subprocess.run([exe], stdin=subprocess.DEVNULL, capture_output=True, timeout=120).stdout
    want = open(os.path.join(ROOT, "tests", "golden", "ga_%s.txt" % prog), "rb").read()
    assert out == want
    assert hashlib.md5(out).hexdigest().startswith(MD5[prog])
