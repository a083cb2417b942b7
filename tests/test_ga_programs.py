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


@pytest.mark.parametrize("prog", ["BPRGA_128_W", "BPRGA_128_M", "BPRGA_1024_W"])
def test_matrix_ga_program_reproduces_reference_table(prog):
    """the three programs that read the M matrices on stdin: same stdin format as the reference, same stdout"""
    host = os.path.join(ROOT, "polardecoding_b200", "host")
    exe = os.path.join(BIN, prog)
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", host, "bin/" + prog], stdout=subprocess.DEVNULL)
    m = subprocess.run([exe, "--emit-m"], capture_output=True, timeout=120).stdout
    n, N = (10, 1024) if "1024" in prog else (7, 128)
    head = m.decode().split("\n")[:n]
    assert all(len(row.split()) == N for row in head)                        # n rows of N column weights first
    assert head[0].split()[:4] == ["2", "2", "2", "2"] and head[0].split()[N // 2] == "1"
    out = subprocess.run([exe], input=m, capture_output=True, timeout=300).stdout
    assert out == open(os.path.join(ROOT, "tests", "golden", "ga_%s.txt" % prog), "rb").read()
    bad = subprocess.run([exe], stdin=subprocess.DEVNULL, capture_output=True, timeout=60)
    assert bad.returncode == 2 and b"stdin" in bad.stderr
