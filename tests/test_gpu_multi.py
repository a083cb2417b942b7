"""GPU (needs >= 2 devices, skipped otherwise): partitioning the frame space over ranks must not change the result.
Counter-based Philox keyed by the global frame index makes the multi-GPU run reproduce the single-GPU counters exactly."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "polardecoding_b200", "host", "bin")


def ngpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except OSError:
        return 0


@pytest.mark.parametrize("prog,args", [("CASCL_128", ["--ebn0", "1.5:0.5:2.5", "--ble", "150"]), ("BP_128", ["--ebn0", "2.0", "--ble", "60"]),
                                       ("CASCL_1024_L8", ["--ebn0", "1.5", "--max-frames", "20000"]),
                                       ("CASCL_1024_L8", ["--ebn0", "1.5", "--ble", "4000"]),          # several pipelined rounds, exact stop
                                       ("CASCL_1024_L8", ["--ebn0", "2.0", "--max-frames", "3000001"]),  # frame budget: one all-reduce, ragged end
                                       ("BP_1024", ["--ebn0", "2.0", "--ble", "300"])])
def test_multi_gpu_equals_single_gpu(prog, args):
    n = ngpus()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    outs = []
    for g in (1, 2) + ((4,) if n >= 4 else ()) + ((8,) if n >= 8 else ()):
        r = subprocess.run([os.path.join(BIN, prog), "--seed", "11", "--gpus", str(g)] + args, stdin=subprocess.DEVNULL, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout)
    for g, o in enumerate(outs[1:]):
        assert o == outs[0], "outputs differ:\n--- 1 GPU\n%s\n--- more GPUs (#%d)\n%s" % (outs[0], g, o)
