"""compute-sanitizer over a tiny pass through every kernel family (the reference has no
race detection; its Chamfer kernels launch on the legacy stream and scatter with atomics).  memcheck catches
out-of-bounds / misaligned accesses (e.g. an unguarded NN index), racecheck shared-memory hazards between the
warp roles of the tcgen05 / TMA pipelines."""
import os
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sanitizer():
    for c in (shutil.which("compute-sanitizer"), "/usr/local/cuda/bin/compute-sanitizer"):
        if c and os.path.exists(c):
            return c
    return None


@pytest.mark.parametrize("tool", ["memcheck", "racecheck"])
def test_compute_sanitizer_is_clean(tool):
    exe = _sanitizer()
    if exe is None:
        pytest.skip("compute-sanitizer not installed")
    # (racecheck cannot follow a device-armed conditional graph node -- the tool itself dies, with 0 hazards reported;
    # the whole-loop form replays exactly the kernels of the 'replay' form, which it does check)
    parts = "eager,replay,whole,batch,autograd,bruteforce,chamfer" if tool == "memcheck" else "eager,replay,batch,autograd,bruteforce,chamfer"
    cmd = [exe, "--tool", tool, "--error-exitcode", "9", "--print-limit", "20", sys.executable,
           os.path.join(ROOT, "tools", "probes", "sanitize_case.py"), parts]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert "sanitize_case ok" in r.stdout, tail
    assert "ERROR SUMMARY: 0 errors" in r.stdout + r.stderr or "RACECHECK SUMMARY: 0 hazards" in r.stdout + r.stderr, tail
