"""Live cross-check of the oracle against the reference's OWN code (oracle/live_check.py via oracle/refshim): pruned-weight
index selection bit-exact, forward outputs and parameter gradients of freshly drawn pruned networks equal.  Needs
/root/reference, which exists in the builder container only -- skipped on the GPU box, where tests/golden stands in."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/pdm"), reason="reference tree not present (GPU box)")
def test_oracle_matches_reference_code_live():
    # subprocess: the shim registers a fake `diffusers` in sys.modules, which must not leak into the other tests
    r = subprocess.run([sys.executable, "-m", "oracle.live_check"], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "live check ok" in r.stdout
