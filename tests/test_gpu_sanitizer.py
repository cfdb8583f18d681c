"""compute-sanitizer over the fused plan: memcheck / racecheck / synccheck / initcheck on
tools/sanitize_target.py.  NOT part of the default GPU suite: the profiling notes of this pool ask for at most one
sanitizer tool per GPU call, so the test runs only when MGB200_TEST_SANITIZER names ONE tool; the logs of the runs
made while building are kept under profiles/ (r2_sanitizer_*.log)."""
import os
import shutil
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
TOOL = os.environ.get("MGB200_TEST_SANITIZER", "")


@pytest.mark.skipif(TOOL not in ("memcheck", "racecheck", "synccheck", "initcheck"),
                    reason="set MGB200_TEST_SANITIZER=memcheck|racecheck|synccheck|initcheck (one tool per GPU call)")
def test_fused_plan_under_compute_sanitizer():
    cs = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    r = subprocess.run([cs, "--tool", TOOL, "--error-exitcode", "9", sys.executable, os.path.join(ROOT, "tools", "sanitize_target.py"), "64", "256"],
                       capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, tail
    assert "ERROR SUMMARY: 0 errors" in r.stdout + r.stderr, tail
    assert "sanitize_target:" in r.stdout, tail
