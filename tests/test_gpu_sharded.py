"""Row-slab sharded solver on >= 2 GPUs (one process per GPU over NCCL) against the single-GPU result
of the same build: same arithmetic, different partitioning => bit-identical field and identical
cycle counts (SURVEY.md section 8e "Validation").  Skipped on a single-GPU box."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_equals_single_gpu(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29400 + world), os.path.join(ROOT, "tests", "sharded_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("SHARDED_RESULT ")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-4000:]
    out = json.loads(line[0][len("SHARDED_RESULT "):])
    for n, res in out.items():
        assert res["sharded_levels"] >= 2, res
        assert res["cycles"] == res["ref_cycles"], res
        assert res["bitwise"], res
        assert res["hist_rel"] <= 1e-9, res
