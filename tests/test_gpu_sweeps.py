"""The scaling harnesses (tools/sweeps.py): the grid-size sweep of mg_timer.cu:210-268 and the
GPU-count sweep in the spirit of multigrid_strongsc.cpp:246-262 write the tables the reference's plotters read."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
TOOL = os.path.join(ROOT, "tools", "sweeps.py")


def table(path):
    rows = []
    for line in open(path):
        a, b = line.split("\t")                 # speedupplot.py:38-50 / strongsc_plot.py:51-60: two tab-separated columns
        rows.append((int(a), float(b)))
    return rows


def test_grid_size_sweep_writes_cudatime(tmp_path):
    out = str(tmp_path / "cudatime.txt")
    r = subprocess.run([sys.executable, TOOL, "nsweep", "--nmin", "32", "--nmax", "256", "--steps", "10", "--reps", "1", "--out", out],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    rows = table(out)
    assert [n for n, _ in rows] == [32, 64, 128, 256]
    assert all(0.0 < s < 5.0 for _, s in rows)
    # the reference's line format (mg_timer.cu:262)
    assert r.stdout.count("Time elapsed for grid size") == 4


def test_gpu_count_sweep_writes_strong_scale(tmp_path):
    out = str(tmp_path / "strong_scale.txt")
    r = subprocess.run([sys.executable, TOOL, "strong", "--n", "1024", "--max-gpus", "1", "--steps", "3", "--out", out],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-2000:])
    rows = table(out)
    assert [g for g, _ in rows] == [1] and 0.0 < rows[0][1] < 5.0
