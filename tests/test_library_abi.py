"""CPU-side checks of the product library: it builds for sm_100a, loads, exports every symbol that
include/mgb200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def libpath():
    from hpcclassmultigridproject_b200 import _build
    return _build.build()


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "mgb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mgb200_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported(libpath):
    lib = ctypes.CDLL(libpath)
    names = declared_symbols()
    assert len(names) >= 30
    for nm in names:
        assert hasattr(lib, nm), f"{nm} declared in include/mgb200.h but not exported"


def test_library_contains_sm100a_code_only(libpath):
    out = subprocess.run(["cuobjdump", "-lelf", libpath], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback(libpath):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import hpcclassmultigridproject_b200 as mg
    with pytest.raises(mg.MgError, match="no CUDA device|CUDA"):
        mg.Solver(64, -4e-4, 1e-3, 1 / 64, 1e-6)


def test_product_never_touches_the_oracle():
    """only tests/, bench.py and __graft_entry__.py may reference oracle/"""
    pkg = os.path.join(ROOT, "hpcclassmultigridproject_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and "oracle" not in txt, (dirpath, f)
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "oracle" not in open(p).read().lower()


def test_options_struct_layout_matches_header(libpath):
    import hpcclassmultigridproject_b200 as mg
    o = mg.api.default_options()
    assert o.struct_size == ctypes.sizeof(mg.Options)
    assert (o.shape, o.niter, o.coarse_maxit, o.max_cycle) == (1, 3, 1000, 50)    # multigrid.cpp:41,60,94
    assert o.coarse_tol == 1e-5 and o.correct_towers == 0


def test_host_affinity_helper_is_harmless_without_topology():
    """hpcclassmultigridproject_b200/affinity.py: cpulist parser; no GPU / one NUMA node -> nothing is changed, nothing raised"""
    import os
    from hpcclassmultigridproject_b200.affinity import bind_near_gpu, parse_cpulist
    assert parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert parse_cpulist("") == []
    before = os.sched_getaffinity(0)
    r = bind_near_gpu(0)
    assert r is None or set(os.sched_getaffinity(0)) <= set(before)
    os.sched_setaffinity(0, before)
