"""Opt-in checks of experimental kernels that have not run on hardware yet (skipped unless DSA_EXPERIMENTAL=1, so that an
unvalidated kernel can never turn the parity suite red or poison its CUDA context).  Each variant runs in its own process.

    DSA_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_zz_experimental.py -q -m gpu -s
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("DSA_EXPERIMENTAL") != "1", reason="experimental kernels: set DSA_EXPERIMENTAL=1")]


def _run(tmp_path, name, env_extra, size=(), script="run_spmv_variant.py"):
    out = str(tmp_path / f"{name}.npz")
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", script), out, *map(str, size)], env=env, timeout=300,
                       capture_output=True, text=True)
    print(r.stdout.strip())
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(out)


def _bits(a):
    return a.view(np.int64) if a.dtype == np.float64 else a


@pytest.mark.parametrize("size", [(100_000, 100_000, 10_000_000), (20_000, 30_000, 2_000_000), (300, 200, 3_000)])
def test_spmv_bulk_variants_are_bit_identical_to_flat(tmp_path, size):
    """k_spmv_bulk keeps k_spmv_flat<., 4>'s chunking and arithmetic: same bits, dense and sparse x, both orientations.
    The small size has capacity < tile for some variants: they must fall back to the flat kernel."""
    ref = _run(tmp_path, "flat", {"DSA_SPMV_BULK": "0", "DSA_SPMV_STEPS": "4"}, size)   # modes 1-5 and 8 were validated in round 1
    for mode in os.environ.get("DSA_EXP_SPMV_MODES", "1,2,3,4,5,8").split(","):
        got = _run(tmp_path, f"bulk{mode}", {"DSA_SPMV_BULK": mode, "DSA_SPMV_STEPS": "4"}, size)
        for key in ("y", "yt", "ys_k", "ys_v"):
            assert np.array_equal(_bits(ref[key]), _bits(got[key])), (mode, key)


# default = two streams + one-pass scan (validated at the end of round 1); every variant below must leave bit-identical layouts
DEFAULT = {"DSA_TWO_STREAMS": "1", "DSA_SCAN_ONEPASS": "1", "DSA_SPMV_BULK": "0", "DSA_ILP": "0"}
SWITCHES = [{"DSA_ILP": "2"}, {"DSA_ILP": "4"}, {"DSA_TWO_STREAMS": "0"}, {"DSA_SCAN_ONEPASS": "0"}]


@pytest.mark.parametrize("size", [(100_000, 10_000_000, 1_000_000, 10), (3_000, 200_000, 50_000, 6)])
def test_update_switches_leave_the_same_layout(tmp_path, size):
    """DSA_ILP only changes how many ops a thread of the per-op kernels handles, DSA_TWO_STREAMS which stream the twin
    orientation's kernels run on, DSA_SCAN_ONEPASS how prefix sums are computed: both layouts, the column maps and the SpMV
    result must be bit-identical to the default run (the step time of every variant is printed)."""
    ref = _run(tmp_path, "default", DEFAULT, size, script="run_update_variant.py")
    for i, sw in enumerate(SWITCHES):
        got = _run(tmp_path, f"sw{i}", dict(DEFAULT, **sw), size, script="run_update_variant.py")
        assert str(ref["col"]) == str(got["col"]) and str(ref["row"]) == str(got["row"]) and int(ref["nnz"]) == int(got["nnz"]), sw
        assert np.array_equal(_bits(ref["y"]), _bits(got["y"])), sw


@pytest.mark.parametrize("sw", SWITCHES[:2])   # the unvalidated ones
def test_parity_suite_under_switch(sw):
    """the parity tests (oracle comparisons) with one switch on"""
    env = dict(os.environ, **sw)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-m", "gpu", "-k", "not full_size"],
                       env=env, timeout=300, capture_output=True, text=True, cwd=ROOT)
    print(sw, r.stdout[-300:])
    assert r.returncode == 0, r.stdout[-3000:]
