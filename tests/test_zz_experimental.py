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


def _run(tmp_path, name, env_extra, size=()):
    out = str(tmp_path / f"{name}.npz")
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_spmv_variant.py"), out, *map(str, size)], env=env, timeout=300,
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
    ref = _run(tmp_path, "flat", {"DSA_SPMV_BULK": "0", "DSA_SPMV_STEPS": "4"}, size)
    for mode in ("1", "2", "3"):
        got = _run(tmp_path, f"bulk{mode}", {"DSA_SPMV_BULK": mode, "DSA_SPMV_STEPS": "4"}, size)
        for key in ("y", "yt", "ys_k", "ys_v"):
            assert np.array_equal(_bits(ref[key]), _bits(got[key])), (mode, key)
