"""Multi-GPU parity of the sharded path as a driver-run test: spawns tests/run_sharded_gpu.py under torch.distributed.run on
min(device_count, 8) GPUs (one rank per GPU, NCCL).  On a single-GPU box the same worker runs with one rank: routing, unpack and
the device-side-count pipeline are still exercised, only the peer stores and NCCL are not."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launch(nproc, extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "run_sharded_gpu.py")]
    r = subprocess.run(cmd, env=env, cwd=ROOT, timeout=900, capture_output=True, text=True)
    tail = (r.stdout[-3000:] + "\n" + r.stderr[-6000:])
    if r.returncode != 0 and os.path.isdir(os.path.join(ROOT, "gpurun_out")):   # the whole worker output, for the post-mortem
        with open(os.path.join(ROOT, "gpurun_out", f"sharded_worker_n{nproc}.log"), "w") as f:
            f.write(r.stdout + "\n=====STDERR=====\n" + r.stderr)
    assert r.returncode == 0, tail
    assert f"sharded parity ok on {nproc} GPU(s)" in r.stdout, tail
    return r.stdout


def _ngpus():
    import dsa_b200
    return min(dsa_b200.device_count(), 8)


def test_sharded_parity_all_gpus():
    out = _launch(_ngpus())
    if _ngpus() > 1:
        assert "transport peer-memory" in out, out   # the fused route+push path is the one that ran


@pytest.mark.skipif("_ngpus() < 2")
def test_sharded_parity_two_gpus_nccl_transport():
    _launch(2, {"DSA_DIST_TRANSPORT": "nccl"})
