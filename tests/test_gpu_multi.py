"""Multi-GPU parity inside the driver-run suite: tests/multi_gpu_worker.py under torchrun on 2 ranks and on every
visible GPU (marker-sharded products / LOCO / diagonal <= 1e-10, PCG iteration counts equal, tau / alpha <= 1e-6, the
sample-tile-sharded dense GRM, and a rank-sharded step-2 scan equal to the single-rank table).  Skipped with < 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import ctypes
    try:
        cudart = ctypes.CDLL("libcudart.so.12")
        n = ctypes.c_int(0)
        return n.value if cudart.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0
    except OSError:
        return 0


def _run(world):
    port = 29600 + (os.getpid() + 7 * world) % 300
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + "\n" + r.stderr)[-4000:]
    assert r.returncode == 0, tail
    assert r.stdout.count("-> OK") == world, tail


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
def test_two_ranks_match_the_oracle():
    _run(2)


@pytest.mark.skipif(_n_gpus() < 3, reason="needs > 2 GPUs")
def test_all_visible_gpus_match_the_oracle():
    _run(_n_gpus())
