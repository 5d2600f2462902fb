"""The N > 1 path.  CPU: world_size-2 (and 4) gloo runs of the drop-in's host side in process-rank mode.
GPU: the NCCL face-halo path on 2 / 4 / 8 GPUs of one box, every operator bit-exact against the oracle's emulation of the
same npX x npY reference run (skipped on boxes with a single GPU)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "mp_worker.py")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _launch(mode, nproc, timeout):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), WORKER, mode]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)
    assert r.returncode == 0 and ("MP_WORKER_OK %s %d" % (mode, nproc)) in r.stdout, (r.stdout[-3000:], r.stderr[-6000:])


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_host_side_process_ranks_gloo(nproc):
    _launch("host", nproc, 600)


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_multi_gpu_parity(nproc):
    if _gpu_count() < nproc:
        pytest.skip("needs %d GPUs on this box" % nproc)
    _launch("gpu", nproc, 1500)
