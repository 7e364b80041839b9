"""Slab decomposition with one process per GPU over NCCL (hh_create_slab_nccl): needs >= 2 GPUs, skipped otherwise
(tests/test_gpu_slab.py runs the same kernels with in-process slabs on one GPU)."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_slab_solve_over_nccl_matches_whole_grid(gpu_pkg):
    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    nproc = 4 if ngpu >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "slab_nccl_check.py"), "65"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SLAB_NCCL_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
