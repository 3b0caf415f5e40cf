"""N > 1: traces sharded over ranks, one small all-reduce between the two phases.
CPU (gloo, world_size 2): the sharding / exchange algebra with the numpy oracle.
GPU (nccl, 2 GPUs when present): the C-ABI plan API end to end."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

WORKER = os.path.join(ROOT, "tests", "_dist_worker.py")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _launch(mode, nproc, timeout):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), WORKER, mode]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)


def test_sharded_exchange_algebra_gloo_world2():
    r = _launch("gloo-oracle", 2, 600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "GLOO_ORACLE_OK" in r.stdout


@pytest.mark.gpu
def test_sharded_plan_api_nccl_world2(gpu_lib):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    r = _launch("nccl-plan", 2, 900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "NCCL_PLAN_OK" in r.stdout
