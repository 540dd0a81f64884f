"""Multi-GPU parity (needs >= 2 GPUs; run with `gpurun --gpus 2|4|8`): one process per GPU,
NCCL halo / current-sum / particle exchange, each rank checked against the multi-rank oracle."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _run(name, world):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "multi_worker.py"), name],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=600)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{out[-3000:]}"


@pytest.mark.parametrize("name", ["thermal2d_x", "thermal2d_y", "thermal1d", "thermal3d", "reflect2d", "foil2d", "laser2d", "solver2d", "laser2d_y", "mixed2d",
                                  "cpml2d", "cpml2d_y", "cpml3d", "window2d"])   # window1d / window3d exist in parity_check.make_deck; not yet run on two GPUs
def test_two_ranks(name):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(name, 2)


def test_rebalance_two_ranks():
    """epb_redistribute (balance_workload's data movement): re-cut slabs in mid-run, nothing lost, physics unchanged"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run("rebalance2d", 2)


def test_rebalance_single_rank():
    """the same machinery on one rank (every cell and particle 'moves' to the same rank through the new state)"""
    if _ngpu() < 1:
        pytest.skip("needs a GPU")
    _run("rebalance2d", 1)


@pytest.mark.parametrize("name", ["thermal2d_xy", "foil2d", "rebalance2d", "foil2d_xy", "cpml2d"])
def test_four_ranks(name):
    if _ngpu() < 4:
        pytest.skip("needs 4 GPUs")
    _run(name, 4)


def test_eight_ranks_3d():
    if _ngpu() < 8:
        pytest.skip("needs 8 GPUs")
    _run("thermal3d", 8)
