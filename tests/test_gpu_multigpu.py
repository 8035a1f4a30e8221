"""Multi-GPU equality: the row-partitioned propagation (fused peer-store all-gather) and the
block-sharded spreading must reproduce the single-GPU result on every rank.  Needs >= 2 GPUs on the box; with one
GPU the tests skip, and the same equalities are checked inside `bench.py --gpus N` (field "parity_vs_1gpu"), which
the driver runs at N = 2, 4, 8."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worlds():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return [w for w in (2, 4, 8) if w <= n]


def _torchrun(world, script, *args, timeout=900):
    port = 29700 + (os.getpid() + world) % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", script), *args]
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("shape", ["ml-1m", "ml-20m"])
def test_row_partitioned_propagation_equals_single_gpu(world, shape):
    if world not in _worlds():
        pytest.skip(f"needs {world} GPUs")
    r = _torchrun(world, "check_multigpu.py", shape)
    assert r.returncode == 0 and "MULTIGPU_OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_spreading_equals_single_gpu(world):
    if world not in _worlds():
        pytest.skip(f"needs {world} GPUs")
    r = _torchrun(world, "check_multigpu_spread.py", "ml-1m")
    assert r.returncode == 0 and "MULTIGPU_SPREAD_OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])


@pytest.mark.parametrize("world", [2, 4, 8])
def test_distributed_training_step_equals_single_gpu(world):
    if world not in _worlds():
        pytest.skip(f"needs {world} GPUs")
    r = _torchrun(world, "check_multigpu_train.py", "ml-100k")
    assert r.returncode == 0 and "MULTIGPU_TRAIN_OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
