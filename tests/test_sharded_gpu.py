"""The sharded path on real GPUs: world size 1 in-process (every step except the wire), and — when the box has
two or more GPUs — two ranks under torchrun (tests/sharded_worker.py)."""
import os
import subprocess
import sys

import pytest

import libmems_b200 as mems
from checkers import Oracle
from gpu_util import gpu_context
from libmems_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world_of_one_equals_oracle():
    ctx = gpu_context()
    seed = mems.get_seed(15)
    gs = synth.genome_family(4, 50_000, seed=3)
    comm = ctx.create_comm(mems.comm_unique_id(), 0, 1)
    flat, info = ctx.find_matches_sharded(comm, gs, [len(g) for g in gs], seed, order=mems.ORDER_CANONICAL)
    want, winfo = Oracle().find_matches(0, gs, seed)
    assert mems.flat_to_matches(flat) == sorted(set(want))
    assert info["n_hits"] == winfo["hits"]
    with pytest.raises(mems.MemsError):
        ctx.find_matches_sharded(comm, gs, [len(g) for g in gs], seed, order=mems.ORDER_REFERENCE)
    comm.close()
    ctx.close()


@pytest.mark.parametrize("transport", ["peer_copy", "peer_scatter", "nccl"])
def test_two_ranks_under_torchrun(transport):
    """Two ranks, every way the exchanges can travel: DMA copies into the peers' exchange windows (default), the
    partition kernel scattering straight into them (MEMS_PEER_SCATTER), NCCL send/recv (MEMS_NO_PEER_WINDOWS)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ)
    env.pop("MEMS_PEER_SCATTER", None)
    env.pop("MEMS_NO_PEER_WINDOWS", None)
    if transport == "peer_scatter":
        env["MEMS_PEER_SCATTER"] = "1"
    elif transport == "nccl":
        env["MEMS_NO_PEER_WINDOWS"] = "1"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(ROOT, "tests", "sharded_worker.py"), "5", "60000", "15"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "sharded ok" in r.stdout
