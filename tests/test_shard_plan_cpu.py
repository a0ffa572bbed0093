"""Host-side sharding plan of the multi-GPU path, exercised with 2 gloo ranks on CPU: every rank derives the
same sequence blocks and the same key-range owners from the all-reduced histogram (no GPU needed)."""
import os
import socket
import sys

import numpy as np
import pytest

import libmems_b200 as mems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import libmems_b200 as m
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    n_seqs = 7
    first, count = m.shard_sequence_range(n_seqs, rank, world)
    # local top-digit histogram of this rank's block (skewed like canonical keys: density ~ 2(1-x))
    rng = np.random.default_rng(100 + rank)
    x = 1.0 - np.sqrt(rng.random(200_000 * max(count, 1)))
    hist = np.bincount((x * 256).astype(np.int64).clip(0, 255), minlength=256).astype(np.int64)
    t = torch.from_numpy(hist.copy())
    dist.all_reduce(t)
    owners = m.shard_bucket_owners(t.numpy().astype(np.uint64), world)
    send_counts = np.bincount(owners, weights=hist, minlength=world).astype(np.int64)
    # the count matrix every rank needs for the all-to-all-v
    mat = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(mat, torch.from_numpy(send_counts))
    recv = np.array([int(mat[p][rank]) for p in range(world)])
    np.save(os.path.join(out_dir, "r%d.npy" % rank),
            np.concatenate([[first, count], owners.astype(np.int64), t.numpy(), send_counts, recv]))
    dist.destroy_process_group()


def test_two_rank_plan(tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / ("r%d.npy" % k)) for k in range(world)]
    blocks = [(int(x[0]), int(x[1])) for x in r]
    assert blocks == [(0, 4), (4, 3)]  # contiguous, disjoint, covering 7 sequences
    owners = [x[2:258] for x in r]
    assert np.array_equal(owners[0], owners[1])  # same plan on every rank
    assert np.all(np.diff(owners[0]) >= 0) and owners[0][0] == 0 and owners[0][-1] == world - 1  # contiguous key ranges
    ghist = r[0][258:514]
    share = np.bincount(owners[0], weights=ghist, minlength=world) / ghist.sum()
    assert np.all(np.abs(share - 1.0 / world) < 0.02)  # balanced although the key density is skewed ~15:1
    send = [x[514:514 + world] for x in r]
    recv = [x[514 + world:514 + 2 * world] for x in r]
    for p in range(world):
        for q in range(world):
            assert send[p][q] == recv[q][p]


def test_plan_edge_cases():
    assert mems.shard_sequence_range(3, 0, 8) == (0, 1)
    assert mems.shard_sequence_range(3, 5, 8) == (3, 0)  # more ranks than sequences: empty block
    assert [mems.shard_sequence_range(16, r, 8) for r in range(8)] == [(2 * r, 2) for r in range(8)]
    h = np.zeros(256, np.uint64)
    assert np.all(mems.shard_bucket_owners(h, 4) == 0)  # nothing to balance
    h[7] = 10
    o = mems.shard_bucket_owners(h, 4)
    assert np.all(np.diff(o.astype(int)) >= 0)
    with pytest.raises(mems.MemsError):
        mems.shard_sequence_range(4, 2, 2)
