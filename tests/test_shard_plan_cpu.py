"""Host-side sharding plan of the multi-GPU path, exercised with 2 gloo ranks on CPU: every rank derives the
same sequence blocks, the same key-range owners and the same exchange plan (count matrix, slice offsets into the
peers' exchange windows) from the gathered histograms (no GPU needed)."""
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
    # the exchange as the library plans it from the GATHERED histograms (what mems_find_matches_sharded does): every
    # rank then "sends" its records — (rank, bucket, serial) triples in partitioned order — to the offsets of the plan
    rows = [torch.zeros(256, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(rows, torch.from_numpy(hist.copy()))
    hist_all = np.stack([r.numpy() for r in rows]).astype(np.uint32)
    owners2 = m.shard_bucket_owners(hist_all.sum(axis=0).astype(np.uint64), world)
    counts, src, dst, max_recv = m.shard_exchange_plan(hist_all, world, rank, owners2)
    records = np.concatenate([np.stack([np.full(hist[b], rank), np.full(hist[b], b), np.arange(hist[b])], axis=1)
                              for b in range(256)]).astype(np.int64)  # partitioned order = bucket order
    for p in range(world):
        n = int(counts[rank, p])
        np.save(os.path.join(out_dir, "x_%d_to_%d.npy" % (rank, p)),
                np.concatenate([[int(dst[p]), int(max_recv)], records[int(src[p]):int(src[p]) + n].ravel()]))
    np.save(os.path.join(out_dir, "plan%d.npy" % rank), np.concatenate([counts.ravel().astype(np.int64), owners2.astype(np.int64)]))
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


def test_two_rank_exchange_plan_tiles_the_windows(tmp_path):
    """Simulated exchange: writing every rank's slices at the planned offsets fills each receive region exactly —
    no gap, no overlap, slices in sender order, and every record lands on the owner of its bucket."""
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    plans = [np.load(tmp_path / ("plan%d.npy" % k)) for k in range(world)]
    assert np.array_equal(plans[0], plans[1])  # identical on every rank
    counts = plans[0][:world * world].reshape(world, world)
    owners = plans[0][world * world:]
    for p in range(world):
        n_recv = int(counts[:, p].sum())
        region = np.full((n_recv, 3), -1, np.int64)
        filled = np.zeros(n_recv, bool)
        for q in range(world):
            x = np.load(tmp_path / ("x_%d_to_%d.npy" % (q, p)))
            at, max_recv, recs = int(x[0]), int(x[1]), x[2:].reshape(-1, 3)
            assert max_recv == counts.sum(axis=0).max()
            assert len(recs) == counts[q, p]
            assert not filled[at:at + len(recs)].any()
            region[at:at + len(recs)] = recs
            filled[at:at + len(recs)] = True
        assert filled.all()
        assert np.all(owners[region[:, 1]] == p)  # only this rank's key range
        assert np.all(np.diff(region[:, 0]) >= 0)  # slices in sender order
        for q in range(world):
            mine = region[region[:, 0] == q]
            assert np.all(np.diff(mine[:, 1]) >= 0)  # each sender's slice still in bucket order


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
