#!/usr/bin/env python3
"""Sharded anchoring of G synthetic genomes of N bases on all ranks of a torchrun launch (BASELINE configs[4]:
16 x 100 Mbp, w19, 8 GPUs).  Every rank generates only the genomes of its own block (the base genome comes from a
shared seed, genome g's mutations from seed + g), so host memory stays small.  Prints one JSON line on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
        tools/run_sharded.py 16 100000000 19 [steps]
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libmems_b200 as mems  # noqa: E402
from libmems_b200 import synth  # noqa: E402


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_genomes, length, weight = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    seed = mems.get_seed(weight)
    first, count = mems.shard_sequence_range(n_genomes, rank, world)
    base = synth.random_genome(length, np.random.default_rng(12345))
    mine = {}
    for g in range(first, first + count):
        mine[g] = base if g == 0 else synth.mutate(base, np.random.default_rng(12345 + g))
    # every rank needs all lengths: exchange them
    lens_t = torch.zeros(n_genomes, dtype=torch.int64, device="cuda")
    for g, a in mine.items():
        lens_t[g] = len(a)
    dist.all_reduce(lens_t)
    lens = [int(x) for x in lens_t.tolist()]
    del base
    host = {g: torch.from_numpy(a).pin_memory() for g, a in mine.items()}
    ctx = mems.Context(local)
    uid = [mems.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0, device=torch.device("cuda", local))
    comm = ctx.create_comm(uid[0], rank, world)
    seqs = [(host[g].data_ptr(), host[g].numel()) if g in host else None for g in range(n_genomes)]

    def step():
        return ctx.find_matches_sharded(comm, seqs, lens, seed)

    step()  # warm-up (NCCL connections, memory pool)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        flat, info = step()
    torch.cuda.synchronize()
    dist.barrier()
    dt = (time.perf_counter() - t0) / steps
    t = torch.tensor([dt], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([info["n_matches"], info["n_hits"], info["max_run"]], device="cuda", dtype=torch.int64)
    mx = cnt.clone()
    dist.all_reduce(cnt)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    if rank == 0:
        mbp = sum(lens) / 1e6
        print(json.dumps({"workload": "%d x %.0f Mbp, w%d, sharded over %d GPUs" % (n_genomes, length / 1e6, weight, world),
                          "e2e_mbp_per_s": mbp / float(t.item()), "s_per_step": float(t.item()), "total_mbp": mbp,
                          "matches": int(cnt[0]), "hits": int(cnt[1]), "max_run": int(mx[2]), "steps": steps}))
    comm.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
