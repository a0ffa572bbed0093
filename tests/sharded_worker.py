"""One rank of the sharded (multi-GPU) parity check.  Launch with torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/sharded_worker.py [n_genomes length weight]
Every rank extracts its block of genomes, the ranks exchange seed ranges and hits over NCCL, and rank 0 checks the
union of all ranks' matches against the oracle (small sizes) or against the single-GPU result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import libmems_b200 as mems  # noqa: E402
from libmems_b200 import synth  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # host channel for the NCCL id and for gathering results
    n_genomes = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    length = int(sys.argv[2]) if len(sys.argv) > 2 else 60_000
    weight = int(sys.argv[3]) if len(sys.argv) > 3 else 15
    seed = mems.get_seed(weight)
    ctx = mems.Context(local)
    uid = [mems.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = ctx.create_comm(uid[0], rank, world)
    # several calls on one communicator: the exchange windows are reused, grown (second call is larger) and reused again
    for call, (n_len, g_seed) in enumerate([(length, 11), (length * 2, 12), (length, 13)]):
        gs = synth.genome_family(n_genomes, n_len, seed=g_seed)
        first, count = mems.shard_sequence_range(n_genomes, rank, world)
        seqs = [g if first <= i < first + count else None for i, g in enumerate(gs)]
        flat, info = ctx.find_matches_sharded(comm, seqs, [len(g) for g in gs], seed, order=mems.ORDER_CANONICAL)
        mine = mems.flat_to_matches(flat)
        gathered = [None] * world
        dist.all_gather_object(gathered, (mine, info["n_hits"]))
        if rank == 0:
            union = [m for part, _ in gathered for m in part]
            assert len(union) == len(set(union)), "ranks returned overlapping matches"
            hits = sum(h for _, h in gathered)
            if n_len <= 400_000:
                from checkers import Oracle
                want, winfo = Oracle().find_matches(0, gs, seed)
                assert sorted(union) == sorted(set(want)), "sharded MatchList differs from the oracle"
                assert hits == winfo["hits"]
            smls = ctx.create_smls(gs, seed)
            single, sinfo = ctx.find_matches(smls, order=mems.ORDER_CANONICAL)
            for sml in smls:
                sml.close()
            assert sorted(union) == mems.flat_to_matches(single), "sharded MatchList differs from the single-GPU one"
            assert hits == sinfo["n_hits"]
            print("sharded ok: call=%d world=%d genomes=%d x %d matches=%d hits=%d per-rank matches=%s" %
                  (call, world, n_genomes, n_len, len(union), hits, [len(p) for p, _ in gathered]))
    # a failure local to ONE rank (a '-' in its block) must come back as the same error on EVERY rank instead of
    # leaving the others in the next collective; a fresh communicator afterwards works as before
    gs = synth.genome_family(n_genomes, 20_000, seed=14)
    first, count = mems.shard_sequence_range(n_genomes, world - 1, world)
    bad = bytearray(gs[first].tobytes())
    bad[len(bad) // 2] = ord("-")
    gs[first] = np.frombuffer(bytes(bad), dtype=np.uint8)
    first, count = mems.shard_sequence_range(n_genomes, rank, world)
    seqs = [g if first <= i < first + count else None for i, g in enumerate(gs)]
    try:
        ctx.find_matches_sharded(comm, seqs, [len(g) for g in gs], seed)
        code = 0
    except mems.MemsError as e:
        code = e.code
    codes = [None] * world
    dist.all_gather_object(codes, code)
    assert codes == [2] * world, "every rank must report MEMS_ERR_GAP, got %s" % codes
    if rank == 0:
        print("sharded ok: gap on the last rank's block reported by all %d ranks" % world)
    dist.barrier()
    comm.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
