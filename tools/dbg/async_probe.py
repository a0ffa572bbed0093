#!/usr/bin/env python3
"""Builder probe: does the asynchronous delivery of large MatchLists overlap the next call (and does the arena stop
growing)?  python tools/dbg/async_probe.py [genomes] [length]   (MEMS_TRACE_SLOW=1 shows driver calls of the allocator)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
import libmems_b200 as mems  # noqa: E402
from libmems_b200 import synth  # noqa: E402

g = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
gs = synth.genome_family(g, n, seed=3)
dev = [torch.from_numpy(x).cuda() for x in gs]
ctx = mems.Context(0)
seed = mems.get_seed(15)
bufs = [(d.data_ptr(), d.numel()) for d in dev]


def step(wait):
    smls = ctx.create_smls(bufs, seed)
    r, info = ctx.find_matches(smls, wait=wait)
    for s in smls:
        s.close()
    return r, info


for wait in (True, False, True, False):
    pend = None
    for _ in range(3):
        pend, info = step(wait)
    if not wait:
        pend.wait()
    torch.cuda.synchronize()
    walls = []
    t_all = time.perf_counter()
    for _ in range(6):
        t0 = time.perf_counter()
        prev = pend
        pend, info = step(wait)
        prev = None
        walls.append(1e3 * (time.perf_counter() - t0))
    if not wait:
        pend.wait()
    torch.cuda.synchronize()
    total = 1e3 * (time.perf_counter() - t_all) / 6
    nflat = pend.n_flat if not wait else len(pend)
    print("wait=%s: %.2f ms/step, host per step %s, matches %d, records %.1f MB" %
          (wait, total, " ".join("%.1f" % w for w in walls), info["n_matches"], nflat * 8 / 1e6), flush=True)
