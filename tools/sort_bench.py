#!/usr/bin/env python3
"""Time the SML build stages (pack, extract, radix passes) on the GPU for one workload: the numbers behind the
roofline table in DESIGN.md and the quick check after a change to those kernels."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import libmems_b200 as mems
from libmems_b200 import synth

n_genomes = int(sys.argv[1]) if len(sys.argv) > 1 else 8
length = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
weight = int(sys.argv[3]) if len(sys.argv) > 3 else 15
rng = np.random.default_rng(1)
gs = [synth.random_genome(length, rng) for _ in range(n_genomes)]
dev = [torch.from_numpy(g).cuda() for g in gs]
ctx = mems.Context(0)
seed = mems.get_seed(weight)
bufs = [(d.data_ptr(), d.numel()) for d in dev]
for _ in range(3):
    for s in ctx.create_smls(bufs, seed):
        s.close()
ctx.profile_reset()
ctx.profile_enable(True)
steps = 5
for _ in range(steps):
    for s in ctx.create_smls(bufs, seed):
        s.close()
prof = ctx.profile()
out = {k: {"ms_per_launch": v["ms"] / v["launches"], "launches_per_step": v["launches"] / steps,
           "gbs": v["bytes"] / v["ms"] / 1e6 if v["ms"] > 0 and v["bytes"] else None} for k, v in prof.items()}
print(json.dumps({"genomes": n_genomes, "length": length,
                  "weight": weight, "kernels": out}))
