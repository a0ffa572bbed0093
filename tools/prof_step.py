#!/usr/bin/env python3
"""One step of a BASELINE workload for ncu: two warm-up steps, then one step (pack, planes, extract, radix passes, run
scan, hits, extension, emit) with device-resident inputs.   python tools/prof_step.py [c1|c2|c3]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import libmems_b200 as mems  # noqa: E402
from libmems_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
g, n, w, mode, _ = synth.BASELINE_WORKLOADS[name]
gs = synth.baseline_genomes(name)
dev = [torch.from_numpy(x).cuda() for x in gs]
ctx = mems.Context(0)
seed = mems.get_seed(w)
bufs = [(d.data_ptr(), d.numel()) for d in dev]
for _ in range(3):
    smls = ctx.create_smls(bufs, seed)
    flat, info = ctx.find_matches(smls, mode=mems.MODE_REPEAT if mode == "repeat" else mems.MODE_MEMHASH)
    for s in smls:
        s.close()
print(name, info["n_matches"], "matches", ctx.launch_count(), "launches in 3 steps")
