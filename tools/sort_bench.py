#!/usr/bin/env python3
"""Time the SML build stages (pack, planes, extract, radix passes) on the GPU for one workload — the numbers behind the
roofline table in DESIGN.md and the quick check after a change to those kernels — and, beside them, the yardstick SURVEY.md
§7-4 names: cub::DeviceRadixSort::SortPairs from the CUDA toolkit on the same number of pairs and key bits
(build/cub_sort_bench, test tooling only; built by __graft_entry__.build()).

    python tools/sort_bench.py [n_genomes] [length] [weight]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import libmems_b200 as mems  # noqa: E402
from libmems_b200 import synth  # noqa: E402

n_genomes = int(sys.argv[1]) if len(sys.argv) > 1 else 8
length = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
weight = int(sys.argv[3]) if len(sys.argv) > 3 else 15
rng = np.random.default_rng(1)
gs = [synth.random_genome(length, rng) for _ in range(n_genomes)]
dev = [torch.from_numpy(g).cuda() for g in gs]
ctx = mems.Context(0)
seed = mems.get_seed(weight)
w = mems.get_seed_weight(seed)
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
ctx.profile_enable(False)
out = {k: {"ms_per_launch": v["ms"] / v["launches"], "launches_per_step": v["launches"] / steps,
           "gbs": v["bytes"] / v["ms"] / 1e6 if v["ms"] > 0 and v["bytes"] else None} for k, v in prof.items()}
res = {"genomes": n_genomes, "length": length, "weight": w, "key_bits": 2 * w + 1, "kernels": out}
n_pairs = sum(len(g) - mems.get_seed_length(seed) + 1 for g in gs)
del dev
ctx.close()
torch.cuda.empty_cache()
exe = os.path.join(ROOT, "build", "cub_sort_bench")
if os.path.exists(exe):
    r = subprocess.run([exe, str(n_pairs), str(2 * w + 1), "10"], capture_output=True, text=True)
    try:
        cub = json.loads(r.stdout.strip().splitlines()[-1])
        ours = out.get("radix_pass")
        cub["ours_ms_per_pass"] = ours["ms_per_launch"] if ours else None
        cub["ours_ms_per_sort"] = ours["ms_per_launch"] * ours["launches_per_step"] if ours else None
        res["cub_yardstick"] = cub
    except (ValueError, IndexError):
        res["cub_yardstick"] = {"error": (r.stderr or r.stdout)[-300:]}
else:
    res["cub_yardstick"] = {"error": "build/cub_sort_bench not built"}
print(json.dumps(res))
