#!/usr/bin/env python3
"""Pin parity at BASELINE.json's full sizes: run the UNMODIFIED reference (oracle/_ref/libmems_ref.so) on configs
1, 2 and 3 with the repository's generator and store digests of its MatchLists in tests/golden/full_<cfg>.json.

Build container only (about 8 s / 4 min / 1.5 min of single-thread reference time for c1 / c2 / c3):

    make -C oracle all && python tools/gen_golden_full.py [c1 c2 c3]

Stored per config: match count, sum of lengths, MemCount, MemCollisionCount, SHA-256 of the MatchList in the
reference's own output order and of the canonically sorted distinct records, plus the reference's
stage times.  tests/test_full_size_gpu.py compares the GPU result of the same inputs with these digests.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from checkers import Reference  # noqa: E402
from libmems_b200 import synth  # noqa: E402

R = Reference()
OUT = os.path.join(ROOT, "tests", "golden")
for name in (sys.argv[1:] or ["c1", "c3", "c2"]):
    g, n, w, mode, gen_seed = synth.BASELINE_WORKLOADS[name]
    seed = R.get_seed(w)
    gs = synth.baseline_genomes(name)
    t0 = time.time()
    matches, info = R.find_matches(1 if mode == "repeat" else 0, gs, seed)
    canon = sorted(set(matches))
    rec = {"config": name, "genomes": g, "length": n, "generator_seed": gen_seed, "seed_weight": w, "seed_pattern": seed,
           "mode": mode, "input_sha256": [__import__("hashlib").sha256(x.tobytes()).hexdigest() for x in gs],
           "n_matches": len(matches), "n_distinct": len(canon), "sum_length": int(sum(m[1] for m in matches)),
           "mem_count": int(info["mem_count"]), "collisions": int(info["collisions"]),
           "sha256_reference_order": synth.matchlist_digest(matches), "sha256_canonical": synth.matchlist_digest(canon),
           "reference_sml_s": info["sml_s"], "reference_find_s": info["find_s"]}
    with open(os.path.join(OUT, "full_%s.json" % name), "w") as f:
        json.dump(rec, f, indent=1)
    print("wrote full_%s.json: %d matches, %.1f s" % (name, len(matches), time.time() - t0), flush=True)
