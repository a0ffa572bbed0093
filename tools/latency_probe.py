#!/usr/bin/env python3
"""Small-problem throughput: pairs of short related sequences (the gap re-anchoring callers' case,
ProgressiveAligner.cpp:589-678) one create + find call per problem versus mems_find_matches_many over batches of problems.
Prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libmems_b200 as mems  # noqa: E402
from libmems_b200 import synth  # noqa: E402

ctx = mems.Context(0)
out = {}
for n in (1000, 10000, 100000):
    seed = mems.get_seed(mems.get_default_seed_weight(n))
    problems = [synth.genome_family(2, n, seed=1 + k, n_indels=2, max_indel=10) for k in range(128)]
    for gs in problems[:5]:
        ctx.find_matches(ctx.create_smls(gs, seed))
    t = time.perf_counter()
    total = 0
    for gs in problems:
        flat, info = ctx.find_matches(ctx.create_smls(gs, seed))
        total += info["n_matches"]
    per_call = (time.perf_counter() - t) / len(problems)
    ctx.find_matches_many(problems, seed)
    t = time.perf_counter()
    reps = 5
    for _ in range(reps):
        res = ctx.find_matches_many(problems, seed)
    per_problem = (time.perf_counter() - t) / (reps * len(problems))
    assert sum(i["n_matches"] for _, i in res) == total
    out["%d_bp_pairs" % n] = {"one_call_per_problem_ms": per_call * 1e3, "many_ms_per_problem": per_problem * 1e3,
                              "problems_per_s_single": 1 / per_call, "problems_per_s_many": 1 / per_problem,
                              "speedup": per_call / per_problem, "problems_per_many_call": len(problems)}
print(json.dumps(out))
