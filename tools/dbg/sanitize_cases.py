"""Small cases through every extension path for compute-sanitizer (memcheck): sequence ends, reverse members, even
weights, non-palindromic patterns, RepeatHash, many(), from-position."""
import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
import libmems_b200 as mems
from libmems_b200 import synth
ctx = mems.Context(0)
rng = np.random.default_rng(3)
for w, rank in ((15, 0), (16, 0), (19, 2), (9, 0)):
    seed = mems.get_seed(w, rank)
    gs = synth.genome_family(4, 6000, seed=w, snp_rate=0.02, n_indels=3, max_indel=20)
    gs.append(synth.revcomp(gs[1]))
    gs.append(gs[0][:40])
    smls = ctx.create_smls(gs, seed)
    for order in (mems.ORDER_ANY, mems.ORDER_REFERENCE):
        flat, info = ctx.find_matches(smls, order=order)
    flat, info = ctx.find_matches(smls, order=mems.ORDER_REFERENCE, start_points=[10, 0, 500, 0, 0, 0])
    smls[0].read(); smls[2].seed_occurrence(); smls[1].find_mer(12345)
    print(w, info["n_matches"])
g = synth.repeat_genome(30000, seed=5, families=5, copies=6, min_len=60, max_len=400)
flat, info = ctx.find_matches([ctx.create_sml(g, mems.get_seed(13))], mode=mems.MODE_REPEAT)
print("repeat", info["n_matches"])
T = synth.random_genome(9000, rng); M = synth.random_genome(60, rng); X = np.concatenate([T, M, T])
ctx.set_test_hooks(walk_budget=1)
flat, info = ctx.find_matches(ctx.create_smls([X, synth.revcomp(X)], mems.get_seed(15)))
ctx.set_test_hooks()
print("long", info["n_matches"])
res = ctx.find_matches_many([synth.genome_family(2, 300 + 50 * k, seed=k) for k in range(20)], mems.get_seed(9))
print("many", sum(i["n_matches"] for _, i in res))
flat, info = ctx.find_matches(ctx.create_smls(synth.genome_family(3, 3000, seed=2), mems.get_seed(11)), mode=mems.MODE_PAIRWISE)
print("pairwise", info["n_matches"])
