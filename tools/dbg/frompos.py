import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import libmems_b200 as mems
from libmems_b200 import synth
from checkers import Reference
R = Reference()
ctx = mems.Context(0)
seed = mems.get_seed(13)
gs = synth.genome_family(3, 30000, seed=5, snp_rate=0.03)
smls = ctx.create_smls(gs, seed)
for sp in ([0,0,0],[5000,0,0],[0,100,0],[0,0,20000],[5000,100,20000],[1,0,0],[10001,0,0]):
    want, wi = R.find_matches_from(gs, seed, sp)
    flat, info = ctx.find_matches(smls, order=mems.ORDER_REFERENCE, start_points=sp)
    got = mems.flat_to_matches(flat)
    w, g = set(want), set(got)
    print(sp, "want", len(want), "got", len(got), "hits", info["n_hits"], "only_want", sorted(w-g)[:3], "only_got", sorted(g-w)[:3])
