import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import libmems_b200 as mems
from libmems_b200 import synth
from checkers import Oracle
orc = Oracle()
ctx = mems.Context(0)
def run(tag, gs, seed, mode=0):
    want, _ = orc.find_matches(mode, gs, seed)
    flat, info = ctx.find_matches(ctx.create_smls(gs, seed), mode=mode, order=mems.ORDER_CANONICAL)
    got = mems.flat_to_matches(flat)
    w, g = set(want), set(got)
    print(tag, "want", len(w), "got", len(g), "only_want", len(w - g), "only_got", len(g - w))
    for m in sorted(w - g)[:4]: print("   W", m)
    for m in sorted(g - w)[:4]: print("   G", m)
seed = mems.get_seed(15)
run("pair_fwd", synth.genome_family(2, 3000, seed=1, n_indels=0), seed)
run("pair", synth.genome_family(2, 12000, seed=1, n_indels=2), seed)
run("three", synth.genome_family(3, 12000, seed=2, n_indels=2), seed)
run("twelve", synth.genome_family(12, 12000, seed=75, n_indels=4, max_indel=25), seed)
g = synth.genome_family(1, 3000, seed=8)[0]
run("ident", [g, g], seed)
run("rc", [g, synth.revcomp(g)], seed)
