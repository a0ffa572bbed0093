import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np, libmems_b200 as mems
from libmems_b200 import synth
ctx=mems.Context(0)
for n in (1000, 10000, 100000):
    gs=synth.genome_family(2,n,seed=1,n_indels=2,max_indel=10)
    seed=mems.get_seed(mems.get_default_seed_weight(n))
    for _ in range(5):
        s=ctx.create_smls(gs,seed); ctx.find_matches(s)
    t=time.perf_counter()
    for _ in range(50):
        s=ctx.create_smls(gs,seed); f,i=ctx.find_matches(s)
    dt=(time.perf_counter()-t)/50
    print(n, "bp pair: %.3f ms per create+find"%(dt*1e3), i['n_matches'], ctx.launch_count())
