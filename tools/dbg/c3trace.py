import sys, os, time
sys.path.insert(0, '/root/repo')
import torch, libmems_b200 as mems
from libmems_b200 import synth
gs = synth.baseline_genomes('c3')
dev = [torch.from_numpy(g).cuda() for g in gs]
ctx = mems.Context(0)
seed = mems.get_seed(19)
bufs = [(d.data_ptr(), d.numel()) for d in dev]
for it in range(3):
    t0 = time.perf_counter()
    smls = ctx.create_smls(bufs, seed)
    t1 = time.perf_counter()
    flat, info = ctx.find_matches(smls, mode=mems.MODE_REPEAT)
    t2 = time.perf_counter()
    print("create %.2f ms find %.2f ms replay %.2f" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, info["host_replay_ms"]), file=sys.stderr)
    for s in smls: s.close()
