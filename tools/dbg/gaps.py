import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import libmems_b200 as mems
from libmems_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
gs = synth.baseline_genomes(name)
W = synth.BASELINE_WORKLOADS[name]
dev = [torch.from_numpy(g).cuda() for g in gs]
stream = torch.cuda.Stream()
ctx = mems.Context(0, stream=stream.cuda_stream)
seed = mems.get_seed(W[2])
mode = mems.MODE_REPEAT if W[3] == 'repeat' else mems.MODE_MEMHASH
bufs = [(d.data_ptr(), d.numel()) for d in dev]
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(stream); return e
for _ in range(3):
    s = ctx.create_smls(bufs, seed); f, i = ctx.find_matches(s, mode=mode)
    for x in s: x.close()
K = 10
tc = tf = 0.0
for _ in range(K):
    torch.cuda.synchronize()
    a = ev(); t0 = time.perf_counter()
    s = ctx.create_smls(bufs, seed)
    t1 = time.perf_counter(); b = ev()
    torch.cuda.synchronize(); t1s = time.perf_counter()
    c = ev(); t2 = time.perf_counter()
    f, i = ctx.find_matches(s, mode=mode)
    t3 = time.perf_counter(); d = ev()
    torch.cuda.synchronize()
    tc += a.elapsed_time(b); tf += c.elapsed_time(d)
    for x in s: x.close()
print("create: device span %.3f ms (host call %.3f ms), find: device span %.3f ms (host call %.3f ms)" % (tc / K, (t1 - t0) * 1e3, tf / K, (t3 - t2) * 1e3))
ctx.profile_reset(); ctx.profile_enable(True)
for _ in range(K):
    s = ctx.create_smls(bufs, seed)
    for x in s: x.close()
p = ctx.profile(); print("create kernels: %.3f ms" % (sum(v['ms'] for v in p.values()) / K), {k: round(v['ms'] / K, 3) for k, v in p.items()})
