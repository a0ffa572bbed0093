import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import libmems_b200 as mems
from libmems_b200 import synth
from checkers import Oracle
orc = Oracle()
seed = mems.get_seed(15)
L = mems.get_seed_length(seed)
gs = synth.genome_family(12, 12000, seed=75, n_indels=4, max_indel=25)
m = (12, 175, 7316, 0, 0, 0, 7324, 7324, 7375, 7281, 7358, 7307, 7379, 7298)
starts = m[2:]
members = [(g, st - 1) for g, st in enumerate(starts) if st]
mask = ~np.uint64(0) << np.uint64(64 - 30)
def W(k):
    ref = None
    for g, p in members:
        q = p + k
        if q < 0 or q > len(gs[g]) - L: return False
        mer = int(orc.seed_mers(gs[g], seed, [q])[1][0])
        t = (mer >> 1 << 1, mer & 1)
        if ref is None: ref = t
        elif t != ref: return False
    return True
row = "".join("1" if W(k) else "0" for k in range(-30, 200))
print(row)
# per-base agreement
code = {65:0, 67:1, 71:2, 84:3}
def base(g, x): return code[int(gs[g][x])]
agree = []
for b in range(-30, 230):
    vals = set(base(g, p + b) for g, p in members)
    agree.append(len(vals) == 1)
pat = [(seed >> (L - 1 - o)) & 1 for o in range(L)]
row2 = "".join("1" if all(agree[k + 30 + o] for o in range(L) if pat[o]) else "0" for k in range(-30, 200))
print(row2)
print(row == row2)
