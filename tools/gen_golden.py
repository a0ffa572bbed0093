#!/usr/bin/env python3
"""Generate tests/golden/*.json from the UNMODIFIED reference (oracle/_ref/libmems_ref.so).

Run in the build container only (the reference sources are not on the GPU box):

    make -C oracle all && python tools/gen_golden.py

Each fixture stores the exact inputs (ASCII sequences), the seed pattern and the reference's own
outputs: forward/canonical seed mers, the sorted mer list, and MatchLists in the reference's output
order.  Tests compare the C restatement (oracle/) and the CUDA path against these files.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from checkers import Reference  # noqa: E402
from libmems_b200 import synth  # noqa: E402

R = Reference()
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def s(a):
    return a.tobytes().decode() if isinstance(a, np.ndarray) else a


def dump(name, obj):
    with open(os.path.join(OUT, name), "w") as f:
        json.dump(obj, f, separators=(",", ":"))
    print("wrote", name)


# --- seed table facts (SeedMasks.h) -------------------------------------------------------------
seeds = {}
for w in range(3, 32):
    for r in range(0, 6):
        sd = R.get_seed(w, r)
        seeds["%d,%d" % (w, r)] = [sd, R.seed_length(sd), R.seed_weight(sd)]
defaults = {str(n): R.default_seed_weight(n) for n in
            [0, 1, 10, 31, 32, 33, 1000, 4096, 65536, 1000000, 5000000, 100000000, 4000000000, 2 ** 40]}
dump("seeds.json", {"seeds": seeds, "default_weight": defaults})

# --- SURVEY Appendix B known-answer sequence + a few more ----------------------------------------
kat = []
cases = [
    ("ACGTTGCATGGACCTAGGATCCAATTGGCCAGTCAGTACA", [(5, 0), (7, 0), (9, 1)]),
    ("acgtNNacgtRYKMSWbdhvACGTTTGACCAGTAGGACCATTAGGACCAGTTTAGACCAGGGATTTACACACAGTTAGACC", [(5, 0), (11, 0), (15, 0)]),
    (s(synth.genome_family(1, 300, seed=11)[0]), [(12, 0), (16, 0), (19, 0), (21, 1), (24, 0), (31, 0)]),
]
for seq, wl in cases:
    for w, r in wl:
        sd = R.get_seed(w, r)
        L = R.seed_length(sd)
        if len(seq) < L:
            continue
        pos = np.arange(len(seq) - L + 1, dtype=np.uint64)
        fwd, dna = R.seed_mers(seq.encode(), sd, pos)
        p, m = R.sml_build(seq.encode(), sd)
        kat.append({"seq": seq, "seed": sd, "L": L, "w": R.seed_weight(sd),
                    "seed_mask": R.last["seed_mask"], "mer_mask": R.last["mer_mask"],
                    "fwd": [int(x) for x in fwd], "dna": [int(x) for x in dna],
                    "sml_pos": [int(x) for x in p], "sml_mer": [int(x) for x in m]})
dump("kat_mers.json", kat)

# --- MatchLists ------------------------------------------------------------------------------------
ml = []


def add(mode, seqs, sd, tag):
    matches, info = R.find_matches(mode, [x.encode() if isinstance(x, str) else x for x in seqs], sd)
    ml.append({"tag": tag, "mode": mode, "seed": sd, "seqs": [s(x) for x in seqs],
               "matches": [list(m) for m in matches], "mem_count": info["mem_count"],
               "collisions": info["collisions"]})


a = "ACGTTGCATGGACCTAGGATCCAATTGGCCAGTCAGTACAGGCTTAACGGATACCGTATTGACCA"
b = a[:30] + "T" + a[31:]
add(0, [a, b], R.get_seed(5), "appendixB_snp_bridged")
add(0, synth.genome_family(2, 3000, seed=21, n_indels=4, max_indel=20), R.get_seed(11), "pair_w11_is_w12")
add(0, synth.genome_family(3, 4000, seed=22, n_indels=4, max_indel=20), R.get_seed(15), "three_w15")
add(0, synth.genome_family(5, 2500, seed=23, n_indels=3, max_indel=20, snp_rate=0.03), R.get_seed(9), "five_w9")
add(0, synth.genome_family(4, 3000, seed=24, n_indels=3, max_indel=20), R.get_seed(16), "four_w16_even")
add(0, synth.genome_family(2, 3000, seed=25, n_indels=3, max_indel=20), R.get_seed(24), "pair_solid24")
add(0, [s(synth.genome_family(1, 50, seed=26)[0]), "ACGT", ""], R.get_seed(15), "ragged_short_and_empty")
add(1, [synth.repeat_genome(8000, seed=27, families=4, copies=5, min_len=60, max_len=300)], R.get_seed(13), "repeat_w13")
add(1, [synth.repeat_genome(12000, seed=28, families=6, copies=7, min_len=60, max_len=400, divergence=0.04)],
    R.get_seed(11), "repeat_w12_drops")
add(2, synth.genome_family(3, 3000, seed=29, n_indels=3, max_indel=20), R.get_seed(13), "pairwise_w13")
dump("matchlists.json", ml)

# --- façade members beyond FindMatches (tests/test_facade_gpu.py, tests/test_table_cpu.py) -------------------------
# MemHash::FindMatchesFromPosition with LogProgress / SetMatchLog attached, MemHash::WriteFile, MemHash::LoadFile
fac = []
for tag, G, n, w, gseed, sp in [("three_w13", 3, 30000, 13, 5, [5000, 100, 20000]), ("pair_w11", 2, 25000, 11, 6, [0, 0]),
                                ("four_w15", 4, 15000, 15, 7, [0, 14000, 3, 20000]), ("pair_w16_past_end", 2, 12000, 16, 9, [100, 50000])]:
    sd = R.get_seed(w)
    gs = synth.genome_family(G, n, seed=gseed, snp_rate=0.03)
    matches, info = R.find_matches_from(gs, sd, sp)
    text = R.mems_write_file(gs, sd)
    lines = [ln for ln in text.split("\n")[2 + 2 * G + 1:] if ln]
    variants = {"as_written": lines, "reversed": lines[::-1], "with_repeats": lines + lines[:10]}
    loads = {}
    for k, v in variants.items():
        ml, il = R.mems_load_file(gs, sd, "\n".join(v) + "\n")
        loads[k] = {"lines": v, "matches": [list(m) for m in ml], "mem_count": il["mem_count"], "collisions": il["collisions"]}
    fac.append({"tag": tag, "weight": w, "seed": sd, "seqs": [s(x) for x in gs], "start_points": sp,
                "matches": [list(m) for m in matches], "mem_count": info["mem_count"], "collisions": info["collisions"],
                "progress": info["progress"], "match_log": info["match_log"], "mems_file": text, "loads": loads})
dump("facade.json", fac)

# --- EliminateOverlaps (Aligner.cpp:62-180) on MemHash results, in the order MemHash returns them -------------------
ov = []
for tag, G, n, w, gseed in [("four_w11", 4, 20000, 11, 7), ("three_w13_inverted", 3, 30000, 13, 5), ("six_w9", 6, 8000, 9, 12), ("pair_w15", 2, 40000, 15, 3)]:
    sd = R.get_seed(w)
    gs = synth.genome_family(G, n, seed=gseed, snp_rate=0.03)
    matches, _ = R.find_matches(0, gs, sd)
    ov.append({"tag": tag, "input": [list(m) for m in matches], "output": [list(m) for m in R.eliminate_overlaps(matches)]})
dump("overlaps.json", ov)
