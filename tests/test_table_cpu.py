"""MemHash's table as the C-ABI exposes it (mems_table_add / mems_table_matches): host arithmetic, no GPU.  The golden
file holds what the UNMODIFIED reference's MemHash::LoadFile made of the same lines (tools/gen_golden.py)."""
import json
import os

import libmems_b200 as mems

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def parse_like_load_file(lines):
    """Records as MemHash::LoadFile reads them: the first line's starts are read from the beginning of the line
    (MemHash.cpp:280-285), so its length doubles as start 0."""
    recs = []
    for k, ln in enumerate(lines):
        v = [int(x) for x in ln.split()]
        recs.append((len(v) - 1, v[0]) + tuple(v[:-1] if k == 0 else v[1:]))
    return recs


def test_table_add_equals_reference_load_file():
    for case in json.load(open(os.path.join(GOLD, "facade.json"))):
        for name, load in case["loads"].items():
            T = mems.HashTable()
            inserted = [T.add(r, mersize=31) for r in parse_like_load_file(load["lines"])]  # DNA_MER_SIZE before any search
            got = T.matches()
            assert got == [tuple(m) for m in load["matches"]], (case["tag"], name)  # the reference's table order
            assert sum(inserted) == load["mem_count"] and len(inserted) - sum(inserted) == load["collisions"], (case["tag"], name)
            T.clear()
            assert T.matches() == []
