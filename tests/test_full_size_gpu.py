"""BASELINE.json configs 1, 2, 3 at their stated sizes against digests of the UNMODIFIED reference's MatchLists
(tests/golden/full_<cfg>.json, written by tools/gen_golden_full.py in the build container — the reference needs
8 s / 190 s / 70 s for them).  The GPU result must equal the reference's list record for record: in the reference's
own output order with equal MemCount / MemCollisionCount (MEMS_ORDER_REFERENCE), and as the canonically sorted
set (MEMS_ORDER_CANONICAL and MEMS_ORDER_ANY)."""
import hashlib
import json
import os

import numpy as np
import pytest

import libmems_b200 as mems
from gpu_util import gpu_context
from libmems_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def flat_digest(flat):
    return hashlib.sha256(np.ascontiguousarray(flat, dtype="<i8").tobytes()).hexdigest()


@pytest.mark.parametrize("name", ["c1", "c2", "c3"])
def test_full_size_digest(name):
    gold = json.load(open(os.path.join(GOLD, "full_%s.json" % name)))
    ctx = gpu_context()
    gs = synth.baseline_genomes(name)
    assert [hashlib.sha256(g.tobytes()).hexdigest() for g in gs] == gold["input_sha256"]  # same inputs as the reference saw
    seed = mems.get_seed(gold["seed_weight"])
    assert seed == gold["seed_pattern"]
    mode = mems.MODE_REPEAT if gold["mode"] == "repeat" else mems.MODE_MEMHASH
    smls = ctx.create_smls(gs, seed)
    # the reference's own order and counters
    flat, info = ctx.find_matches(smls, mode=mode, order=mems.ORDER_REFERENCE)
    assert info["n_matches"] == gold["n_matches"] and info["max_run"] <= 1000
    assert info["mem_count"] == gold["mem_count"] and info["collisions"] == gold["collisions"]
    assert flat_digest(flat) == gold["sha256_reference_order"]
    if mode == mems.MODE_MEMHASH:
        flat, info = ctx.find_matches(smls, mode=mode, order=mems.ORDER_CANONICAL)
        assert info["n_matches"] == gold["n_distinct"]
        assert flat_digest(flat) == gold["sha256_canonical"]
        # device order: the same set, no duplicates
        flat, info = ctx.find_matches(smls, mode=mode, order=mems.ORDER_ANY)
        got = sorted(mems.flat_to_matches(flat))
        assert len(got) == gold["n_distinct"] and synth.matchlist_digest(got) == gold["sha256_canonical"]
        assert int(sum(m[1] for m in got)) == gold["sum_length"]
    ctx.close()
