"""Randomised parity sweep: many small inputs of random shape (sequence count, ragged and tiny lengths, seed weight and
rank incl. even weights and the non-palindromic table entries, SNP / indel / inversion rates, mode) through the C-ABI
against the oracle — the corners no hand-written case names."""
import numpy as np
import pytest

import libmems_b200 as mems
from checkers import Oracle
from gpu_util import gpu_context
from libmems_b200 import synth

pytestmark = pytest.mark.gpu


def random_case(rng):
    w = int(rng.integers(5, 26))
    seed = mems.get_seed(w, int(rng.integers(0, 3)))
    G = int(rng.integers(2, 10))
    n = int(rng.choice([int(rng.integers(1, 80)), int(rng.integers(80, 2000)), int(rng.integers(2000, 30000))]))
    gs = synth.genome_family(G, n, seed=int(rng.integers(1, 1 << 30)), snp_rate=float(rng.choice([0.0, 0.005, 0.02, 0.08])),
                             n_indels=int(rng.integers(0, 6)) if n >= 300 else 0, max_indel=int(rng.integers(1, 40)))
    if rng.random() < 0.3:
        gs[int(rng.integers(0, G))] = synth.revcomp(gs[int(rng.integers(0, G))])
    if rng.random() < 0.3:
        k = int(rng.integers(0, G))
        gs[k] = gs[k][:int(rng.integers(0, len(gs[k]) + 1))]  # ragged, possibly shorter than the seed or empty
    if rng.random() < 0.2:
        gs.append(np.concatenate([gs[0], gs[0]]))  # every seed twice in one sequence
    return seed, gs


@pytest.mark.parametrize("block", range(6))
def test_random_memhash_and_pairwise(block):
    ctx, orc = gpu_context(), Oracle()
    rng = np.random.default_rng(9000 + block)
    for it in range(40):
        seed, gs = random_case(rng)
        mode = mems.MODE_PAIRWISE if it % 4 == 3 else mems.MODE_MEMHASH
        want, winfo = orc.find_matches(2 if mode == mems.MODE_PAIRWISE else 0, gs, seed)
        smls = ctx.create_smls(gs, seed)
        flat, info = ctx.find_matches(smls, mode=mode, order=mems.ORDER_REFERENCE)
        tag = (block, it, hex(seed), [len(g) for g in gs])
        assert mems.flat_to_matches(flat) == want, tag
        assert info["n_hits"] == winfo["hits"] and info["collisions"] == winfo["collisions"], tag
        flat, _ = ctx.find_matches(smls, mode=mode, order=mems.ORDER_ANY)
        got = mems.flat_to_matches(flat)
        assert len(got) == len(set(got)) and sorted(got) == sorted(set(want)), tag
    ctx.close()


@pytest.mark.parametrize("block", range(3))
def test_random_repeathash(block):
    ctx, orc = gpu_context(), Oracle()
    rng = np.random.default_rng(9500 + block)
    for it in range(25):
        seed = mems.get_seed(int(rng.integers(5, 24)), int(rng.integers(0, 2)))
        n = int(rng.integers(50, 40000))
        g = synth.repeat_genome(n, seed=int(rng.integers(1, 1 << 30)), families=int(rng.integers(0, 8)), copies=int(rng.integers(2, 15)),
                                min_len=20, max_len=max(21, min(600, n // 3)), divergence=float(rng.choice([0.0, 0.02, 0.06])))
        want, winfo = orc.find_matches(1, [g], seed)
        flat, info = ctx.find_matches([ctx.create_sml(g, seed)], mode=mems.MODE_REPEAT)
        assert mems.flat_to_matches(flat) == want, (block, it, hex(seed), n)
        assert info["collisions"] == winfo["collisions"], (block, it)
    ctx.close()


def test_random_sml(ctx=None):
    ctx, orc = gpu_context(), Oracle()
    rng = np.random.default_rng(9900)
    for it in range(40):
        seed = mems.get_seed(int(rng.integers(3, 32)), int(rng.integers(0, 4)))
        if seed == 0:
            continue
        g = synth.random_genome(int(rng.integers(0, 5000)), rng)
        sml = ctx.create_sml(g, seed)
        pos, mers = sml.read()
        opos, omers = orc.sml_build(g, seed)
        assert np.array_equal(pos, opos) and np.array_equal(mers, omers), (it, hex(seed), len(g))
    ctx.close()
