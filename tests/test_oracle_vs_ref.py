"""Randomised comparison of the C restatement with the unmodified reference (oracle/_ref).
Skipped where the reference library was not built (it needs /root/reference at build time).
CPU only."""
import numpy as np
import pytest

from checkers import Oracle, Reference
from libmems_b200 import synth

pytestmark = pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def libs():
    return Oracle(), Reference()


def test_sml_keys_match(libs):
    O, R = libs
    for w in (5, 9, 11, 12, 15, 16, 19, 21, 22, 27, 31):
        seed = O.get_seed(w)
        g = synth.genome_family(1, 6000, seed=w)[0]
        po, mo = O.sml_build(g, seed)
        pr, mr = R.sml_build(g, seed)
        assert (mo == mr).all()
        assert sorted(zip(mo.tolist(), po.tolist())) == sorted(zip(mr.tolist(), pr.tolist()))


@pytest.mark.parametrize("it", range(12))
def test_memhash_order_exact(libs, it):
    O, R = libs
    rng = np.random.default_rng(1000 + it)
    w = int(rng.integers(5, 25))
    seed = O.get_seed(w, int(rng.integers(0, 3)))
    G = int(rng.integers(2, 7))
    n = int(rng.integers(200, 20000))
    gs = synth.genome_family(G, n, seed=100 + it, snp_rate=float(rng.choice([0.0, 0.01, 0.05])),
                             n_indels=int(rng.integers(0, 8)), max_indel=30)
    mo, io = O.find_matches(0, gs, seed)
    mr, ir = R.find_matches(0, gs, seed)
    assert mo == mr
    assert io["collisions"] == ir["collisions"] and io["mem_count"] == ir["mem_count"]


@pytest.mark.parametrize("it", range(8))
def test_repeathash_order_exact(libs, it):
    O, R = libs
    rng = np.random.default_rng(2000 + it)
    seed = O.get_seed(int(rng.integers(7, 25)), int(rng.integers(0, 2)))
    g = synth.repeat_genome(int(rng.integers(5000, 40000)), seed=it, families=int(rng.integers(1, 8)),
                            copies=int(rng.integers(2, 9)), min_len=50, max_len=600, divergence=0.03)
    mo, io = O.find_matches(1, [g], seed)
    mr, ir = R.find_matches(1, [g], seed)
    assert mo == mr
    assert io["collisions"] == ir["collisions"]


@pytest.mark.parametrize("it", range(4))
def test_pairwise_order_exact(libs, it):
    O, R = libs
    rng = np.random.default_rng(3000 + it)
    seed = O.get_seed(int(rng.integers(7, 22)))
    gs = synth.genome_family(int(rng.integers(2, 5)), int(rng.integers(500, 10000)), seed=300 + it,
                             n_indels=3, max_indel=20)
    mo, _ = O.find_matches(2, gs, seed)
    mr, _ = R.find_matches(2, gs, seed)
    assert mo == mr


def test_seed_occurrence(libs):
    O, R = libs
    g = synth.repeat_genome(5000, seed=5, families=3, copies=4, min_len=50, max_len=200)
    seed = O.get_seed(11)
    assert np.array_equal(O.seed_occurrence(g, seed), R.seed_occurrence(g, seed))


def test_masked_memhash(libs):
    O, R = libs
    gs = synth.genome_family(4, 6000, seed=41, snp_rate=0.03, n_indels=3, max_indel=20)
    seed = O.get_seed(11)
    for mask in (0, 0b1111, 0b1100, 0b0101, 0b1011):
        mo, io = O.find_matches_masked(gs, seed, mask)
        mr, ir = R.find_matches_masked(gs, seed, mask)
        assert mo == mr, mask
        assert io["collisions"] == ir["collisions"]


@pytest.mark.parametrize("it", range(4))
def test_multi_seed_accumulation(libs, it):
    """Three seed ranks accumulated in one MemHash table (ProgressiveAligner.cpp:619-653)."""
    O, R = libs
    gs = synth.genome_family(2 + it % 2, 8000 + 1000 * it, seed=70 + it, snp_rate=0.04, n_indels=5, max_indel=20)
    w = 11 + 2 * it
    seeds = [O.get_seed(w, r) for r in range(3)]
    mo, io = O.find_matches_multi_seed(gs, seeds)
    mr, ir = R.find_matches_multi_seed(gs, seeds)
    assert mo == mr
    assert io["mem_count"] == ir["mem_count"] and io["collisions"] == ir["collisions"]
