"""MemHash / RepeatHash match finding on the GPU against the oracle and the reference's golden MatchLists.
All calls go through the C-ABI."""
import json
import os

import numpy as np
import pytest

import libmems_b200 as mems
from checkers import Oracle
from gpu_util import gpu_context
from libmems_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ctx():
    c = gpu_context()
    yield c
    c.close()


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def gpu_matches(ctx, seqs, seed, mode, order=mems.ORDER_CANONICAL):
    smls = ctx.create_smls(seqs, seed)
    flat, info = ctx.find_matches(smls, mode=mode, order=order)
    return mems.flat_to_matches(flat), info


def canonical(matches):
    """tuples (SeqCount, Length, Start...) sorted lexicographically, compared as sets."""
    return sorted(set(matches))


def test_golden_matchlists(ctx):
    for case in json.load(open(os.path.join(GOLD, "matchlists.json"))):
        seqs = [s.encode() for s in case["seqs"]]
        want = [tuple(m) for m in case["matches"]]
        got, info = gpu_matches(ctx, seqs, case["seed"], case["mode"], mems.ORDER_REFERENCE)
        assert got == want, case["tag"]  # same matches in the reference's own output order
        assert info["mem_count"] == case["mem_count"] and info["collisions"] == case["collisions"], case["tag"]
        if case["mode"] == mems.MODE_MEMHASH:
            got, _ = gpu_matches(ctx, seqs, case["seed"], case["mode"], mems.ORDER_CANONICAL)
            assert got == canonical(want), case["tag"]


@pytest.mark.parametrize("it", range(16))
def test_memhash_vs_oracle(ctx, orc, it):
    rng = np.random.default_rng(5000 + it)
    w = int(rng.integers(5, 25))
    seed = mems.get_seed(w, int(rng.integers(0, 3)))
    G = int(rng.integers(2, 9))
    n = int(rng.integers(200, 60000))
    gs = synth.genome_family(G, n, seed=700 + it, snp_rate=float(rng.choice([0.0, 0.01, 0.05])),
                             n_indels=int(rng.integers(0, 8)), max_indel=30)
    want, winfo = orc.find_matches(0, gs, seed)
    got, info = gpu_matches(ctx, gs, seed, mems.MODE_MEMHASH)
    assert got == canonical(want)
    assert info["n_hits"] == winfo["hits"]
    got_ref, info_ref = gpu_matches(ctx, gs, seed, mems.MODE_MEMHASH, mems.ORDER_REFERENCE)
    assert got_ref == want
    assert info_ref["collisions"] == winfo["collisions"] and info_ref["mem_count"] == winfo["mem_count"]


@pytest.mark.parametrize("it", range(10))
def test_repeathash_vs_oracle(ctx, orc, it):
    rng = np.random.default_rng(6000 + it)
    seed = mems.get_seed(int(rng.integers(7, 25)), int(rng.integers(0, 2)))
    g = synth.repeat_genome(int(rng.integers(5000, 80000)), seed=40 + it, families=int(rng.integers(1, 10)),
                            copies=int(rng.integers(2, 12)), min_len=50, max_len=800, divergence=0.03)
    want, winfo = orc.find_matches(1, [g], seed)
    got, info = gpu_matches(ctx, [g], seed, mems.MODE_REPEAT)
    assert got == want  # RepeatHash is always replayed in reference order (its drops are order dependent)
    assert info["collisions"] == winfo["collisions"]


def test_singles_then_find_equals_batch(ctx, orc):
    """DNAMemorySML::Create one by one, then FindMatches (the reference's calling pattern)."""
    seed = mems.get_seed(15)
    gs = synth.genome_family(4, 20000, seed=77)
    smls = [ctx.create_sml(g, seed) for g in gs]
    flat, _ = ctx.find_matches(smls, mode=mems.MODE_MEMHASH)
    want, _ = orc.find_matches(0, gs, seed)
    assert sorted(mems.flat_to_matches(flat)) == canonical(want)  # ORDER_ANY: distinct, device order
    # a subset, in another order, of a batch
    b = ctx.create_smls(gs, seed)
    flat, _ = ctx.find_matches([b[2], b[0]], mode=mems.MODE_MEMHASH, order=mems.ORDER_CANONICAL)
    want, _ = orc.find_matches(0, [gs[2], gs[0]], seed)
    assert mems.flat_to_matches(flat) == canonical(want)


def test_edge_cases(ctx, orc):
    seed = mems.get_seed(15)
    g = synth.genome_family(1, 3000, seed=8)[0]
    # identical sequences: one match spanning everything
    got, _ = gpu_matches(ctx, [g, g], seed, mems.MODE_MEMHASH)
    assert got == [(2, 3000, 1, 1)]
    # a sequence and its reverse complement
    got, _ = gpu_matches(ctx, [g, synth.revcomp(g)], seed, mems.MODE_MEMHASH)
    assert got == canonical(orc.find_matches(0, [g, synth.revcomp(g)], seed)[0])
    # nothing in common / too short / empty
    h = synth.genome_family(1, 3000, seed=9)[0]
    assert gpu_matches(ctx, [g, h], seed, mems.MODE_MEMHASH)[0] == canonical(orc.find_matches(0, [g, h], seed)[0])
    assert gpu_matches(ctx, [g, b"ACGT", b""], seed, mems.MODE_MEMHASH)[0] == []
    # different seeds are rejected like MatchFinder.cpp:190-199
    a = ctx.create_sml(g, mems.get_seed(15))
    b = ctx.create_sml(g, mems.get_seed(13))
    with pytest.raises(mems.MemsError) as e:
        ctx.find_matches([a, b])
    assert e.value.code == 5


def test_medium_size_vs_oracle(ctx, orc):
    seed = mems.get_seed(15)
    gs = synth.genome_family(4, 400_000, seed=123)
    want, winfo = orc.find_matches(0, gs, seed)
    got, info = gpu_matches(ctx, gs, seed, mems.MODE_MEMHASH)
    assert got == canonical(want)
    assert info["n_hits"] == winfo["hits"]


@pytest.mark.parametrize("w", [15, 19, 16])
def test_long_walks_between_sparse_hits(ctx, orc, w):
    """Few hits, long matching diagonal: the extension walks exceed a warp's probe budget and are finished
    by the CTA-wide walker, including the case where a long walk has to link two distant hit groups.  w19 runs the
    64-bit-key instances of every kernel, w16 the even-weight rule (a mer equal to its own reverse complement never
    matches across orientations)."""
    seed = mems.get_seed(w)
    rng = np.random.default_rng(99)
    T = synth.random_genome(30_000, rng)
    M1, M2 = synth.random_genome(60, rng), synth.random_genome(60, rng)
    # every window inside a copy of T is repeated (no hit); only windows touching M1/M2 are unique
    for X in (np.concatenate([T, M1, T]), np.concatenate([T, M1, T, M2, T])):
        for other in (X, synth.revcomp(X)):
            want, winfo = orc.find_matches(0, [X, other], seed)
            got, info = gpu_matches(ctx, [X, other], seed, mems.MODE_MEMHASH)
            assert got == canonical(want)
            assert info["n_hits"] == winfo["hits"]
            assert any(m[1] == len(X) for m in got)  # the full-length diagonal is found
            got_ref, _ = gpu_matches(ctx, [X, other], seed, mems.MODE_MEMHASH, mems.ORDER_REFERENCE)
            assert got_ref == want


@pytest.mark.parametrize("it", range(6))
def test_pairwise_vs_oracle(ctx, orc, it):
    """PairwiseMatchFinder policy: every pair of sequences that hold a seed exactly once."""
    rng = np.random.default_rng(7000 + it)
    seed = mems.get_seed(int(rng.integers(7, 22)))
    gs = synth.genome_family(int(rng.integers(2, 7)), int(rng.integers(500, 40000)), seed=900 + it,
                             n_indels=3, max_indel=20)
    if it % 2:
        gs.append(np.concatenate([gs[1], gs[1]]))  # holds every seed twice: excluded, must not block the other pairs
    want, winfo = orc.find_matches(2, gs, seed)
    got, info = gpu_matches(ctx, gs, seed, mems.MODE_PAIRWISE)
    assert got == canonical(want)
    assert info["n_hits"] == winfo["hits"]
    got_ref, info_ref = gpu_matches(ctx, gs, seed, mems.MODE_PAIRWISE, mems.ORDER_REFERENCE)
    assert got_ref == want
    assert info_ref["collisions"] == winfo["collisions"]


def test_masked_memhash_vs_oracle(ctx, orc):
    """MaskedMemHash: only hits whose sequence set equals the mask (sequence 0 = most significant bit)."""
    gs = synth.genome_family(4, 30000, seed=41, snp_rate=0.03, n_indels=3, max_indel=20)
    seed = mems.get_seed(11)
    smls = ctx.create_smls(gs, seed)
    for mask in (0b1111, 0b1100, 0b0101, 0b1011):
        want, winfo = orc.find_matches_masked(gs, seed, mask)
        flat, info = ctx.find_matches(smls, order=mems.ORDER_REFERENCE, seq_mask=mask)
        assert mems.flat_to_matches(flat) == want, mask
        assert info["collisions"] == winfo["collisions"] and info["n_hits"] == winfo["hits"]


def test_diagonal_hash_collisions_are_harmless(orc):
    """Force many diagonals into few hash buckets (mems_test_hooks): segments get split by foreign entries,
    some components are found twice, and the de-dup of the marked components must restore the exact MatchList."""
    seed = mems.get_seed(11)
    gs = synth.genome_family(5, 40000, seed=51, snp_rate=0.03, n_indels=6, max_indel=30)
    want, winfo = orc.find_matches(0, gs, seed)
    for bits in (1, 3, 6):
        c = gpu_context()
        c.set_test_hooks(hash_bits=bits)
        smls = c.create_smls(gs, seed)
        flat, info = c.find_matches(smls)  # ORDER_ANY: device order, distinct
        got = mems.flat_to_matches(flat)
        assert len(got) == len(set(got)), bits
        assert sorted(got) == canonical(want), bits
        flat, info = c.find_matches(smls, order=mems.ORDER_REFERENCE)
        assert mems.flat_to_matches(flat) == want, bits
        c.close()


@pytest.mark.parametrize("w", range(11, 22))
def test_seed_weight_sweep(ctx, orc, w):
    """BASELINE config 4 at test size: the requested weights 11..21 map to getSeed(w, 0) — including the
    "weight 11" row that really is a weight-12 pattern — over a dozen related genomes."""
    seed = mems.get_seed(w)
    gs = synth.genome_family(12, 12000, seed=60 + w, n_indels=4, max_indel=25)
    want, winfo = orc.find_matches(0, gs, seed)
    got, info = gpu_matches(ctx, gs, seed, mems.MODE_MEMHASH)
    assert got == canonical(want)
    assert info["n_hits"] == winfo["hits"]


@pytest.mark.parametrize("it", range(4))
def test_multi_seed_accumulation(ctx, orc, it):
    """Three seed ranks accumulated in one table (ClearSequences + FindMatches per pattern, ProgressiveAligner.cpp:619-653)."""
    gs = synth.genome_family(2 + it % 2, 20000 + 3000 * it, seed=70 + it, snp_rate=0.04, n_indels=5, max_indel=20)
    w = 11 + 2 * it
    seeds = [mems.get_seed(w, r) for r in range(3)]
    want, winfo = orc.find_matches_multi_seed(gs, seeds)
    table = mems.HashTable()
    for sd in seeds:
        smls = ctx.create_smls(gs, sd)
        flat, info = ctx.find_matches(smls, table=table)
    assert mems.flat_to_matches(flat) == want
    assert info["mem_count"] == winfo["mem_count"] and info["collisions"] == winfo["collisions"]
    table.clear()
    flat, info = ctx.find_matches(ctx.create_smls(gs, seeds[0]), table=table)
    assert mems.flat_to_matches(flat) == orc.find_matches(0, gs, seeds[0])[0]


@pytest.mark.parametrize("w", [15, 19])
def test_grid_wide_walks(orc, w):
    """With the warp and CTA budgets shrunk to 1, the long diagonals of the sparse-hit inputs are finished by the
    grid-cooperative walker (all CTAs on one walk, grid barrier per round), including the linking case; w19 = the
    64-bit-key instances (giant_walk_kernel<u64>, long_walk_*<u64>)."""
    c = gpu_context()
    c.set_test_hooks(walk_budget=1)
    seed = mems.get_seed(w)
    rng = np.random.default_rng(98)
    T = synth.random_genome(40_000, rng)
    M1, M2 = synth.random_genome(60, rng), synth.random_genome(60, rng)
    for X in (np.concatenate([T, M1, T]), np.concatenate([T, M1, T, M2, T])):
        for other in (X, synth.revcomp(X)):
            want, _ = orc.find_matches(0, [X, other], seed)
            smls = c.create_smls([X, other], seed)
            flat, _ = c.find_matches(smls, order=mems.ORDER_REFERENCE)
            assert mems.flat_to_matches(flat) == want
    gs = synth.genome_family(4, 150_000, seed=77)
    want, _ = orc.find_matches(0, gs, seed)
    flat, _ = c.find_matches(c.create_smls(gs, seed), order=mems.ORDER_CANONICAL)
    assert mems.flat_to_matches(flat) == canonical(want)
    c.close()


def test_many_sequences(ctx, orc):
    """40 related sequences in one call (BASELINE config 4 has 50): 6 tag bits for the sequence id."""
    seed = mems.get_seed(13)
    gs = synth.genome_family(40, 5000, seed=91, n_indels=2, max_indel=15)
    want, winfo = orc.find_matches(0, gs, seed)
    got, info = gpu_matches(ctx, gs, seed, mems.MODE_MEMHASH)
    assert got == canonical(want)
    assert info["n_hits"] == winfo["hits"]
    with pytest.raises(mems.MemsError) as e:
        ctx.find_matches(ctx.create_smls(synth.genome_family(65, 300, seed=92, n_indels=0), seed))
    assert e.value.code == 4  # more than MEMS_MAX_SEQS


def test_repeathash_megabase_vs_oracle(ctx, orc):
    """RepeatHash on a 2 Mbp sequence with planted repeat families (BASELINE config 3 at a size the oracle
    finishes in seconds), default seed weight for that length."""
    g = synth.repeat_genome(2_000_000, seed=93, families=40, copies=20)
    seed = mems.get_seed(mems.get_default_seed_weight(len(g)))
    want, winfo = orc.find_matches(1, [g], seed)
    got, info = gpu_matches(ctx, [g], seed, mems.MODE_REPEAT)
    assert got == want
    assert info["collisions"] == winfo["collisions"] and info["max_run"] <= 1000


def _window_matches(smls, lens, L, seed_mask, match, k):
    """The reference's window test (MatchFinder.h:264-308) for window k of `match`, evaluated with the mers the
    library itself reports at those positions."""
    _, length, *starts = match
    ref = None
    for g, st in enumerate(starts):
        if st == 0:
            continue
        p = st - 1 + k if st > 0 else -st - 1 + (length - L - k)
        if p < 0 or p > lens[g] - L:
            return False
        mer = int(smls[g].seed_mers([p])[1][0])
        strand = mer & 1
        tagged = (mer & seed_mask, (strand == 0) if st > 0 else (strand == 1))
        if ref is None:
            ref = tagged
        elif tagged != ref:
            return False
    return True


def test_full_size_match_properties(ctx):
    """BASELINE config 2 at full size (8 x 5 Mbp, w15), where the oracle would need minutes: every reported match
    must start and end on a matching seed window and be maximal (no matching window within L beyond either end),
    matches must be distinct, and the run must be reproducible."""
    seed = mems.get_seed(15)
    gs = synth.genome_family(8, 5_000_000, seed=2)
    smls = ctx.create_smls(gs, seed)
    flat, info = ctx.find_matches(smls)
    matches = mems.flat_to_matches(flat)
    assert len(matches) == info["n_matches"] == len(set(matches))
    flat2, _ = ctx.find_matches(smls)
    assert np.array_equal(flat, flat2)
    L, mask, lens = smls[0].info["seed_length"], smls[0].info["seed_mask"], [len(g) for g in gs]
    rng = np.random.default_rng(3)
    picks = list(rng.choice(len(matches), size=40, replace=False)) + [int(np.argmax([m[1] for m in matches]))]
    for i in picks:
        m = matches[i]
        assert sum(1 for s in m[2:] if s) >= 2 and m[1] >= L
        assert _window_matches(smls, lens, L, mask, m, 0)
        assert _window_matches(smls, lens, L, mask, m, m[1] - L)
        for d in range(1, L + 1):
            assert not _window_matches(smls, lens, L, mask, m, -d)
            assert not _window_matches(smls, lens, L, mask, m, m[1] - L + d)


def test_contexts_are_independent_across_threads(orc):
    """One context per host thread, used concurrently (the reference keeps one MemHash per OpenMP thread,
    Aligner.h:198 / ProgressiveAligner.cpp:695): results must equal the oracle's for every thread."""
    import threading
    seed = mems.get_seed(13)
    jobs = [synth.genome_family(3, 20000 + 3000 * t, seed=200 + t, n_indels=4, max_indel=20) for t in range(4)]
    want = [canonical(orc.find_matches(0, gs, seed)[0]) for gs in jobs]
    got, errors = [None] * len(jobs), []

    def work(t):
        try:
            c = gpu_context()
            for _ in range(5):
                flat, _ = c.find_matches(c.create_smls(jobs[t], seed), order=mems.ORDER_CANONICAL)
                got[t] = mems.flat_to_matches(flat)
                assert got[t] == want[t]
            c.close()
        except Exception as e:  # surfaced below: a failing assert in a thread would otherwise be lost
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(len(jobs))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    assert got == want


def test_u64_medium_size_vs_oracle(ctx, orc):
    """64-bit keys (w19, the seed of BASELINE configs 3 and 5) on several 300 kbp genomes with inversions: every kernel's
    <u64> instance against the oracle, in the reference's order too."""
    seed = mems.get_seed(19)
    gs = synth.genome_family(5, 300_000, seed=321)
    want, winfo = orc.find_matches(0, gs, seed)
    got, info = gpu_matches(ctx, gs, seed, mems.MODE_MEMHASH)
    assert got == canonical(want)
    assert info["n_hits"] == winfo["hits"]
    got_ref, info_ref = gpu_matches(ctx, gs, seed, mems.MODE_MEMHASH, mems.ORDER_REFERENCE)
    assert got_ref == want and info_ref["collisions"] == winfo["collisions"]


def test_non_palindromic_patterns(ctx, orc):
    """The two entries of the reference's seed table that are not palindromes (weight 19 rank 2, weight 21 rank 1,
    SURVEY.md 0-7) and a hand-made asymmetric pattern: cared base i sits at another offset on the other strand, so
    reverse members are compared offset by offset.  Inputs carry an inversion, i.e. reverse members."""
    for seed in (mems.get_seed(19, 2), mems.get_seed(21, 1), 0b1101000111):
        gs = synth.genome_family(4, 40_000, seed=77, snp_rate=0.02, n_indels=4, max_indel=20)
        gs.append(synth.revcomp(gs[1]))
        want, winfo = orc.find_matches(0, gs, seed)
        got, info = gpu_matches(ctx, gs, seed, mems.MODE_MEMHASH)
        assert got == canonical(want), hex(seed)
        assert info["n_hits"] == winfo["hits"]


def test_many_small_problems_in_one_call(ctx, orc):
    """mems_find_matches_many: dozens of independent gap-sized problems (pairs and small sets, ragged, some empty of
    matches) in one launch set; every problem's list must equal the oracle's for that problem alone."""
    rng = np.random.default_rng(17)
    for seed, mode in ((mems.get_seed(9), mems.MODE_MEMHASH), (mems.get_seed(11), mems.MODE_MEMHASH), (mems.get_seed(9), mems.MODE_PAIRWISE)):
        problems = []
        for k in range(70):
            G = int(rng.integers(2, 5)) if k % 7 else 2
            n = int(rng.integers(30, 12000))
            gs = synth.genome_family(G, n, seed=1000 + k, snp_rate=0.03, n_indels=int(rng.integers(0, 4)), max_indel=15)
            if k % 11 == 0:
                gs[1] = synth.random_genome(n, rng)  # unrelated: no matches
            if k % 13 == 0:
                gs[-1] = gs[-1][:10]  # shorter than the seed
            problems.append(gs)
        res = ctx.find_matches_many(problems, seed, mode=mode, order=mems.ORDER_CANONICAL)
        assert len(res) == len(problems)
        for k, (gs, (flat, info)) in enumerate(zip(problems, res)):
            want, _ = orc.find_matches(2 if mode == mems.MODE_PAIRWISE else 0, gs, seed)
            assert mems.flat_to_matches(flat) == canonical(want), k
            assert info["seq_count"] == len(gs) and info["n_matches"] == len(canonical(want))
    # limits: more than MEMS_MAX_SEQS sequences in one problem, more than 256 in all
    with pytest.raises(mems.MemsError):
        ctx.find_matches_many([[b"ACGT" * 20] * 65], mems.get_seed(7))
    with pytest.raises(mems.MemsError):
        ctx.find_matches_many([[b"ACGT" * 20] * 2] * 129, mems.get_seed(7))


def test_asynchronous_delivery(ctx, orc):
    """MEMS_ORDER_ANY records leave the device behind the call's last kernel (mems_b200.h, mems_matches_wait): calls
    issued back to back before any of their MatchLists is read — the device buffers of one call are reused by the next
    while its copy may still be in flight — must each deliver exactly the oracle's set, in any order of reading."""
    seed = mems.get_seed(13)
    jobs = [synth.genome_family(3 + (t % 3), 30000 + 7000 * t, seed=300 + t, n_indels=4, max_indel=20) for t in range(6)]
    want = [canonical(orc.find_matches(0, gs, seed)[0]) for gs in jobs]
    pending = []
    for rnd in range(2):
        for gs in jobs:
            smls = ctx.create_smls(gs, seed)
            pend, info = ctx.find_matches(smls, wait=False)
            for s in smls:
                s.close()
            pending.append((pend, info))
    # read them newest first, half of them after an explicit wait
    for k in reversed(range(len(pending))):
        pend, info = pending[k]
        if k % 2:
            pend.wait()
        got = mems.flat_to_matches(pend.records())
        assert len(got) == info["n_matches"]
        assert canonical(got) == want[k % len(jobs)]
    # a MatchList that is never read is simply given back
    smls = ctx.create_smls(jobs[0], seed)
    pend, info = ctx.find_matches(smls, wait=False)
    del pend
    flat, _ = ctx.find_matches(smls)
    assert canonical(mems.flat_to_matches(flat)) == want[0]


def test_context_trim_gives_idle_memory_back(orc):
    """mems_ctx_trim: the arena's slabs without a live object go back to the driver; what is still alive (an SML) keeps
    its slab and stays usable, and the context works as before afterwards."""
    c = gpu_context()
    seed = mems.get_seed(13)
    gs = synth.genome_family(3, 40000, seed=77, n_indels=4, max_indel=20)
    want = canonical(orc.find_matches(0, gs, seed)[0])
    smls = c.create_smls(gs, seed)
    flat, _ = c.find_matches(smls, order=mems.ORDER_CANONICAL)
    assert mems.flat_to_matches(flat) == want
    held = c.trim()
    assert held > 0  # the SMLs are alive
    flat, _ = c.find_matches(smls, order=mems.ORDER_CANONICAL)
    assert mems.flat_to_matches(flat) == want
    for s in smls:
        s.close()
    del flat
    assert c.trim() == 0
    flat, _ = c.find_matches(c.create_smls(gs, seed), order=mems.ORDER_CANONICAL)
    assert mems.flat_to_matches(flat) == want
    c.close()
