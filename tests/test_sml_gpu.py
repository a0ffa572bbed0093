"""SML build on the GPU (pack -> spaced-seed extraction -> LSD radix sort) against the oracle and the
reference-generated golden vectors.  Everything goes through the C-ABI."""
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


def test_golden_kat(ctx):
    for case in json.load(open(os.path.join(GOLD, "kat_mers.json"))):
        sml = ctx.create_sml(case["seq"].encode(), case["seed"])
        assert sml.info["seed_length"] == case["L"] and sml.info["seed_weight"] == case["w"]
        assert sml.info["seed_mask"] == case["seed_mask"] and sml.info["mer_mask"] == case["mer_mask"]
        n = len(case["fwd"])
        assert sml.sml_length() == n
        fwd, dna = sml.seed_mers(np.arange(n))
        assert fwd.tolist() == case["fwd"]
        assert dna.tolist() == case["dna"]
        pos, mers = sml.read()
        assert mers.tolist() == case["sml_mer"]
        # ties are unspecified in the reference (std::sort); compare per-key position sets
        assert sorted(zip(case["sml_mer"], case["sml_pos"])) == sorted(zip(mers.tolist(), pos.tolist()))


def test_appendix_b_packed_words(ctx):
    sml = ctx.create_sml(b"ACGTTGCATGGACCTAGGATCCAATTGGCCAGTCAGTACA", mems.get_seed(5))
    assert sml.packed().tolist() == [0x1BE4E85C, 0xA350FA52, 0xD2C40000, 0, 0]


@pytest.mark.parametrize("w", [5, 9, 11, 12, 15, 16, 19, 21, 22, 27, 31])
def test_sml_matches_oracle(ctx, orc, w):
    seed = mems.get_seed(w)
    g = synth.genome_family(1, 20000 + 137 * w, seed=w)[0]
    sml = ctx.create_sml(g, seed)
    pos, mers = sml.read()
    opos, omers = orc.sml_build(g, seed)
    assert np.array_equal(mers, omers)
    assert np.array_equal(pos, opos)  # stable sort == the oracle's ascending-position tie-break
    assert np.array_equal(sml.packed(), orc.pack(g))


def test_ambiguity_codes_and_case(ctx, orc):
    seq = b"acgtNNacgtRYKMSWbdhvACGTTTGACCAGTAGGACCATTAGGACCAGTTTAGACCAGGGATTTACACACAGTTAGACC" * 3
    seed = mems.get_seed(7)
    sml = ctx.create_sml(seq, seed)
    assert np.array_equal(sml.packed(), orc.pack(seq))
    assert np.array_equal(sml.read()[1], orc.sml_build(seq, seed)[1])


def test_batch_equals_singles(ctx, orc):
    seed = mems.get_seed(15)
    gs = synth.genome_family(5, 30000, seed=3)
    gs[2] = gs[2][:1000]  # ragged
    gs.append(np.frombuffer(b"ACGT", dtype=np.uint8))  # shorter than the seed
    gs.append(np.zeros(0, dtype=np.uint8))  # empty
    smls = ctx.create_smls(gs, seed)
    for g, sml in zip(gs, smls):
        pos, mers = sml.read()
        opos, omers = orc.sml_build(g, seed)
        assert sml.info["length"] == len(g) and sml.sml_length() == len(opos)
        assert np.array_equal(mers, omers) and np.array_equal(pos, opos)
    # partial reads (MemorySML::Read semantics)
    pos, mers = smls[0].read(100, 50)
    opos, omers = orc.sml_build(gs[0], seed)
    assert np.array_equal(pos, opos[100:150]) and np.array_equal(mers, omers[100:150])
    pos, _ = smls[0].read(len(opos) - 10, 50)
    assert len(pos) == 10


def test_find_mer(ctx, orc):
    seed = mems.get_seed(11)
    g = synth.genome_family(1, 5000, seed=9)[0]
    sml = ctx.create_sml(g, seed)
    pos, mers = sml.read()
    for i in (0, 17, len(mers) // 2, len(mers) - 1):
        found, idx = sml.find_mer(int(mers[i]))
        assert found and mers[idx] == mers[i]
    found, idx = sml.find_mer(int(mers[5]) ^ (1 << 39))
    assert not found or mers[idx] == (int(mers[5]) ^ (1 << 39))


def test_gap_rejected(ctx):
    with pytest.raises(mems.MemsError) as e:
        ctx.create_sml(b"ACGTACGTAC-GTACGTACGTACGATCGATCGATCGACTAGCTAGCTAGCATCGAT", mems.get_seed(5))
    assert e.value.code == 2


def test_bad_seeds_rejected(ctx):
    with pytest.raises(mems.MemsError):
        ctx.create_sml(b"ACGT" * 20, 0)
    with pytest.raises(mems.MemsError):
        ctx.create_sml(b"ACGT" * 20, (1 << 32) - 1)  # span 32: unsupported (reference UB)


@pytest.mark.parametrize("w", [15, 19])
def test_large_sort_properties(ctx, w):
    """Full-size property checks where the oracle would be slow: sortedness, permutation, and that every
    sorted entry's key is the key at its position — for 32-bit keys (w15) and for 64-bit keys (w19: the
    onesweep_kernel<u64> / extract_kernel<u64> instances BASELINE configs 3 and 5 run on)."""
    seed = mems.get_seed(w)
    L = mems.get_seed_length(seed)
    gs = synth.genome_family(3, 3_000_000, seed=5)
    smls = ctx.create_smls(gs, seed)
    for g, sml in zip(gs, smls):
        pos, mers = sml.read()
        n = sml.sml_length()
        assert n == len(g) - L + 1
        assert np.all(mers[1:] >= mers[:-1])
        assert np.array_equal(np.sort(pos), np.arange(n, dtype=np.uint32))
        eq = mers[1:] == mers[:-1]
        assert np.all(pos[1:][eq] > pos[:-1][eq])  # stable: ties ascend by position
        sample = np.random.default_rng(1).integers(0, n, size=4096)
        _, dna = sml.seed_mers(pos[sample])
        assert np.array_equal(dna, mers[sample])


def test_seed_occurrence(ctx, orc):
    """SeedOccurrenceList::construct (per-position seed multiplicity, smoothed) — bit-exact floats."""
    seed = mems.get_seed(11)
    g = synth.repeat_genome(20000, seed=5, families=4, copies=6, min_len=50, max_len=300)
    other = synth.genome_family(1, 7000, seed=6)[0]
    smls = ctx.create_smls([g, other, np.frombuffer(b"ACGTACGTAC", dtype=np.uint8)], seed)
    for s, q in zip(smls, [g, other, b"ACGTACGTAC"]):
        assert np.array_equal(s.seed_occurrence(), orc.seed_occurrence(q, seed))
    single = ctx.create_sml(g, mems.get_seed(15))
    assert np.array_equal(single.seed_occurrence(), orc.seed_occurrence(g, mems.get_seed(15)))


def test_more_than_256_sequences_rejected(ctx):
    """A create call sorts the per-sequence lists with one 8-bit counting pass on the sequence tag: more than 256
    sequences must be refused cleanly (MEMS_ERR_UNSUPPORTED), not overflow that pass."""
    seqs = [synth.random_genome(60, np.random.default_rng(i)) for i in range(257)]
    with pytest.raises(mems.MemsError) as e:
        ctx.create_smls(seqs, mems.get_seed(7))
    assert e.value.code == 4
    smls = ctx.create_smls(seqs[:256], mems.get_seed(7))  # 256 is fine, accessors included
    pos, mers = smls[255].read()
    assert len(pos) == 60 - mems.get_seed_length(mems.get_seed(7)) + 1 and np.all(mers[1:] >= mers[:-1])
