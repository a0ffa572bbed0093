"""Deterministic synthetic genomes for the parity tests and the benchmark.

Base genome: iid uniform ACGT.  Derived genome g>0: the base with a per-base SNP rate, a number of
short indels (random inserted bases, never homopolymers), and one inversion (reverse complement)
of 10 % of the length starting at N/3.  Repeat genomes (RepeatHash config): planted families of
diverged copies, some inverted.  Everything is driven by numpy's PCG64 with explicit seeds, so
tests, bench and the CPU baseline all see byte-identical inputs.  No seed in these inputs exceeds
libMems' MER_REPEAT_LIMIT of 1000 occurrences (MatchFinder.cpp:166).
"""
import numpy as np

_ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[ord("A")], _COMP[ord("C")], _COMP[ord("G")], _COMP[ord("T")] = ord("T"), ord("G"), ord("C"), ord("A")


def random_genome(n, rng):
    return _ALPHA[rng.integers(0, 4, size=n, dtype=np.uint8)]


def revcomp(a):
    return _COMP[a[::-1]]


def mutate(base, rng, snp_rate=0.01, n_indels=20, max_indel=50, inversion=True):
    g = base.copy()
    n = len(g)
    # SNPs: uniform replacement (may re-draw the same base, like the survey's generator)
    hits = np.nonzero(rng.random(n) < snp_rate)[0]
    g[hits] = _ALPHA[rng.integers(0, 4, size=len(hits), dtype=np.uint8)]
    for _ in range(n_indels):
        p = int(rng.integers(0, len(g)))
        k = int(rng.integers(1, max_indel + 1))
        if rng.integers(0, 2):
            g = np.concatenate([g[:p], g[p + k:]])
        else:
            g = np.concatenate([g[:p], random_genome(k, rng), g[p:]])
    if inversion and len(g) >= 30:
        a = len(g) // 3
        b = a + len(g) // 10
        g = np.concatenate([g[:a], revcomp(g[a:b]), g[b:]])
    return np.ascontiguousarray(g)


def genome_family(n_genomes, n, seed=1, snp_rate=0.01, n_indels=20, max_indel=50):
    """n_genomes related sequences of about n bases as a list of uint8 arrays (ASCII)."""
    rng = np.random.default_rng(seed)
    base = random_genome(n, rng)
    out = [base]
    for _ in range(1, n_genomes):
        out.append(mutate(base, rng, snp_rate, n_indels, max_indel))
    return out


def repeat_genome(n, seed=1, families=200, copies=20, min_len=300, max_len=2300, divergence=0.02,
                  inverted_frac=0.4):
    """One sequence of n bases with planted repeat families (RepeatHash workload)."""
    rng = np.random.default_rng(seed)
    g = random_genome(n, rng)
    for _ in range(families):
        ln = int(rng.integers(min_len, max_len + 1))
        if ln * 2 >= n:
            continue
        unit = random_genome(ln, rng)
        for _ in range(copies):
            c = unit.copy()
            hits = np.nonzero(rng.random(ln) < divergence)[0]
            c[hits] = _ALPHA[rng.integers(0, 4, size=len(hits), dtype=np.uint8)]
            if rng.random() < inverted_frac:
                c = revcomp(c)
            p = int(rng.integers(0, n - ln))
            g[p:p + ln] = c
    return np.ascontiguousarray(g)


# BASELINE.json configs at their stated sizes: (genomes, length, seed weight, mode, generator seed).  bench.py, the
# full-size parity tests and tools/gen_golden_full.py all take their inputs from here, so the digests the
# reference produced in the build container (tests/golden/full_*.json) describe exactly what the GPU is given.
BASELINE_WORKLOADS = {
    "c1": (2, 5_000_000, 15, "memhash", 2),
    "c2": (8, 5_000_000, 15, "memhash", 2),
    "c3": (1, 100_000_000, 19, "repeat", 2),
}


def baseline_genomes(name, n_genomes=None, length=None, seed=None):
    """The synthetic sequences of a BASELINE config (optionally at another size: bounded CPU samples)."""
    g, n, _, mode, s = BASELINE_WORKLOADS[name]
    g = g if n_genomes is None else n_genomes
    n = n if length is None else length
    s = s if seed is None else seed
    if mode == "repeat":
        return [repeat_genome(n, seed=s, families=max(2, int(200 * n / 100_000_000)), copies=20)]
    return genome_family(g, n, seed=s)


def matchlist_digest(matches):
    """SHA-256 over the int64 little-endian stream [SeqCount, Length, Start(0..)] of the matches in the given order."""
    import hashlib
    h = hashlib.sha256()
    buf = np.fromiter((x for m in matches for x in m), dtype="<i8")
    h.update(buf.tobytes())
    return h.hexdigest()
