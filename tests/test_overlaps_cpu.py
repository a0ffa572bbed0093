"""EliminateOverlaps (Aligner.cpp:62-180), the step the callers run right behind FindMatches: the façade's host
implementation (libmems_b200/host/libMems/Aligner.h) against what the UNMODIFIED reference function left of the same
lists (tests/golden/overlaps.json; oracle/Makefile cuts the function out of Aligner.cpp and compiles it as it is) —
the same matches in the same order.  Host code only: runs without a GPU."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "build", "facade_demo")
GOLD = os.path.join(ROOT, "tests", "golden")


def run_overlaps(tmp_path, matches):
    if not os.path.exists(DEMO):
        import __graft_entry__
        __graft_entry__.build()
    f = tmp_path / "in.txt"
    f.write_text("".join("\t".join(str(x) for x in m[1:]) + "\n" for m in matches))
    r = subprocess.run([DEMO, "overlaps", str(f)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = []
    for line in r.stdout.splitlines():
        v = [int(x) for x in line.split("\t")]
        out.append((len(v) - 1, v[0]) + tuple(v[1:]))
    return out


def test_eliminate_overlaps_equals_reference_fixture(tmp_path):
    for case in json.load(open(os.path.join(GOLD, "overlaps.json"))):
        got = run_overlaps(tmp_path, [tuple(m) for m in case["input"]])
        assert got == [tuple(m) for m in case["output"]], case["tag"]


@pytest.mark.ref
def test_eliminate_overlaps_equals_reference_live(tmp_path):
    from checkers import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref not built")
    from libmems_b200 import synth
    R = Reference()
    for k in range(6):
        gs = synth.genome_family(2 + k % 4, 6000 + 3000 * k, seed=300 + k, snp_rate=0.02 + 0.01 * (k % 3))
        matches, _ = R.find_matches(0, gs, R.get_seed(9 + 2 * (k % 3)))
        assert run_overlaps(tmp_path, matches) == R.eliminate_overlaps(matches), k
    assert run_overlaps(tmp_path, []) == [] and run_overlaps(tmp_path, [(2, 30, 5, 9)]) == [(2, 30, 5, 9)]
