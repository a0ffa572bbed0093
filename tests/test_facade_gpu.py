"""The C++ façade (libmems_b200/host/libMems/*.h: MatchList, DNAMemorySML, MemHash, RepeatHash with the
reference's member names) driven like the reference's callers, checked against the oracle."""
import os
import subprocess

import pytest

import libmems_b200 as mems
from checkers import Oracle
from libmems_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "build", "facade_demo")


def run_demo(tmp_path, mode, weight, seqs):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(DEMO):
        import __graft_entry__
        __graft_entry__.build()
    files = []
    for i, s in enumerate(seqs):
        p = tmp_path / ("seq%d.raw" % i)
        p.write_bytes(s.tobytes())
        files.append(str(p))
    r = subprocess.run([DEMO, mode, str(weight)] + files, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = []
    for line in r.stdout.splitlines():
        v = [int(x) for x in line.split("\t")]
        out.append((len(v) - 1, v[0]) + tuple(v[1:]))
    return out, r.stderr


def test_memhash_facade(tmp_path):
    gs = synth.genome_family(3, 30000, seed=31)
    got, log = run_demo(tmp_path, "memhash", 15, gs)
    want, info = Oracle().find_matches(0, gs, mems.get_seed(15))
    assert got == want  # reference output order
    assert "MemCount %d MemCollisionCount %d" % (info["mem_count"], info["collisions"]) in log


def test_repeathash_facade(tmp_path):
    g = synth.repeat_genome(30000, seed=32, families=4, copies=5, min_len=60, max_len=400)
    got, _ = run_demo(tmp_path, "repeat", 13, [g])
    want, _ = Oracle().find_matches(1, [g], mems.get_seed(13))
    assert got == want
