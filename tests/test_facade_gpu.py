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


def test_pairwise_facade(tmp_path):
    gs = synth.genome_family(3, 20000, seed=33)
    got, _ = run_demo(tmp_path, "pairwise", 13, gs)
    want, _ = Oracle().find_matches(2, gs, mems.get_seed(13))
    assert got == want


def test_mums_files_round_trip_through_the_reference(tmp_path):
    """WriteList output is parsed by the reference's own ReadList, and the façade's ReadList parses the
    reference's WriteList output (.mums format version 3, MatchList.h:498-634)."""
    from checkers import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref not built")
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    R = Reference()
    gs = synth.genome_family(3, 20000, seed=34)
    files = []
    for i, s in enumerate(gs):
        p = tmp_path / ("seq%d.raw" % i)
        p.write_bytes(s.tobytes())
        files.append(str(p))
    r = subprocess.run([DEMO, "mums", "15"] + files, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    want, _ = Oracle().find_matches(0, gs, mems.get_seed(15))
    assert R.read_list(r.stdout) == want  # our file, their reader
    ref_text = R.write_list(gs, mems.get_seed(15))
    f = tmp_path / "ref.mums"
    f.write_text(ref_text)
    r = subprocess.run([DEMO, "readmums", str(f)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    got = []
    for line in r.stdout.splitlines():
        v = [int(x) for x in line.split("\t")]
        got.append((len(v) - 1, v[0]) + tuple(v[1:]))
    assert got == want  # their file, our reader
