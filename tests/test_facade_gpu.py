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


def test_sml_file_matches_the_reference_image(tmp_path):
    """DNAFileSML::Create writes the file FileSML::Create would (FileSML.cpp:344-366): same header fields at the
    same offsets, same packed words, same positions; only the bytes the reference leaves uninitialised
    (word_size, little_endian, struct padding) are excluded.  LoadFile reads the reference's image back."""
    from checkers import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref not built")
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(DEMO):
        import __graft_entry__
        __graft_entry__.build()
    R = Reference()
    g = synth.genome_family(1, 40_001, seed=35)[0]
    seed = mems.get_seed(15)
    want, layout = R.sml_file_image(g, seed)
    r = subprocess.run([DEMO, "smllayout"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    assert [int(x) for x in r.stdout.split()] == layout[:15]
    raw = tmp_path / "g.raw"
    raw.write_bytes(g.tobytes())
    out = tmp_path / "g.sml"
    r = subprocess.run([DEMO, "writesml", "15", str(raw), str(out)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    got = out.read_bytes()
    assert len(got) == len(want)
    size, off = layout[0], dict(zip(["version", "alphabet_bits", "seed", "seed_length", "seed_weight", "length", "unique_mers",
                                     "word_size", "little_endian", "id", "circular", "translation_table", "description"],
                                    layout[1:14]))
    defined = [("version", 4), ("alphabet_bits", 4), ("seed", 8), ("seed_length", 4), ("seed_weight", 4), ("length", 8),
               ("unique_mers", 4), ("id", 2), ("circular", 1), ("translation_table", 255), ("description", 1)]
    for name, n in defined:
        assert got[off[name]:off[name] + n] == want[off[name]:off[name] + n], name
    # packed words are identical; positions list the same mers in the same order (std::sort leaves ties unspecified)
    # (the two pad words are zero here; translate32 never writes them, SortedMerList.cpp:425-460)
    n_words = (len(g) * 2 + 31) // 32 + 2
    assert got[size:size + 4 * (n_words - 2)] == want[size:size + 4 * (n_words - 2)]
    assert got[size + 4 * (n_words - 2):size + 4 * n_words] == bytes(8)
    import numpy as np
    mine = np.frombuffer(got[size + 4 * n_words:], np.uint32)
    theirs = np.frombuffer(want[size + 4 * n_words:], np.uint32)
    assert sorted(mine.tolist()) == sorted(theirs.tolist())
    pos, mers = R.sml_build(g, seed)
    mer_at = dict(zip(pos.tolist(), mers.tolist()))
    assert [mer_at[p] for p in mine.tolist()] == [mer_at[p] for p in theirs.tolist()]
    # the reference's image through our reader
    ref_file = tmp_path / "ref.sml"
    ref_file.write_bytes(want)
    r = subprocess.run([DEMO, "loadsml", str(ref_file)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    f = r.stdout.split()
    assert int(f[0], 16) == seed and int(f[1]) == len(g) and int(f[2]) == len(pos)
    assert f[3] == "%d:%d" % (pos[0], mers[0]) or int(f[3].split(":")[1]) == int(mers[0])
    # a truncated file is rejected like FileSML::LoadFile does
    bad = tmp_path / "bad.sml"
    bad.write_bytes(want[:size + 100])
    r = subprocess.run([DEMO, "loadsml", str(bad)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "sequence data" in r.stderr


def test_load_smls_creates_then_reuses_the_files(tmp_path):
    """MatchList::LoadSMLs (MatchList.h:262-349): the first run creates <seq>.sml, the second loads them."""
    gs = synth.genome_family(3, 25000, seed=36)
    want, _ = Oracle().find_matches(0, gs, mems.get_seed(15))
    got, log = run_demo(tmp_path, "smlmemhash", 15, gs)
    assert got == want
    assert log.count("Creating sorted mer list") == 3
    assert all(os.path.exists(str(tmp_path / ("seq%d.raw.sml" % i))) for i in range(3))
    got, log = run_demo(tmp_path, "smlmemhash", 15, gs)
    assert got == want
    assert log.count("Sorted mer list loaded successfully") == 3 and "Creating" not in log
    got, log = run_demo(tmp_path, "smlmemhash", 13, gs)  # other weight: seed mismatch, lists are recreated
    assert got == Oracle().find_matches(0, gs, mems.get_seed(13))[0]
    assert log.count("Default seed mismatch") == 3


# ---- façade members beyond FindMatches, against what the unmodified reference produced (tests/golden/facade.json) ----
import json  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _facade_cases():
    return json.load(open(os.path.join(GOLD, "facade.json")))


def _write_seqs(tmp_path, seqs):
    files = []
    for i, s in enumerate(seqs):
        p = tmp_path / ("seq%d.raw" % i)
        p.write_text(s)
        files.append(str(p))
    return files


def _demo(args):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(DEMO):
        import __graft_entry__
        __graft_entry__.build()
    r = subprocess.run([DEMO] + args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    return r.stdout


def _parse_matches(text):
    out = []
    for line in text.splitlines():
        v = [int(x) for x in line.split("\t")]
        out.append((len(v) - 1, v[0]) + tuple(v[1:]))
    return out


def test_find_matches_from_position_c_abi():
    """MemHash::FindMatchesFromPosition through mems_match_params_t.start_points."""
    from gpu_util import gpu_context
    ctx = gpu_context()
    for case in _facade_cases():
        seqs = [s.encode() for s in case["seqs"]]
        smls = ctx.create_smls(seqs, case["seed"])
        flat, info = ctx.find_matches(smls, order=mems.ORDER_REFERENCE, start_points=case["start_points"])
        assert mems.flat_to_matches(flat) == [tuple(m) for m in case["matches"]], case["tag"]
        assert info["mem_count"] == case["mem_count"] and info["collisions"] == case["collisions"], case["tag"]
        flat, _ = ctx.find_matches(smls, order=mems.ORDER_CANONICAL, start_points=case["start_points"])
        assert mems.flat_to_matches(flat) == sorted(set(tuple(m) for m in case["matches"])), case["tag"]
    ctx.close()


def test_from_position_logs_and_mems_files_facade(tmp_path):
    for case in _facade_cases():
        d = tmp_path / case["tag"]
        d.mkdir()
        files = _write_seqs(d, case["seqs"])
        out = _demo(["frompos", str(case["weight"]), ",".join(str(x) for x in case["start_points"])] + files)
        body, rest = out.split("#counts ", 1)
        counts, rest = rest.split("\n#progress\n", 1)
        progress, rest = rest.split("#matchlog\n", 1)
        matchlog, offsets = rest.split("#offsets\n", 1)
        assert _parse_matches(body) == [tuple(m) for m in case["matches"]], case["tag"]
        assert counts.split() == [str(case["mem_count"]), str(case["collisions"])], case["tag"]
        assert progress == case["progress"], case["tag"]  # LogProgress: the reference's text, replayed
        # SetMatchLog: the same records (the reference logs in order of discovery, the façade in table order)
        assert sorted(matchlog.splitlines()) == sorted(case["match_log"].splitlines()), case["tag"]
        assert offsets == ""  # SetOffsetLog: written only when a search restarts, which neither does on these inputs
        # MemHash::WriteFile: byte for byte the reference's .mems text
        assert _demo(["memsfile", str(case["weight"])] + files) == case["mems_file"], case["tag"]
        # MemHash::LoadFile
        for name, load in case["loads"].items():
            f = d / (name + ".mems")
            f.write_text("\n".join(load["lines"]) + "\n")
            out = _demo(["loadmems", str(f)])
            body, counts = out.split("#counts ", 1)
            assert _parse_matches(body) == [tuple(m) for m in load["matches"]], (case["tag"], name)
            assert counts.split() == [str(load["mem_count"]), str(load["collisions"])], (case["tag"], name)


def test_sorted_mer_list_clone(tmp_path):
    g = synth.genome_family(1, 5000, seed=35)[0]
    p = tmp_path / "seq.raw"
    p.write_bytes(g.tobytes())
    out = _demo(["clone", "13", str(p)]).split()
    assert out[0] == "1" and out[1] == out[2]  # same first entry after the original is gone; GetSeedMer is the canonical mer
    # and through the C-ABI
    from gpu_util import gpu_context
    ctx = gpu_context()
    a = ctx.create_sml(g, mems.get_seed(13))
    b = a.clone()
    want = a.read()
    a.close()
    got = b.read()
    assert (got[0] == want[0]).all() and (got[1] == want[1]).all()
    ctx.close()
