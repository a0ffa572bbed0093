"""CPU-side checks of the C-ABI library: it loads, exports every declared symbol, its host-side seed
arithmetic matches the reference's golden table, and compute calls fail loudly without a GPU."""
import ctypes
import json
import os
import re

import pytest

import libmems_b200 as mems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mems_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mems_[a-z_0-9]+)\s*\(", text)))


def test_library_loads_and_exports_header_symbols():
    lib = mems.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(mems.EXPORTS) == declared


def test_seed_table_matches_reference_golden():
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "seeds.json")))
    for key, (seed, length, weight) in g["seeds"].items():
        w, r = (int(x) for x in key.split(","))
        assert mems.get_seed(w, r) == seed
        assert mems.get_seed_length(seed) == length
        assert mems.get_seed_weight(seed) == weight
    for n, w in g["default_weight"].items():
        assert mems.get_default_seed_weight(int(n)) == w
    assert mems.get_seed(40, 0) == (1 << 32) - 1  # weight > 31 -> solid 32 (SeedMasks.h:309-310)
    assert mems.get_seed(13, 9) == (1 << 13) - 1  # rank > 5 -> solid


def test_arena_bookkeeping_selftest():
    """The device-memory arena of a context (csrc/context.cu) is plain host bookkeeping over cudaMalloc'ed slabs: its
    randomised self-check (made-up addresses, no device) must hold for blocks never overlapping or leaving their slab,
    free neighbours merging, whole slabs coming back once everything is returned."""
    lib = mems.load()
    lib.mems_selftest_arena.argtypes = [ctypes.c_uint64, ctypes.c_int]
    lib.mems_selftest_arena.restype = ctypes.c_int
    for seed in range(8):
        assert lib.mems_selftest_arena(seed, 20000) == 0, seed


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mems.MemsError) as e:
        mems.Context(0)
    assert e.value.code == 3  # MEMS_ERR_CUDA


def test_sml_header_layout_is_the_reference_s():
    """The façade's copy of SMLHeader (.sml files, SortedMerList.h:48-63) has the size and field offsets of the
    reference's struct as compiled against the shim (oracle/_ref), and the image the reference would write for
    a small sequence has the expected three parts."""
    import subprocess
    from checkers import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref not built")
    demo = os.path.join(ROOT, "build", "facade_demo")
    if not os.path.exists(demo):
        import __graft_entry__
        __graft_entry__.build()
    r = subprocess.run([demo, "smllayout"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    seq = b"ACGTTGCAAGGCTTAACCGGTTAGCATCGATCGGATCCGATTACAGGCAT" * 3
    image, layout = Reference().sml_file_image(seq, mems.get_seed(11))
    assert [int(x) for x in r.stdout.split()] == layout[:15]
    n_words = (len(seq) * 2 + 31) // 32 + 2
    n_pos = len(seq) - mems.get_seed_length(mems.get_seed(11)) + 1
    assert len(image) == layout[0] + 4 * n_words + 4 * n_pos
