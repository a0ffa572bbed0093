"""libmems_b200 — ctypes binding to the B200-native anchoring path (include/mems_b200.h).

This module is plumbing only: it loads ``libmems_b200.so`` (hand-written sm_100a CUDA kernels behind
a C-ABI) and exposes the reference's vocabulary — sorted mer lists (``SortedMerList``), ``MemHash`` /
``RepeatHash`` match finding, match lists.  There is NO CPU fallback: importing works without a GPU
(so the symbol table can be checked), but every compute call raises ``MemsError`` unless an sm_100
device is present, and ``load()`` raises if the shared library has not been built.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmems_b200.so")

MODE_MEMHASH, MODE_REPEAT, MODE_PAIRWISE = 0, 1, 2
ORDER_ANY, ORDER_REFERENCE, ORDER_CANONICAL = 0, 1, 2

_u64 = ctypes.c_uint64
_vp = ctypes.c_void_p


class MemsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mems error %d: %s" % (code, msg))
        self.code = code


class SmlInfo(ctypes.Structure):
    _fields_ = [("length", _u64), ("sml_length", _u64), ("seed", _u64), ("seed_length", ctypes.c_uint32),
                ("seed_weight", ctypes.c_uint32), ("seed_mask", _u64), ("mer_mask", _u64)]


class MatchParams(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_int), ("order", ctypes.c_int), ("table_size", ctypes.c_uint32),
                ("reserved", ctypes.c_uint32), ("table", _vp), ("seq_mask", _u64), ("start_points", _vp)]


class MatchesInfo(ctypes.Structure):
    _fields_ = [("n_matches", _u64), ("n_flat", _u64), ("n_hits", _u64), ("mem_count", _u64), ("collisions", _u64),
                ("max_run", _u64), ("n_segments", _u64), ("seq_count", ctypes.c_uint32), ("seed_length", ctypes.c_uint32),
                ("host_replay_ms", ctypes.c_double)]


class ProfileEntry(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 48), ("launches", _u64), ("ms", ctypes.c_double), ("bytes", ctypes.c_double)]


# every symbol include/mems_b200.h declares
EXPORTS = [
    "mems_get_seed", "mems_get_solid_seed", "mems_get_seed_length", "mems_get_seed_weight",
    "mems_get_default_seed_weight", "mems_ctx_create", "mems_ctx_destroy", "mems_last_error",
    "mems_ctx_synchronize", "mems_host_alloc", "mems_host_free", "mems_sml_create", "mems_sml_create_batch",
    "mems_sml_destroy", "mems_sml_clone", "mems_sml_info", "mems_sml_read", "mems_sml_seed_mers", "mems_sml_find_mer",
    "mems_sml_packed", "mems_sml_seed_occurrence", "mems_find_matches", "mems_find_matches_many", "mems_table_create", "mems_table_clear", "mems_table_destroy", "mems_table_add", "mems_table_matches", "mems_matches_info", "mems_matches_copy", "mems_matches_data", "mems_matches_wait", "mems_matches_destroy", "mems_selftest_arena", "mems_ctx_trim",
    "mems_comm_unique_id", "mems_comm_create", "mems_comm_destroy", "mems_shard_sequence_range",
    "mems_shard_bucket_owners", "mems_shard_exchange_plan", "mems_find_matches_sharded", "mems_profile_enable", "mems_profile_reset", "mems_profile_get", "mems_launch_count",
    "mems_test_hooks",
]

_lib = None


def load():
    """Load the CUDA library; raises if it was not built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MemsError(-1, "%s not found — build it with `make -C libmems_b200/csrc` "
                            "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.mems_get_seed.restype = _u64
    lib.mems_get_seed.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.mems_get_solid_seed.restype = _u64
    lib.mems_get_seed_length.argtypes = [_u64]
    lib.mems_get_seed_weight.argtypes = [_u64]
    lib.mems_get_default_seed_weight.argtypes = [_u64]
    lib.mems_get_default_seed_weight.restype = ctypes.c_uint
    lib.mems_last_error.restype = ctypes.c_char_p
    lib.mems_last_error.argtypes = [_vp]
    lib.mems_ctx_create.argtypes = [ctypes.c_int, _vp, ctypes.POINTER(_vp)]
    lib.mems_ctx_destroy.argtypes = [_vp]
    lib.mems_ctx_synchronize.argtypes = [_vp]
    lib.mems_host_alloc.argtypes = [ctypes.POINTER(_vp), _u64]
    lib.mems_host_free.argtypes = [_vp]
    lib.mems_sml_create.argtypes = [_vp, _vp, _u64, _u64, ctypes.POINTER(_vp)]
    lib.mems_sml_create_batch.argtypes = [_vp, ctypes.c_int, ctypes.POINTER(_vp), ctypes.POINTER(_u64), _u64,
                                          ctypes.POINTER(_vp)]
    lib.mems_sml_destroy.argtypes = [_vp]
    lib.mems_sml_clone.argtypes = [_vp, ctypes.POINTER(_vp)]
    lib.mems_sml_info.argtypes = [_vp, ctypes.POINTER(SmlInfo)]
    lib.mems_sml_read.argtypes = [_vp, _u64, _u64, _vp, _vp, ctypes.POINTER(_u64)]
    lib.mems_sml_seed_mers.argtypes = [_vp, _vp, _u64, _vp, _vp]
    lib.mems_sml_find_mer.argtypes = [_vp, _u64, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(_u64)]
    lib.mems_sml_packed.argtypes = [_vp, _vp, ctypes.POINTER(_u64)]
    lib.mems_sml_seed_occurrence.argtypes = [_vp, _vp]
    lib.mems_find_matches.argtypes = [_vp, ctypes.c_int, ctypes.POINTER(_vp), ctypes.POINTER(MatchParams),
                                      ctypes.POINTER(_vp)]
    lib.mems_find_matches_many.argtypes = [_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(_vp),
                                           ctypes.POINTER(_u64), _u64, ctypes.POINTER(MatchParams), ctypes.POINTER(_vp)]
    lib.mems_table_create.argtypes = [ctypes.c_uint32, ctypes.POINTER(_vp)]
    lib.mems_table_clear.argtypes = [_vp]
    lib.mems_table_destroy.argtypes = [_vp]
    lib.mems_table_add.argtypes = [_vp, ctypes.c_uint32, _u64, _vp, ctypes.c_uint32, ctypes.POINTER(ctypes.c_int)]
    lib.mems_table_matches.argtypes = [_vp, ctypes.POINTER(_vp)]
    lib.mems_matches_info.argtypes = [_vp, ctypes.POINTER(MatchesInfo)]
    lib.mems_matches_copy.argtypes = [_vp, _vp]
    lib.mems_matches_data.argtypes = [_vp]
    lib.mems_matches_wait.argtypes = [_vp]
    lib.mems_matches_data.restype = ctypes.POINTER(ctypes.c_int64)
    lib.mems_matches_destroy.argtypes = [_vp]
    lib.mems_comm_unique_id.argtypes = [_vp]
    lib.mems_comm_create.argtypes = [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_vp)]
    lib.mems_comm_destroy.argtypes = [_vp]
    lib.mems_shard_sequence_range.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                              ctypes.POINTER(ctypes.c_int)]
    lib.mems_shard_bucket_owners.argtypes = [_vp, ctypes.c_int, _vp]
    lib.mems_shard_exchange_plan.argtypes = [_vp, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, ctypes.POINTER(_u64)]
    lib.mems_find_matches_sharded.argtypes = [_vp, _vp, ctypes.c_int, ctypes.POINTER(_vp), ctypes.POINTER(_u64), _u64,
                                              ctypes.POINTER(MatchParams), ctypes.POINTER(_vp)]
    lib.mems_profile_enable.argtypes = [_vp, ctypes.c_int]
    lib.mems_profile_reset.argtypes = [_vp]
    lib.mems_profile_get.argtypes = [_vp, _vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    lib.mems_launch_count.argtypes = [_vp]
    lib.mems_launch_count.restype = _u64
    lib.mems_test_hooks.argtypes = [_vp, ctypes.c_int, ctypes.c_int]
    _lib = lib
    return lib


# ---- seed patterns (SeedMasks.h) --------------------------------------------------------------------
def get_seed(weight, rank=0):
    return int(load().mems_get_seed(int(weight), int(rank)))


def get_seed_length(seed):
    return int(load().mems_get_seed_length(seed))


def get_seed_weight(seed):
    return int(load().mems_get_seed_weight(seed))


def get_default_seed_weight(avg_len):
    return int(load().mems_get_default_seed_weight(int(avg_len)))


def shard_sequence_range(n_seqs, rank, world):
    """(first, count): the contiguous block of sequences a rank extracts when the path is sharded."""
    f, c = ctypes.c_int(), ctypes.c_int()
    rc = load().mems_shard_sequence_range(n_seqs, rank, world, ctypes.byref(f), ctypes.byref(c))
    if rc:
        raise MemsError(rc, "bad shard arguments")
    return f.value, c.value


def shard_bucket_owners(hist256, world):
    """Owner rank of each of the 256 top-key-digit buckets, balanced by the global histogram."""
    h = np.ascontiguousarray(hist256, dtype=np.uint64)
    assert h.size == 256
    out = np.zeros(256, np.uint8)
    rc = load().mems_shard_bucket_owners(h.ctypes.data, world, out.ctypes.data)
    if rc:
        raise MemsError(rc, "bad shard arguments")
    return out


def shard_exchange_plan(hist_all, world, rank, owners):
    """Count matrix [sender, receiver], this rank's slice starts (source / destination) and the largest receive region
    of the seed-record exchange, from the gathered (world x 256) top-digit histograms."""
    h = np.ascontiguousarray(hist_all, dtype=np.uint32).reshape(world, 256)
    o = np.ascontiguousarray(owners, dtype=np.uint8)
    counts = np.zeros((world, world), np.uint64)
    src, dst = np.zeros(world, np.uint64), np.zeros(world, np.uint64)
    mx = ctypes.c_uint64()
    rc = load().mems_shard_exchange_plan(h.ctypes.data, world, rank, o.ctypes.data, counts.ctypes.data, src.ctypes.data,
                                         dst.ctypes.data, ctypes.byref(mx))
    if rc:
        raise MemsError(rc, "bad shard arguments")
    return counts, src, dst, mx.value


def comm_unique_id():
    buf = ctypes.create_string_buffer(128)
    rc = load().mems_comm_unique_id(ctypes.addressof(buf))
    if rc:
        raise MemsError(rc, load().mems_last_error(None).decode())
    return buf.raw


def _host_ptr(seq):
    """(pointer, length, keepalive) for bytes / numpy uint8 / (ptr, len) inputs."""
    if isinstance(seq, np.ndarray):
        a = np.ascontiguousarray(seq, dtype=np.uint8)
        return a.ctypes.data, a.size, a
    if isinstance(seq, tuple):
        return int(seq[0]), int(seq[1]), None
    b = bytes(seq)
    buf = ctypes.create_string_buffer(b, len(b)) if len(b) else ctypes.create_string_buffer(1)
    return ctypes.addressof(buf), len(b), buf


class _MatchHandle:
    def __init__(self, lib, h):
        self.lib, self.h = lib, h

    def __del__(self):
        try:
            self.lib.mems_matches_destroy(self.h)
        except Exception:
            pass


class PendingMatches:
    """The records of a call made with wait=False: they are still on their way from the device (mems_b200.h,
    mems_matches_wait).  records() waits for them and returns the zero-copy view."""

    def __init__(self, lib, keep, n_flat):
        self.lib, self._keep, self.n_flat = lib, keep, n_flat

    def records(self):
        if self.n_flat == 0:
            return np.zeros(0, np.int64)
        view = np.ctypeslib.as_array(self.lib.mems_matches_data(self._keep.h), shape=(self.n_flat,))  # waits
        flat = view.view(_OwnedArray)
        flat._keep = self._keep
        return flat

    def wait(self):
        self.lib.mems_matches_wait(self._keep.h)
        return self


class _OwnedArray(np.ndarray):
    """ndarray view that keeps the owning match handle alive."""
    _keep = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self._keep = getattr(obj, "_keep", None)


class Context:
    """One stream + scratch pool; the analogue of holding one MemHash per thread (Aligner.h:198)."""

    def __init__(self, device=0, stream=None):
        self.lib = load()
        h = _vp()
        rc = self.lib.mems_ctx_create(int(device), _vp(stream) if stream else None, ctypes.byref(h))
        if rc:
            raise MemsError(rc, self.lib.mems_last_error(None).decode())
        self.h = h

    def _check(self, rc):
        if rc:
            raise MemsError(rc, self.lib.mems_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.mems_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self._check(self.lib.mems_ctx_synchronize(self.h))

    def trim(self):
        """Give the device memory no live object occupies back to the driver; returns the bytes still held."""
        left = _u64()
        self._check(self.lib.mems_ctx_trim(self.h, ctypes.byref(left)))
        return int(left.value)

    # -- SML construction -----------------------------------------------------------------------------
    def create_sml(self, seq, seed):
        """DNAMemorySML::Create (MemorySML.cpp:45-60)."""
        p, n, keep = _host_ptr(seq)
        h = _vp()
        self._check(self.lib.mems_sml_create(self.h, p, n, seed, ctypes.byref(h)))
        return SortedMerList(self, h)

    def create_smls(self, seqs, seed):
        """MatchList::CreateMemorySMLs (MatchList.h:408-435): one SML per sequence, built as one batch."""
        parts = [_host_ptr(s) for s in seqs]
        n = len(parts)
        ptrs = (_vp * n)(*[p[0] for p in parts])
        lens = (_u64 * n)(*[p[1] for p in parts])
        out = (_vp * n)()
        self._check(self.lib.mems_sml_create_batch(self.h, n, ptrs, lens, seed, out))
        return [SortedMerList(self, _vp(out[i])) for i in range(n)]

    # -- match finding --------------------------------------------------------------------------------
    def find_matches(self, smls, mode=MODE_MEMHASH, order=ORDER_ANY, table_size=0, seq_mask=0, table=None, start_points=None,
                     wait=True):
        """MemHash / RepeatHash / PairwiseMatchFinder ::FindMatches (start_points: FindMatchesFromPosition).  Returns
        (flat, info): flat = int64 records [SeqCount, Length, Start(0), ...] (see flat_to_matches); with wait=False a
        PendingMatches in place of flat (the records may still be on their way from the device)."""
        n = len(smls)
        arr = (_vp * n)(*[s.h for s in smls])
        sp = (_u64 * n)(*[int(x) for x in start_points]) if start_points is not None else None
        params = MatchParams(mode, order, table_size, 0, table.h if table is not None else None, seq_mask,
                             ctypes.cast(sp, _vp) if sp is not None else None)
        h = _vp()
        self._check(self.lib.mems_find_matches(self.h, n, arr, ctypes.byref(params), ctypes.byref(h)))
        keep = _MatchHandle(self.lib, h)
        info = MatchesInfo()
        self._check(self.lib.mems_matches_info(h, ctypes.byref(info)))
        d = {k: (float if k == "host_replay_ms" else int)(getattr(info, k)) for k, _ in MatchesInfo._fields_}
        if not wait:
            return PendingMatches(self.lib, keep, int(info.n_flat)), d
        if info.n_flat == 0:
            return np.zeros(0, np.int64), d
        # zero-copy view of the library's (page-locked) result buffer; it lives as long as the array does
        view = np.ctypeslib.as_array(self.lib.mems_matches_data(h), shape=(int(info.n_flat),))
        flat = view.view(_OwnedArray)
        flat._keep = keep
        return flat, d

    def find_matches_many(self, problems, seed, mode=MODE_MEMHASH, order=ORDER_ANY):
        """Many independent small problems (each a list of sequences) in one launch set; returns one (flat, info) per
        problem, equal to create_smls + find_matches on that problem alone."""
        parts = [_host_ptr(s) for prob in problems for s in prob]
        n = len(parts)
        counts = (ctypes.c_int * len(problems))(*[len(prob) for prob in problems])
        ptrs = (_vp * n)(*[p[0] for p in parts])
        lens = (_u64 * n)(*[p[1] for p in parts])
        params = MatchParams(mode, order, 0, 0, None, 0, None)
        out = (_vp * len(problems))()
        self._check(self.lib.mems_find_matches_many(self.h, len(problems), counts, ptrs, lens, seed, ctypes.byref(params), out))
        res = []
        for g in range(len(problems)):
            h = _vp(out[g])
            info = MatchesInfo()
            self._check(self.lib.mems_matches_info(h, ctypes.byref(info)))
            d = {k: (float if k == "host_replay_ms" else int)(getattr(info, k)) for k, _ in MatchesInfo._fields_}
            flat = np.zeros(0, np.int64)
            if info.n_flat:
                flat = np.ctypeslib.as_array(self.lib.mems_matches_data(h), shape=(int(info.n_flat),)).copy()
            self.lib.mems_matches_destroy(h)
            res.append((flat, d))
        return res

    # -- sharded (one process per GPU) ----------------------------------------------------------------
    def create_comm(self, unique_id, rank, world):
        h = _vp()
        buf = ctypes.create_string_buffer(unique_id, 128)
        self._check(self.lib.mems_comm_create(self.h, ctypes.addressof(buf), rank, world, ctypes.byref(h)))
        return Communicator(self, h, rank, world)

    def find_matches_sharded(self, comm, seqs, lens, seed, mode=MODE_MEMHASH, order=ORDER_ANY, wait=True):
        """Collective.  seqs: list over ALL sequences; entries outside this rank's block may be None.
        Returns this rank's share of the matches (flat, info); wait=False as in find_matches."""
        n = len(lens)
        parts = [(_host_ptr(s) if s is not None else (0, 0, None)) for s in seqs]
        ptrs = (_vp * n)(*[p[0] for p in parts])
        ls = (_u64 * n)(*[int(x) for x in lens])
        params = MatchParams(mode, order, 0, 0, None, 0, None)
        h = _vp()
        self._check(self.lib.mems_find_matches_sharded(self.h, comm.h, n, ptrs, ls, seed, ctypes.byref(params),
                                                       ctypes.byref(h)))
        keep = _MatchHandle(self.lib, h)
        info = MatchesInfo()
        self._check(self.lib.mems_matches_info(h, ctypes.byref(info)))
        d = {k: (float if k == "host_replay_ms" else int)(getattr(info, k)) for k, _ in MatchesInfo._fields_}
        if not wait:
            return PendingMatches(self.lib, keep, int(info.n_flat)), d
        if info.n_flat == 0:
            return np.zeros(0, np.int64), d
        view = np.ctypeslib.as_array(self.lib.mems_matches_data(h), shape=(int(info.n_flat),))
        flat = view.view(_OwnedArray)
        flat._keep = keep
        return flat, d

    # -- measurement ----------------------------------------------------------------------------------
    def profile_enable(self, on=True):
        self._check(self.lib.mems_profile_enable(self.h, 1 if on else 0))

    def profile_reset(self):
        self._check(self.lib.mems_profile_reset(self.h))

    def profile(self):
        n = ctypes.c_int()
        ent = (ProfileEntry * 64)()
        self._check(self.lib.mems_profile_get(self.h, ent, 64, ctypes.byref(n)))
        return {ent[i].name.decode(): {"launches": int(ent[i].launches), "ms": float(ent[i].ms),
                                       "bytes": float(ent[i].bytes)} for i in range(min(n.value, 64))}

    def launch_count(self):
        return int(self.lib.mems_launch_count(self.h))

    def set_test_hooks(self, hash_bits=0, walk_budget=0):
        """Testing only: shrink the diagonal hash / the walk budgets of this context (see mems_test_hooks)."""
        self._check(self.lib.mems_test_hooks(self.h, int(hash_bits), int(walk_budget)))


class HashTable:
    """Persistent MemHash table: pass it to find_matches(table=...) to accumulate several seed patterns."""

    def __init__(self, table_size=0):
        self.lib = load()
        h = _vp()
        rc = self.lib.mems_table_create(table_size, ctypes.byref(h))
        if rc:
            raise MemsError(rc, "cannot create table")
        self.h = h

    def clear(self):
        self.lib.mems_table_clear(self.h)

    def add(self, match, mersize=0):
        """AddHashEntry of an already extended match (SeqCount, Length, starts...); True if it was inserted."""
        st = (ctypes.c_int64 * int(match[0]))(*[int(x) for x in match[2:]])
        ins = ctypes.c_int()
        rc = self.lib.mems_table_add(self.h, int(match[0]), int(match[1]), st, int(mersize), ctypes.byref(ins))
        if rc:
            raise MemsError(rc, "bad table entry")
        return bool(ins.value)

    def matches(self):
        """The table's content in the reference's output order, as a list of tuples."""
        h = _vp()
        rc = self.lib.mems_table_matches(self.h, ctypes.byref(h))
        if rc:
            raise MemsError(rc, "cannot list table")
        info = MatchesInfo()
        self.lib.mems_matches_info(h, ctypes.byref(info))
        flat = np.zeros(0, np.int64)
        if info.n_flat:
            flat = np.ctypeslib.as_array(self.lib.mems_matches_data(h), shape=(int(info.n_flat),)).copy()
        self.lib.mems_matches_destroy(h)
        return flat_to_matches(flat)

    def __del__(self):
        try:
            self.lib.mems_table_destroy(self.h)
        except Exception:
            pass


class Communicator:
    """NCCL communicator of the sharded path (one per process / GPU)."""

    def __init__(self, ctx, handle, rank, world):
        self.ctx, self.h, self.rank, self.world = ctx, handle, rank, world

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.mems_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def flat_to_matches(flat):
    """[SeqCount, Len, starts...]* -> list of tuples."""
    out, i, n = [], 0, len(flat)
    while i < n:
        k = int(flat[i])
        out.append(tuple(int(x) for x in flat[i:i + 2 + k]))
        i += 2 + k
    return out


class SortedMerList:
    """Device-resident sorted mer list (SortedMerList.h:69-282 accessors)."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle
        info = SmlInfo()
        ctx._check(ctx.lib.mems_sml_info(self.h, ctypes.byref(info)))
        self.info = {k: int(getattr(info, k)) for k, _ in SmlInfo._fields_}

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.mems_sml_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sml_length(self):
        return self.info["sml_length"]

    def clone(self):
        """MemorySML::Clone: another handle to the same device-resident list."""
        h = _vp()
        self.ctx._check(self.ctx.lib.mems_sml_clone(self.h, ctypes.byref(h)))
        return SortedMerList(self.ctx, h)

    def read(self, offset=0, count=None):
        """MemorySML::Read: (positions, mers) of sorted-list entries [offset, offset+count)."""
        if count is None:
            count = self.info["sml_length"]
        pos = np.zeros(max(count, 1), np.uint32)
        mers = np.zeros(max(count, 1), np.uint64)
        n = _u64()
        self.ctx._check(self.ctx.lib.mems_sml_read(self.h, offset, count, pos.ctypes.data, mers.ctypes.data,
                                                   ctypes.byref(n)))
        return pos[:n.value], mers[:n.value]

    def seed_mers(self, positions):
        """(GetSeedMer forward, GetDnaSeedMer canonical) at the given positions."""
        p = np.ascontiguousarray(positions, dtype=np.uint64)
        fwd = np.zeros(max(len(p), 1), np.uint64)
        dna = np.zeros(max(len(p), 1), np.uint64)
        self.ctx._check(self.ctx.lib.mems_sml_seed_mers(self.h, p.ctypes.data, len(p), fwd.ctypes.data, dna.ctypes.data))
        return fwd[:len(p)], dna[:len(p)]

    def find_mer(self, mer):
        found = ctypes.c_int()
        idx = _u64()
        self.ctx._check(self.ctx.lib.mems_sml_find_mer(self.h, mer, ctypes.byref(found), ctypes.byref(idx)))
        return bool(found.value), int(idx.value)

    def seed_occurrence(self):
        """SeedOccurrenceList::construct: smoothed per-position seed multiplicity (float32, one per base)."""
        out = np.zeros(max(self.info["length"], 1), np.float32)
        self.ctx._check(self.ctx.lib.mems_sml_seed_occurrence(self.h, out.ctypes.data))
        return out[:self.info["length"]]

    def packed(self):
        n = _u64()
        self.ctx._check(self.ctx.lib.mems_sml_packed(self.h, None, ctypes.byref(n)))
        w = np.zeros(n.value, np.uint32)
        self.ctx._check(self.ctx.lib.mems_sml_packed(self.h, w.ctypes.data, ctypes.byref(n)))
        return w
