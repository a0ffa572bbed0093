"""ctypes bindings to the CPU checkers: oracle/liboracle.so (C restatement) and, when present,
oracle/_ref/libmems_ref.so (the unmodified reference).  TEST INFRASTRUCTURE — never imported by the
product package."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
u64 = ctypes.c_uint64
_vp = ctypes.c_void_p


def _ptr(a):
    return a.ctypes.data_as(_vp) if a is not None else None


def _as_bytes(s):
    if isinstance(s, np.ndarray):
        return s.tobytes()
    return bytes(s)


def flat_to_matches(flat):
    """[SeqCount, Len, starts...]* -> list of tuples (SeqCount, Len, start0, start1, ...)"""
    out = []
    i = 0
    n = len(flat)
    while i < n:
        k = int(flat[i])
        out.append(tuple(int(x) for x in flat[i:i + 2 + k]))
        i += 2 + k
    return out


class _Checker:
    prefix = ""

    def __init__(self, path):
        self.lib = ctypes.CDLL(path)
        p = self.prefix
        getattr(self.lib, p + "get_seed").restype = u64
        getattr(self.lib, p + "last_error").restype = ctypes.c_char_p

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def err(self):
        return self._f("last_error")().decode()

    def get_seed(self, weight, rank=0):
        return int(self._f("get_seed")(int(weight), int(rank)))

    def seed_length(self, seed):
        return int(self._f("seed_length")(u64(seed)))

    def seed_weight(self, seed):
        return int(self._f("seed_weight")(u64(seed)))

    def default_seed_weight(self, avg_len):
        return int(self._f("default_seed_weight")(u64(avg_len)))

    def seed_mers(self, seq, seed, positions):
        s = _as_bytes(seq)
        pos = np.ascontiguousarray(positions, dtype=np.uint64)
        fwd = np.zeros(len(pos), np.uint64)
        dna = np.zeros(len(pos), np.uint64)
        rc = self._f("seed_mers")(s, u64(len(s)), u64(seed), _ptr(pos), u64(len(pos)), _ptr(fwd), _ptr(dna))
        if rc:
            raise RuntimeError(self.err())
        return fwd, dna

    def find_matches(self, mode, seqs, seed):
        bufs = [_as_bytes(s) for s in seqs]
        arr = (ctypes.c_char_p * len(bufs))(*bufs)
        lens = (u64 * len(bufs))(*[len(b) for b in bufs])
        flat = ctypes.POINTER(ctypes.c_int64)()
        nflat, nm = u64(), u64()
        return bufs, arr, lens, flat, nflat, nm


class Oracle(_Checker):
    prefix = "orc_"

    def __init__(self):
        super().__init__(os.path.join(ROOT, "oracle", "liboracle.so"))
        self.lib.orc_packed_words.restype = u64

    def pack(self, seq):
        s = _as_bytes(seq)
        nw = int(self.lib.orc_packed_words(u64(len(s))))
        words = np.zeros(nw, np.uint32)
        if self.lib.orc_pack(s, u64(len(s)), _ptr(words)):
            raise RuntimeError(self.err())
        return words

    def sml_build(self, seq, seed):
        s = _as_bytes(seq)
        cap = max(len(s), 1)
        pos = np.zeros(cap, np.uint32)
        mers = np.zeros(cap, np.uint64)
        n = u64()
        if self.lib.orc_sml_build(s, u64(len(s)), u64(seed), _ptr(pos), _ptr(mers), ctypes.byref(n)):
            raise RuntimeError(self.err())
        return pos[:n.value].copy(), mers[:n.value].copy()

    def find_matches(self, mode, seqs, seed):
        bufs, arr, lens, flat, nflat, nm = super().find_matches(mode, seqs, seed)
        counts = (u64 * 4)()
        rc = self.lib.orc_find_matches(int(mode), len(bufs), arr, lens, u64(seed), ctypes.byref(flat),
                                       ctypes.byref(nflat), ctypes.byref(nm), counts)
        if rc:
            raise RuntimeError(self.err())
        out = np.ctypeslib.as_array(flat, shape=(max(nflat.value, 1),))[:nflat.value].copy()
        self.lib.orc_free(flat)
        info = {"mem_count": counts[0], "collisions": counts[1], "hits": counts[2], "max_run": counts[3]}
        return flat_to_matches(out), info

    def find_matches_multi_seed(self, seqs, seeds):
        """One table across several seed patterns (ClearSequences between FindMatches calls)."""
        self.lib.orc_accumulate_begin()
        try:
            for sd in seeds:
                out = self.find_matches(0, seqs, sd)
            return out
        finally:
            self.lib.orc_accumulate_end()

    def find_matches_masked(self, seqs, seed, seq_mask):
        bufs, arr, lens, flat, nflat, nm = _Checker.find_matches(self, 0, seqs, seed)
        counts = (u64 * 4)()
        rc = self.lib.orc_find_matches_masked(len(bufs), arr, lens, u64(seed), u64(seq_mask), ctypes.byref(flat),
                                              ctypes.byref(nflat), ctypes.byref(nm), counts)
        if rc:
            raise RuntimeError(self.err())
        out = np.ctypeslib.as_array(flat, shape=(max(nflat.value, 1),))[:nflat.value].copy()
        self.lib.orc_free(flat)
        return flat_to_matches(out), {"mem_count": counts[0], "collisions": counts[1], "hits": counts[2]}

    def seed_occurrence(self, seq, seed):
        s = _as_bytes(seq)
        out = np.zeros(len(s), np.float32)
        if self.lib.orc_seed_occurrence(s, u64(len(s)), u64(seed), _ptr(out)):
            raise RuntimeError(self.err())
        return out


class Reference(_Checker):
    prefix = "ref_"
    PATH = os.path.join(ROOT, "oracle", "_ref", "libmems_ref.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self):
        super().__init__(self.PATH)

    def sml_build(self, seq, seed):
        s = _as_bytes(seq)
        cap = max(len(s), 1)
        pos = np.zeros(cap, np.uint32)
        mers = np.zeros(cap, np.uint64)
        n, sm, mm = u64(), u64(), u64()
        secs = ctypes.c_double()
        rc = self.lib.ref_sml_build(s, u64(len(s)), u64(seed), _ptr(pos), _ptr(mers), ctypes.byref(n),
                                    ctypes.byref(sm), ctypes.byref(mm), ctypes.byref(secs))
        if rc:
            raise RuntimeError(self.err())
        self.last = {"seed_mask": sm.value, "mer_mask": mm.value, "secs": secs.value}
        return pos[:n.value].copy(), mers[:n.value].copy()

    def sml_file_image(self, seq, seed):
        """Bytes of the .sml file FileSML::Create would write (header, packed words, positions) + header layout."""
        s = _as_bytes(seq)
        layout = np.zeros(16, np.uint32)
        self.lib.ref_sml_file_image.restype = ctypes.c_int64
        total = self.lib.ref_sml_file_image(s, u64(len(s)), u64(seed), None, u64(0), _ptr(layout))
        if total < 0:
            raise RuntimeError(self.err())
        buf = np.zeros(total, np.uint8)
        got = self.lib.ref_sml_file_image(s, u64(len(s)), u64(seed), _ptr(buf), u64(total), None)
        if got != total:
            raise RuntimeError(self.err())
        return buf.tobytes(), [int(x) for x in layout]

    def sml_time(self, seq, seed):
        s = _as_bytes(seq)
        n = u64()
        secs = ctypes.c_double()
        rc = self.lib.ref_sml_build(s, u64(len(s)), u64(seed), None, None, ctypes.byref(n), None, None,
                                    ctypes.byref(secs))
        if rc:
            raise RuntimeError(self.err())
        return secs.value

    def find_matches(self, mode, seqs, seed):
        bufs, arr, lens, flat, nflat, nm = super().find_matches(mode, seqs, seed)
        times = (ctypes.c_double * 2)()
        counts = (u64 * 2)()
        rc = self.lib.ref_find_matches(int(mode), len(bufs), arr, lens, u64(seed), ctypes.byref(flat),
                                       ctypes.byref(nflat), ctypes.byref(nm), times, counts)
        if rc:
            raise RuntimeError(self.err())
        out = np.ctypeslib.as_array(flat, shape=(max(nflat.value, 1),))[:nflat.value].copy()
        self.lib.ref_free(flat)
        info = {"mem_count": counts[0], "collisions": counts[1], "sml_s": times[0], "find_s": times[1]}
        return flat_to_matches(out), info

    def find_matches_multi_seed(self, seqs, seeds):
        self.lib.ref_accumulate_begin()
        try:
            for sd in seeds:
                out = self.find_matches(0, seqs, sd)
            return out
        finally:
            self.lib.ref_accumulate_end()

    def find_matches_masked(self, seqs, seed, seq_mask):
        self.lib.ref_set_seq_mask(u64(seq_mask))
        try:
            return self.find_matches(3, seqs, seed)
        finally:
            self.lib.ref_set_seq_mask(u64(0))

    def read_list(self, text):
        """Parse .mums text with the reference's ReadList."""
        flat = ctypes.POINTER(ctypes.c_int64)()
        nflat, nm = u64(), u64()
        if self.lib.ref_read_list(text.encode(), ctypes.byref(flat), ctypes.byref(nflat), ctypes.byref(nm)):
            raise RuntimeError(self.err())
        out = np.ctypeslib.as_array(flat, shape=(max(nflat.value, 1),))[:nflat.value].copy()
        self.lib.ref_free(flat)
        return flat_to_matches(out)

    def write_list(self, seqs, seed):
        """MemHash + the reference's WriteList -> .mums text."""
        bufs = [_as_bytes(s) for s in seqs]
        arr = (ctypes.c_char_p * len(bufs))(*bufs)
        lens = (u64 * len(bufs))(*[len(b) for b in bufs])
        text = ctypes.c_char_p()
        if self.lib.ref_write_list(len(bufs), arr, lens, u64(seed), ctypes.byref(text)):
            raise RuntimeError(self.err())
        return text.value.decode()

    def _seq_args(self, seqs):
        bufs = [_as_bytes(s) for s in seqs]
        arr = (ctypes.c_char_p * len(bufs))(*bufs)
        lens = (u64 * len(bufs))(*[len(b) for b in bufs])
        return bufs, arr, lens

    def _take_flat(self, flat, nflat):
        out = np.ctypeslib.as_array(flat, shape=(max(nflat.value, 1),))[:nflat.value].copy()
        self.lib.ref_free(flat)
        return flat_to_matches(out)

    def _take_text(self, p):
        text = ctypes.string_at(p).decode()
        self.lib.ref_free(p)
        return text

    def find_matches_from(self, seqs, seed, start_points):
        """MemHash::FindMatchesFromPosition with LogProgress / SetMatchLog attached: (matches, info)."""
        bufs, arr, lens = self._seq_args(seqs)
        sp = (u64 * len(bufs))(*[int(x) for x in start_points])
        flat = ctypes.POINTER(ctypes.c_int64)()
        nflat, nm = u64(), u64()
        counts = (u64 * 2)()
        prog, mlog = _vp(), _vp()
        rc = self.lib.ref_find_matches_from(len(bufs), arr, lens, u64(seed), sp, ctypes.byref(flat), ctypes.byref(nflat),
                                            ctypes.byref(nm), counts, ctypes.byref(prog), ctypes.byref(mlog))
        if rc:
            raise RuntimeError(self.err())
        return self._take_flat(flat, nflat), {"mem_count": counts[0], "collisions": counts[1],
                                              "progress": self._take_text(prog), "match_log": self._take_text(mlog)}

    def mems_write_file(self, seqs, seed):
        """MemHash::FindMatches + MemHash::WriteFile -> .mems text."""
        bufs, arr, lens = self._seq_args(seqs)
        text = _vp()
        if self.lib.ref_mems_write_file(len(bufs), arr, lens, u64(seed), ctypes.byref(text)):
            raise RuntimeError(self.err())
        return self._take_text(text)

    def mems_load_file(self, seqs, seed, text):
        """MemHash::LoadFile of bare match lines into a MemHash holding the sequences -> (matches, counters)."""
        bufs, arr, lens = self._seq_args(seqs)
        flat = ctypes.POINTER(ctypes.c_int64)()
        nflat, nm = u64(), u64()
        counts = (u64 * 2)()
        rc = self.lib.ref_mems_load_file(len(bufs), arr, lens, u64(seed), text.encode(), ctypes.byref(flat), ctypes.byref(nflat),
                                         ctypes.byref(nm), counts)
        if rc:
            raise RuntimeError(self.err())
        return self._take_flat(flat, nflat), {"mem_count": counts[0], "collisions": counts[1]}

    def eliminate_overlaps(self, matches):
        """The reference's EliminateOverlaps (Aligner.cpp:62-180) on a list of (SeqCount, Length, starts...) tuples."""
        flat_in = np.array([x for m in matches for x in m], dtype=np.int64)
        flat = ctypes.POINTER(ctypes.c_int64)()
        nflat, nm = u64(), u64()
        if self.lib.ref_eliminate_overlaps(_ptr(flat_in), u64(len(flat_in)), ctypes.byref(flat), ctypes.byref(nflat), ctypes.byref(nm)):
            raise RuntimeError(self.err())
        return self._take_flat(flat, nflat)

    def seed_occurrence(self, seq, seed):
        s = _as_bytes(seq)
        out = np.zeros(len(s), np.float32)
        n = u64()
        if self.lib.ref_seed_occurrence(s, u64(len(s)), u64(seed), _ptr(out), ctypes.byref(n)):
            raise RuntimeError(self.err())
        return out
