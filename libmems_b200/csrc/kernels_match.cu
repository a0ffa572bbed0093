// kernels_match.cu — match finding on the sorted union of a batch (MemHash / RepeatHash / PairwiseMatchFinder).
//
// Reference flow (CPU, sequential): MatchFinder::SearchRange k-way merges the SMLs and hands every
// equal-seed run to EnumerateMatches (MatchFinder.cpp:172-340); MemHash keeps runs that hold the seed at
// most once per sequence and in >= 2 sequences (MemHash.cpp:139-162), RepeatHash keeps every run of its
// single sequence (RepeatHash.cpp:34-62); each kept run is a "hit" that AddHashEntry drops if an
// already-stored match contains it on the same diagonal, else extends with ExtendMatch
// (MatchFinder.h:219-374) and stores (MemHash.cpp:209-251).
//
// The extended result of a hit is the connected component of matching seed windows on the hit's
// diagonal, where windows count as adjacent when they start <= L apart.  That
// turns the sequential table into data-parallel steps:
//   1. run scan     — a warp per 32 union entries decides the runs that start in its first lanes from ballots -> hit list (key order)
//   2. describe     — per hit: first member, 64-bit hash of its diagonal (member set, orientations,
//                     offsets) -> sort key (diagonal hash | first-member position)
//   3. sort         — hits by that key (radix_sort.cu): hits of one diagonal become contiguous, by position
//   4. segments     — neighbours on the same diagonal <= L apart are connected without looking at sequence
//   5. extend       — one warp per segment walks RIGHT from its last hit with the window test: a probe tests
//                     128 windows at once (4 per lane, one load of the position-ordered key array per member
//                     and window), the chain of matches <= L apart is followed lane-parallel (each lane asks
//                     "no match in my next L windows?", ballots find the first); the walk ends when it reaches
//                     the next segment of its diagonal (link) or a gap > L (component end).  Chains of linked
//                     segments are the components; only each chain's first segment walks LEFT.  Walks that
//                     outlast a warp's budget go to 16-warp CTAs (8192 windows per round), the few that outlast
//                     those to the whole grid (cooperative launch)
//   6. emit         — components -> [SeqCount, Length, starts] records; policies: MemHash (+ MaskedMemHash
//                     filter), RepeatHash, PairwiseMatchFinder; ORDER_REFERENCE replays the reference's hash
//                     table on the host (optionally into a persistent table shared by several calls)
//   sharded         — find_matches_sharded at the bottom: the same stages across ranks with two exchanges
//                     (peer-to-peer into CUDA-IPC exchange windows, NCCL as fallback)
// Hash collisions between diagonals only split segments (more window tests), never merge them: every
// merge decision compares the full member lists.
#include <algorithm>
#include <chrono>
#include <exception>
#include <functional>
#include <thread>
#include <cstdlib>
#include <cstdio>
#include <cstring>

#include <cooperative_groups.h>

#include "common.cuh"
#include "mems_b200.h"
#include "seed_dev.cuh"

namespace mems {

constexpr uint32_t kRunCap = MEMS_MER_REPEAT_LIMIT;  // longer runs are dropped and reported via max_run
constexpr uint16_t kFirstStrandBit = 0x8000;         // hit_len bit 15: strand of the hit's first member

struct MatchArgs {
	const void* keys;      // union compact keys
	const uint32_t* vals;  // union (seq << pos_bits) | pos
	uint32_t n;
	int pos_bits;
	uint32_t pos_mask;
	int n_seqs;            // sequences in the batch (all problems together)
	int max_group;         // sequences in its largest problem (= n_seqs unless the batch holds several, SeqMeta::group)
	int mode;
	uint64_t seq_set;  // MaskedMemHash filter: required member set, bit g = sequence g (0 = no filter)
	int hit_key_bits;      // hit sort key = top (hit_key_bits - pos_bits) bits of the diagonal hash, then the first member's position
	int test_hash_bits;  // 0 = use the full diagonal hash; n > 0 keeps only n bits (tests force bucket collisions with it)
	int warp_budget;     // probes one warp spends on a walk before a CTA takes over (kWarpProbeBudget; tests shrink it)
	int cta_budget;      // rounds one CTA spends before the whole grid takes over (kCtaRoundBudget; tests shrink it)
	const uint2* planes;   // bit planes of the packed sequences (Batch::planes); SeqMeta::word_off * 16 = a sequence's first base
	const SeqMeta* meta;
};

template <class KeyT>
__device__ __forceinline__ uint64_t masked_of(const void* keys, uint32_t i) {
	return (uint64_t)(reinterpret_cast<const KeyT*>(keys)[i] >> 1);
}
template <class KeyT>
__device__ __forceinline__ uint32_t strand_of(const void* keys, uint32_t i) {
	return (uint32_t)(reinterpret_cast<const KeyT*>(keys)[i] & 1);
}

// ------------------------------------------------------------------------------------------------ 1. runs -> hits
constexpr int kScanBlock = 256;
constexpr int kRunItems = 8;                         // consecutive union entries per thread
constexpr int kRunTile = kScanBlock * kRunItems;     // per CTA
constexpr uint32_t kRunOk = 0x80000000u;

// Equal-seed runs of the sorted union -> hits (MatchFinder::SearchRange, MatchFinder.cpp:253-337, and the policies
// MemHash::EnumerateMatches, MemHash.cpp:139-162 / RepeatHash.cpp:34-45 / MaskedMemHash.cpp:50-60), ONE kernel.
// A thread owns 8 consecutive union entries (128-bit loads) and decides every run that STARTS among them in one
// forward pass with a few registers of state: run length, the set of sequences seen, "a sequence occurred twice".  A
// run still open at the thread's last entry ends in the next thread's leading entries: their summary comes from the
// next lane by shuffle (long runs, and the last lane of a warp, follow the run through memory instead).
// The hits leave in union order: per-thread counts are scanned inside the CTA and every thread writes (first entry,
// length) of its hits to the CTA's own kRunTile / 2 slots of a staging array — no CTA waits for another (a chained
// look-back over ~20 k small tiles cost more than the kernel's whole memory time) — and hit_gather_kernel closes the gaps
// once the tile counts are scanned: 6 bytes per hit instead of a flag per union entry.
template <class KeyT>
__global__ void __launch_bounds__(kScanBlock)
run_hits_kernel(MatchArgs a, uint32_t* __restrict__ stage_start, uint16_t* __restrict__ stage_len, uint32_t* __restrict__ tile_hits,
                uint32_t* __restrict__ max_run) {
	__shared__ uint32_t s_tot[kScanBlock / 32], s_max[kScanBlock / 32];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t tile = blockIdx.x;
	const KeyT* __restrict__ keys = reinterpret_cast<const KeyT*>(a.keys);
	const uint64_t n = a.n;
	const uint64_t i0 = (uint64_t)tile * kRunTile + (uint64_t)tid * kRunItems;
	KeyT mk[kRunItems];
	uint32_t sq[kRunItems];
	if (i0 + kRunItems <= n && ((reinterpret_cast<uintptr_t>(keys) | reinterpret_cast<uintptr_t>(a.vals)) & 15u) == 0) {
		if (sizeof(KeyT) == 4) {
			const uint4 q0 = *reinterpret_cast<const uint4*>(keys + i0), q1 = *reinterpret_cast<const uint4*>(keys + i0 + 4);
			mk[0] = (KeyT)(q0.x >> 1); mk[1] = (KeyT)(q0.y >> 1); mk[2] = (KeyT)(q0.z >> 1); mk[3] = (KeyT)(q0.w >> 1);
			mk[4] = (KeyT)(q1.x >> 1); mk[5] = (KeyT)(q1.y >> 1); mk[6] = (KeyT)(q1.z >> 1); mk[7] = (KeyT)(q1.w >> 1);
		} else {
#pragma unroll
			for (int k = 0; k < kRunItems; k += 2) {
				const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(keys + i0 + k);
				mk[k] = (KeyT)(q.x >> 1);
				mk[k + 1] = (KeyT)(q.y >> 1);
			}
		}
		const uint4 v0 = *reinterpret_cast<const uint4*>(a.vals + i0), v1 = *reinterpret_cast<const uint4*>(a.vals + i0 + 4);
		sq[0] = v0.x; sq[1] = v0.y; sq[2] = v0.z; sq[3] = v0.w;
		sq[4] = v1.x; sq[5] = v1.y; sq[6] = v1.z; sq[7] = v1.w;
	} else {
#pragma unroll
		for (int k = 0; k < kRunItems; ++k) {
			const bool in = i0 + k < n;
			mk[k] = in ? (KeyT)(keys[i0 + k] >> 1) : (KeyT)0;
			sq[k] = in ? a.vals[i0 + k] : 0u;
		}
	}
	const bool memhash = a.mode == MEMS_MODE_MEMHASH;
	const bool grouped = a.n_seqs > 64;  // a batch of many problems: sequence bits are relative to the run's problem
	// forward pass: rec[k] / rec_at[k] = the run that closed right before entry k (length | kRunOk, first entry)
	uint32_t rec[kRunItems + 1], rec_at[kRunItems + 1];
	bool open = false, dup = false;
	uint32_t len = 0, at = 0, g0 = 0, longest = 0;
	uint64_t seen = 0;
	KeyT cur = 0;
	auto close = [&](uint32_t& r) {  // the open run ends: a hit if the policy takes it
		r = 0;
		if (len >= 2u) {
			longest = max(longest, len);
			bool ok = len <= kRunCap;
			if (memhash) ok = ok && !dup && (!a.seq_set || seen == a.seq_set);  // one occurrence per sequence; MaskedMemHash: exactly this set
			r = len | (ok ? kRunOk : 0u);
		}
	};
	const KeyT before = i0 > 0 && i0 < n ? (KeyT)(keys[i0 - 1] >> 1) : (KeyT)0;
	// the thread's leading entries that continue its predecessor's run: what the thread before needs to close that run
	bool leading = true, lead_dup = false;
	uint32_t lead_len = 0;
	uint64_t lead_seen = 0;
	const uint32_t lead_g0 = grouped && i0 < n ? a.meta[sq[0] >> a.pos_bits].group_first : 0u;
#pragma unroll
	for (int k = 0; k < kRunItems; ++k) {
		const bool in = i0 + k < n;
		const bool same = in && (k == 0 ? (i0 > 0 && mk[0] == before) : mk[k] == mk[k - 1]);
		rec[k] = 0;
		rec_at[k] = at;
		leading = leading && same;
		if (leading) {
			const uint64_t bit = 1ull << ((sq[k] >> a.pos_bits) - lead_g0);
			lead_dup |= (lead_seen & bit) != 0ull;
			lead_seen |= bit;
			++lead_len;
		}
		if (!same) {  // entry k starts a run (or lies past the end)
			if (open) close(rec[k]);
			open = in;
			if (in) {
				at = (uint32_t)(i0 + k);
				len = 1;
				dup = false;
				cur = mk[k];
				const uint32_t g = sq[k] >> a.pos_bits;
				g0 = grouped ? a.meta[g].group_first : 0u;
				seen = 1ull << (g - g0);
			}
		} else if (open) {  // (entries of a run that began before this thread's first entry are someone else's)
			const uint64_t bit = 1ull << ((sq[k] >> a.pos_bits) - g0);
			dup |= (seen & bit) != 0ull;
			seen |= bit;
			++len;
		}
	}
	rec[kRunItems] = 0;
	rec_at[kRunItems] = at;
	{
		// The run still open at the thread's last entry goes on in the next thread's leading entries: from the next lane's
		// registers.  Only if that lane's entries all belong to the run too (or the lane sits in another warp) is the
		// run followed through memory.
		const uint32_t next_len = __shfl_down_sync(0xffffffffu, lead_len, 1);
		const uint64_t next_seen = __shfl_down_sync(0xffffffffu, lead_seen, 1);
		const bool next_dup = __shfl_down_sync(0xffffffffu, (int)lead_dup, 1) != 0;
		if (open) {
			if (lane < 31 && next_len < (uint32_t)kRunItems) {
				dup |= next_dup || (seen & next_seen) != 0ull;
				seen |= next_seen;
				len += next_len;
			} else {
				uint64_t j = i0 + kRunItems;
				while (j < n && (KeyT)(keys[j] >> 1) == cur && len <= kRunCap) {
					const uint64_t bit = 1ull << ((a.vals[j] >> a.pos_bits) - g0);
					dup |= (seen & bit) != 0ull;
					seen |= bit;
					++len;
					++j;
				}
			}
			close(rec[kRunItems]);
		}
	}
	uint32_t mine = 0;
#pragma unroll
	for (int k = 1; k <= kRunItems; ++k) mine += rec[k] >> 31;
	// CTA-wide exclusive scan of the counts, then the chain
	uint32_t incl = mine;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += t;
	}
	longest = __reduce_max_sync(0xffffffffu, longest);
	if (lane == 31) s_tot[warp] = incl;
	if (lane == 0) s_max[warp] = longest;
	__syncthreads();
	uint32_t woff = 0, tot = 0, mx = 0;
#pragma unroll
	for (int w = 0; w < kScanBlock / 32; ++w) {
		const uint32_t t = s_tot[w];
		if (w < warp) woff += t;
		tot += t;
		mx = max(mx, s_max[w]);
	}
	if (tid == 0) {
		tile_hits[tile] = tot;
		if (mx > 1u) atomicMax(max_run, mx);  // one atomic per CTA
	}
	uint32_t out = tile * (uint32_t)(kRunTile / 2) + woff + incl - mine;
#pragma unroll
	for (int k = 1; k <= kRunItems; ++k) {
		if (rec[k] & kRunOk) {
			stage_start[out] = rec_at[k];
			stage_len[out] = (uint16_t)(rec[k] & 0xffffu);
			++out;
		}
	}
}

__global__ void __launch_bounds__(kScanBlock)
hit_gather_kernel(const uint32_t* __restrict__ stage_start, const uint16_t* __restrict__ stage_len, const uint32_t* __restrict__ tile_hits,
                  const uint32_t* __restrict__ tile_off, uint32_t* __restrict__ hit_start, uint16_t* __restrict__ hit_len) {
	const uint32_t tile = blockIdx.x, cnt = tile_hits[tile], off = tile_off[tile], from = tile * (uint32_t)(kRunTile / 2);
	for (uint32_t r = threadIdx.x; r < cnt; r += kScanBlock) {
		hit_start[off + r] = stage_start[from + r];
		hit_len[off + r] = stage_len[from + r];
	}
}

// ------------------------------------------------------------------------------------------------ members
// Members of a hit in ascending (sequence, position) order.  Inside a run the union is ordered by strand
// first (the strand is the key's low bit), then by (sequence, position): two sorted sublists to merge.
template <class KeyT>
struct MemberIter {
	const MatchArgs& a;
	uint32_t i0, m, i1, e;
	uint32_t v0, v1;  // heads of the two sublists (valid while i0 < m / i1 < e)
	__device__ MemberIter(const MatchArgs& args, uint32_t s, uint32_t len) : a(args), i0(s), e(s + len) {
		// first strand-1 entry: eight independent loads at a time instead of a chain of dependent ones
		m = e;
		for (uint32_t base = s; base < e && m == e; base += 8) {
			uint32_t mask = 0;
#pragma unroll
			for (int j = 0; j < 8; ++j)
				if (base + j < e) mask |= strand_of<KeyT>(a.keys, base + j) << j;
			if (mask) m = base + (uint32_t)__ffs((int)mask) - 1u;
		}
		i1 = m;
		v0 = i0 < m ? a.vals[i0] : 0u;
		v1 = i1 < e ? a.vals[i1] : 0u;
	}
	__device__ bool next(uint32_t& val, uint32_t& strand) {
		if (i0 >= m && i1 >= e) return false;
		const bool take0 = i1 >= e || (i0 < m && v0 < v1);
		if (take0) {
			val = v0;
			strand = 0;
			if (++i0 < m) v0 = a.vals[i0];
		} else {
			val = v1;
			strand = 1;
			if (++i1 < e) v1 = a.vals[i1];
		}
		return true;
	}
};

// The members of a hit of at most 8 entries in ascending (sequence, position) order, as registers: all loads are issued
// at once (MemberIter's merge is a chain of dependent loads) and a 19-step sorting network orders them.  Unused slots hold
// 0xffffffff and sort to the end.
__device__ __forceinline__ void order2(uint32_t& va, uint32_t& sa, uint32_t& vb, uint32_t& sb) {
	const bool swap = vb < va;
	const uint32_t v0 = swap ? vb : va, v1 = swap ? va : vb, s0 = swap ? sb : sa, s1 = swap ? sa : sb;
	va = v0;
	vb = v1;
	sa = s0;
	sb = s1;
}
template <class KeyT>
__device__ __forceinline__ void sorted_members8(const MatchArgs& a, uint32_t s, uint32_t len, uint32_t (&v)[8], uint32_t (&st)[8]) {
#pragma unroll
	for (int t = 0; t < 8; ++t) {
		const bool in = (uint32_t)t < len;
		v[t] = in ? a.vals[s + t] : 0xffffffffu;
		st[t] = in ? strand_of<KeyT>(a.keys, s + t) : 0u;
	}
#define MEMS_CE(i, j) order2(v[i], st[i], v[j], st[j])
	MEMS_CE(0, 1); MEMS_CE(2, 3); MEMS_CE(4, 5); MEMS_CE(6, 7);
	MEMS_CE(0, 2); MEMS_CE(1, 3); MEMS_CE(4, 6); MEMS_CE(5, 7);
	MEMS_CE(1, 2); MEMS_CE(5, 6); MEMS_CE(0, 4); MEMS_CE(3, 7);
	MEMS_CE(1, 5); MEMS_CE(2, 6);
	MEMS_CE(1, 4); MEMS_CE(3, 6);
	MEMS_CE(2, 4); MEMS_CE(3, 5);
	MEMS_CE(3, 4);
#undef MEMS_CE
}

__device__ __forceinline__ uint64_t mix64(uint64_t h, uint64_t v) {
	h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
	h *= 0xBF58476D1CE4E5B9ull;
	h ^= h >> 31;
	return h;
}

// ------------------------------------------------------------------------------------------------ 2. describe
constexpr int kMaxHitPasses = 8;

// A hit of at most 8 members as registers: its members in merged (sequence, position) order relative to the first.
// Two hits lie on one diagonal iff their signatures are equal.
struct HitSig {
	uint32_t len;      // members; kSigLong = more than 8: compared through memory (same_diagonal)
	uint32_t so[4];    // (sequence << 1 | orientation) of members 0..7, 16 bits each
	uint32_t diag[7];  // members 1..7: position minus / plus the first member's, by orientation
};
constexpr uint32_t kSigLong = 0xffffffffu;

__device__ __forceinline__ HitSig make_sig(const MatchArgs& a, uint32_t len, const uint32_t (&v)[8], const uint32_t (&st)[8]) {
	HitSig g;
	g.len = len;
#pragma unroll
	for (int t = 0; t < 4; ++t) g.so[t] = 0;
#pragma unroll
	for (int t = 0; t < 7; ++t) g.diag[t] = 0;
	const uint32_t sf = st[0], x0 = v[0] & a.pos_mask;
	g.so[0] = (v[0] >> a.pos_bits) << 1;
#pragma unroll
	for (int t = 1; t < 8; ++t) {
		if ((uint32_t)t < len) {
			const uint32_t o = st[t] ^ sf, p = v[t] & a.pos_mask;
			g.so[t >> 1] |= (((v[t] >> a.pos_bits) << 1) | o) << (16 * (t & 1));
			g.diag[t - 1] = o ? p + x0 : p - x0;  // (32-bit wrap is one-to-one here: |p - x0| < 2^30, p + x0 < 2^31)
		}
	}
	return g;
}
// The signature as hit_describe_kernel leaves it for segment_flag_kernel: 64 bytes per hit, in hit order.  There the
// members of a hit are neighbours in the union and cost one coalesced pass; gathered again in diagonal order (where
// neighbours sit at random places of the union) they were 1.6 GB of DRAM sectors per config-2 step.
struct __align__(16) HitSigRec {
	HitSig sig;
	uint32_t hit_start, members, pad[2];
};
static_assert(sizeof(HitSigRec) == 64, "one 64-byte record per hit");
__device__ __forceinline__ HitSigRec load_sig_rec(const HitSigRec* __restrict__ p) {
	const uint4* q = reinterpret_cast<const uint4*>(p);
	const uint4 a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3];
	HitSigRec r;
	r.sig.len = a0.x; r.sig.so[0] = a0.y; r.sig.so[1] = a0.z; r.sig.so[2] = a0.w;
	r.sig.so[3] = a1.x; r.sig.diag[0] = a1.y; r.sig.diag[1] = a1.z; r.sig.diag[2] = a1.w;
	r.sig.diag[3] = a2.x; r.sig.diag[4] = a2.y; r.sig.diag[5] = a2.z; r.sig.diag[6] = a2.w;
	r.hit_start = a3.x; r.members = a3.y; r.pad[0] = 0; r.pad[1] = 0;
	return r;
}
__device__ __forceinline__ void store_sig_rec(HitSigRec* __restrict__ p, const HitSig& g, uint32_t hit_start, uint32_t members) {
	uint4* q = reinterpret_cast<uint4*>(p);
	q[0] = make_uint4(g.len, g.so[0], g.so[1], g.so[2]);
	q[1] = make_uint4(g.so[3], g.diag[0], g.diag[1], g.diag[2]);
	q[2] = make_uint4(g.diag[3], g.diag[4], g.diag[5], g.diag[6]);
	q[3] = make_uint4(hit_start, members, 0u, 0u);
}

template <class KeyT>
__global__ void __launch_bounds__(256)
hit_describe_kernel(MatchArgs a, const uint32_t* __restrict__ hit_start, uint16_t* __restrict__ hit_len, uint32_t n_hits,
                    uint64_t* __restrict__ hkey, uint32_t* __restrict__ hid, uint32_t* __restrict__ hist, SortPlan plan,
                    HitSigRec* __restrict__ sig) {  // sig: optional
	__shared__ uint32_t s_hist[kMaxHitPasses * 256];
	for (int i = threadIdx.x; i < plan.n_passes * 256; i += blockDim.x) s_hist[i] = 0;
	__syncthreads();
	// grid-stride: a CTA describes many hits and flushes its histograms once (one flush per 256 hits would be
	// 8 same-address global atomics per hit)
	for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < n_hits; h += gridDim.x * blockDim.x) {
		const uint32_t s = hit_start[h];
		const uint32_t len = hit_len[h];
		uint32_t sf;
		int64_t x0;
		uint64_t hash;
		if (len <= 8) {  // the usual case: members as registers, no dependent loads
			uint32_t v[8], st8[8];
			sorted_members8<KeyT>(a, s, len, v, st8);
			if (sig) store_sig_rec(sig + h, make_sig(a, len, v, st8), s, len);
			sf = st8[0];
			x0 = v[0] & a.pos_mask;
			hash = mix64(0x1234567ull + len, v[0] >> a.pos_bits);
#pragma unroll
			for (int t = 1; t < 8; ++t) {
				if ((uint32_t)t < len) {
					const uint32_t o = st8[t] ^ sf;
					const int64_t p = v[t] & a.pos_mask;
					const int64_t diag = o ? p + x0 : p - x0;
					hash = mix64(hash, ((uint64_t)(v[t] >> a.pos_bits) << 1) | o);
					hash = mix64(hash, (uint64_t)diag);
				}
			}
		} else {
			if (sig) {
				HitSig g{};
				g.len = kSigLong;
				store_sig_rec(sig + h, g, s, len);
			}
			MemberIter<KeyT> it(a, s, len);
			uint32_t val, st;
			it.next(val, st);
			sf = st;
			x0 = val & a.pos_mask;
			hash = mix64(0x1234567ull + len, val >> a.pos_bits);
			while (it.next(val, st)) {
				const uint32_t o = st ^ sf;
				const int64_t p = val & a.pos_mask;
				const int64_t diag = o ? p + x0 : p - x0;
				hash = mix64(hash, ((uint64_t)(val >> a.pos_bits) << 1) | o);
				hash = mix64(hash, (uint64_t)diag);
			}
		}
		if (a.test_hash_bits) hash = (hash & ((1ull << a.test_hash_bits) - 1ull)) << (64 - a.test_hash_bits);
		// 56 key bits = seven sort passes: 56 - pos_bits hash bits are plenty (two diagonals that share them only cost
		// extra window tests and the de-dup of the components marked below)
		const uint64_t key = ((hash >> (64 - a.hit_key_bits + a.pos_bits)) << a.pos_bits) | (uint64_t)x0;
		hkey[h] = key;
		hid[h] = h;
		hit_len[h] = (uint16_t)(len | (sf ? kFirstStrandBit : 0));
#pragma unroll
		for (int q = 0; q < kMaxHitPasses; ++q) {
			if (q >= plan.n_passes) break;
			atomicAdd(&s_hist[q * 256 + ((uint32_t)(key >> plan.shift[q]) & ((1u << plan.bits[q]) - 1u))], 1u);
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < plan.n_passes * 256; i += blockDim.x) {
		uint32_t v = s_hist[i];
		if (v) atomicAdd(&hist[i], v);
	}
}

// described hits (their keys came with them): identity permutation + the digit histograms of the sort
__global__ void __launch_bounds__(256)
hit_histogram_kernel(const uint64_t* __restrict__ hkey, uint32_t n_hits, uint32_t* __restrict__ hid, uint32_t* __restrict__ hist,
                     SortPlan plan) {
	__shared__ uint32_t s_hist[kMaxHitPasses * 256];
	for (int i = threadIdx.x; i < plan.n_passes * 256; i += blockDim.x) s_hist[i] = 0;
	__syncthreads();
	for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < n_hits; h += gridDim.x * blockDim.x) {
		const uint64_t key = hkey[h];
		hid[h] = h;
#pragma unroll
		for (int q = 0; q < kMaxHitPasses; ++q) {
			if (q >= plan.n_passes) break;
			atomicAdd(&s_hist[q * 256 + ((uint32_t)(key >> plan.shift[q]) & ((1u << plan.bits[q]) - 1u))], 1u);
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < plan.n_passes * 256; i += blockDim.x) {
		uint32_t v = s_hist[i];
		if (v) atomicAdd(&hist[i], v);
	}
}

// ------------------------------------------------------------------------------------------------ 4. segments
// same diagonal: same member sequences, same relative orientations, same offsets to the first member
template <class KeyT>
__device__ bool same_diagonal(const MatchArgs& a, uint32_t sa, uint32_t la, uint32_t sb, uint32_t lb) {
	if (la != lb) return false;
	MemberIter<KeyT> ia(a, sa, la), ib(a, sb, lb);
	uint32_t va, ta, vb, tb;
	ia.next(va, ta);
	ib.next(vb, tb);
	if ((va >> a.pos_bits) != (vb >> a.pos_bits)) return false;
	const uint32_t fa = ta, fb = tb;
	const int64_t xa = va & a.pos_mask, xb = vb & a.pos_mask;
	while (ia.next(va, ta)) {
		ib.next(vb, tb);
		if ((va >> a.pos_bits) != (vb >> a.pos_bits)) return false;
		const uint32_t oa = ta ^ fa, ob = tb ^ fb;
		if (oa != ob) return false;
		const int64_t pa = va & a.pos_mask, pb = vb & a.pos_mask;
		if ((oa ? pa + xa : pa - xa) != (ob ? pb + xb : pb - xb)) return false;
	}
	return true;
}

constexpr uint8_t kFlagHead = 1;      // starts a new segment
constexpr uint8_t kFlagSameDiag = 2;  // same diagonal as the previous entry in sorted order

// A thread builds the signature of its own hit once (or reads the record hit_describe_kernel left) and takes its
// predecessor's from the neighbouring lane.
template <class KeyT>
__device__ __forceinline__ HitSig load_sig(const MatchArgs& a, uint32_t s, uint32_t len) {
	if (len > 8) {
		HitSig g{};
		g.len = kSigLong;
		return g;
	}
	uint32_t v[8], st[8];
	sorted_members8<KeyT>(a, s, len, v, st);
	return make_sig(a, len, v, st);
}

template <class KeyT>
__global__ void __launch_bounds__(256)
segment_flag_kernel(MatchArgs a, int L, const uint64_t* __restrict__ hkey, const uint32_t* __restrict__ hid,
                    const uint32_t* __restrict__ hit_start, const uint16_t* __restrict__ hit_len, uint32_t n_hits,
                    uint8_t* __restrict__ flags, uint32_t* __restrict__ is_head, uint32_t* __restrict__ collision_seen,
                    uint8_t* __restrict__ suspect, const HitSigRec* __restrict__ sig) {  // sig: optional (hit order)
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	const bool valid = i < n_hits;
	const uint32_t lane = threadIdx.x & 31;
	uint32_t s_mine = 0, l_mine = 0;
	HitSig mine;
	mine.len = 0;
#pragma unroll
	for (int t = 0; t < 4; ++t) mine.so[t] = 0;
#pragma unroll
	for (int t = 0; t < 7; ++t) mine.diag[t] = 0;
	if (valid) {
		const uint32_t h = hid[i];
		if (sig) {
			const HitSigRec r = load_sig_rec(sig + h);
			mine = r.sig;
			s_mine = r.hit_start;
			l_mine = r.members;
		} else {
			s_mine = hit_start[h];
			l_mine = hit_len[h] & ~kFirstStrandBit;
			mine = load_sig<KeyT>(a, s_mine, l_mine);
		}
	}
	// the predecessor's hit: from the lane below, or (first lane of a warp) gathered like the own one
	HitSig prev;
	prev.len = __shfl_up_sync(0xffffffffu, mine.len, 1);
#pragma unroll
	for (int t = 0; t < 4; ++t) prev.so[t] = __shfl_up_sync(0xffffffffu, mine.so[t], 1);
#pragma unroll
	for (int t = 0; t < 7; ++t) prev.diag[t] = __shfl_up_sync(0xffffffffu, mine.diag[t], 1);
	uint32_t s_prev = __shfl_up_sync(0xffffffffu, s_mine, 1), l_prev = __shfl_up_sync(0xffffffffu, l_mine, 1);
	if (!valid) return;
	uint8_t f = kFlagHead;
	if (i > 0) {
		const uint64_t k = hkey[i], kp = hkey[i - 1];
		if ((k >> a.pos_bits) == (kp >> a.pos_bits)) {
			if (lane == 0) {
				const uint32_t hb = hid[i - 1];
				if (sig) {
					const HitSigRec r = load_sig_rec(sig + hb);
					prev = r.sig;
					s_prev = r.hit_start;
					l_prev = r.members;
				} else {
					s_prev = hit_start[hb];
					l_prev = hit_len[hb] & ~kFirstStrandBit;
					prev = load_sig<KeyT>(a, s_prev, l_prev);
				}
			}
			bool same;
			if (mine.len == kSigLong || prev.len == kSigLong) {
				same = same_diagonal<KeyT>(a, s_mine, l_mine, s_prev, l_prev);
			} else {
				same = mine.len == prev.len;
#pragma unroll
				for (int t = 0; t < 4; ++t) same = same && mine.so[t] == prev.so[t];
#pragma unroll
				for (int t = 0; t < 7; ++t) same = same && mine.diag[t] == prev.diag[t];
			}
			if (same) {
				f = kFlagSameDiag;
				const uint64_t gap = (k & a.pos_mask) - (kp & a.pos_mask);
				if (gap > (uint64_t)L) f |= kFlagHead;
			} else {
				// two diagonals in one hash bucket (rare): the hits on either side of the foreign entry cannot see each
				// other, so the components they end up in may be reported twice — mark both for the de-dup
				atomicOr(collision_seen, 1u);
				suspect[i] = 1;
				suspect[i - 1] = 1;
			}
		}
	}
	flags[i] = f;
	is_head[i] = (f & kFlagHead) ? 1u : 0u;
}

__global__ void segment_compact_kernel(const uint32_t* __restrict__ is_head, const uint32_t* __restrict__ seg_of,
                                       const uint64_t* __restrict__ hkey, uint32_t pos_mask, uint32_t n_hits,
                                       uint32_t* __restrict__ seg_head, uint32_t* __restrict__ seg_x,
                                       uint32_t* __restrict__ seg_id) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n_hits && is_head[i]) {
		const uint32_t seg = seg_of[i];
		seg_head[seg] = i;
		seg_x[seg] = (uint32_t)hkey[i] & pos_mask;
		seg_id[seg] = seg;
	}
}

// ------------------------------------------------------------------------------------------------ 5. extend
// The window test of MatchFinder::ExtendMatch (MatchFinder.h:264-308) for the window k positions from a hit
// (k < 0: towards the first member's start): every member's seed position must be valid, all masked keys equal,
// and all (strand xor orientation) parities equal.  The reference recomputes two spaced-seed mers per member and
// window (GetSeedMer, ~84 ns each).  Here the test runs on the packed sequence itself, 32 windows per lane and
// 1024 per warp and probe:
//   * (canonical mer, strand) determines the forward mer and vice versa, so for a member in the first member's
//     orientation "equal key and parity" means "equal forward mers": the w cared bases agree;
//   * for a member in the opposite orientation it means "my reverse-complemented mer equals your forward mer, and
//     that mer is not its own reverse complement" (the two would then carry the same strand flag; possible for even
//     weights only) — with a palindromic pattern: the w cared bases of the two windows agree on opposite strands.
// So per probe a lane loads, for every member, the 64 bases under its 32 windows from the bit planes of the packed
// sequence (3 coalesced 64-bit loads; reverse members load the mirrored chunk and bit-reverse it), XORs them
// against the first member's and ORs the result into one disagreement bit per base; the windows then are
//   match(j) = AND over the w cared offsets o of agree(j + o)
// — w funnel shifts and ANDs for 32 windows.  Non-palindromic patterns (two entries of the reference's table) compare
// reverse members offset by offset instead (cared base i sits at another offset on the other strand), and the
// self-reverse-complement exception is w/2 shifted XORs of the first member's own planes.  Position validity
// collapses to one interval [kmin, kmax] per hit (intersection of the members' valid ranges).
constexpr int kExtendWarps = 4;
constexpr int kMemberTile = 64;  // members staged in shared memory per warp (MEMS_MAX_SEQS fits at once)
constexpr int kProbeWindows = 1024;  // windows one warp tests per probe
#ifndef MEMS_WARP_BUDGET
#define MEMS_WARP_BUDGET 6
#endif
constexpr int kWarpProbeBudget = MEMS_WARP_BUDGET;  // probes a single warp spends on one walk before deferring it to a whole CTA
// Most walks are short (config 2: 61 % end within 256 windows, 78 % within 512, 88 % need ONE 1024-window probe), so
// the walk kernels start every segment on a GROUP of kGroupLanes lanes — 32 windows per lane as before, several
// segments per warp — and only the walks that outlast kGroupProbeBudget group probes continue on the whole warp.
#ifndef MEMS_GROUP_LANES
#define MEMS_GROUP_LANES 8
#endif
#ifndef MEMS_GROUP_BUDGET
#define MEMS_GROUP_BUDGET 2
#endif
constexpr int kGroupLanes = MEMS_GROUP_LANES;
constexpr int kGroupsPerWarp = 32 / kGroupLanes;
constexpr int kGroupMembers = 16;  // members staged per segment (segment_stage_kernel); hits with more go to the whole warp
constexpr int kGroupProbeBudget = MEMS_GROUP_BUDGET;

struct SeedShape {  // what the window test needs of the pattern (SeedDesc::off / mirror), in shared memory
	uint8_t off[32], mirror[32];
	int w, L;
	bool palindromic;
};
__device__ __forceinline__ void load_seed_shape(SeedShape* sh, const SeedDesc& sd) {
	if (threadIdx.x < 32) {
		sh->off[threadIdx.x] = sd.off[threadIdx.x];
		sh->mirror[threadIdx.x] = sd.mirror[threadIdx.x];
	}
	if (threadIdx.x == 0) {
		sh->w = sd.w;
		sh->L = sd.L;
		sh->palindromic = sd.palindromic != 0;
	}
	__syncthreads();
}

struct Chunk {  // 64 consecutive bases as bit planes: bit t <-> base t of the chunk
	uint32_t h0, h1, l0, l1;
};
// the chunk that starts at base `at` of the batch's base array (at >= 0: the buffer carries a lead pad)
__device__ __forceinline__ Chunk load_chunk(const uint2* __restrict__ planes, int64_t at) {
	const uint2* p = planes + (at >> 5);
	const uint32_t sh = (uint32_t)at & 31u;
	const uint2 a = p[0], b = p[1], c = p[2];
	Chunk r;
	r.h0 = __funnelshift_r(a.x, b.x, sh);
	r.h1 = __funnelshift_r(b.x, c.x, sh);
	r.l0 = __funnelshift_r(a.y, b.y, sh);
	r.l1 = __funnelshift_r(b.y, c.y, sh);
	return r;
}
// bases in reverse order (the complement is left to the caller's XOR)
__device__ __forceinline__ Chunk reverse_chunk(const Chunk& c) {
	Chunk r;
	r.h0 = __brev(c.h1);
	r.h1 = __brev(c.h0);
	r.l0 = __brev(c.l1);
	r.l1 = __brev(c.l0);
	return r;
}

// What a walk needs of one member of a hit (union entry j; sf = strand of the hit's first member): the base index of
// its window 0 in the batch's base array (.x low 32 bits, .y high bits; bit 31 of .y = opposite orientation to the first
// member) and the window offsets [lo, hi] at which the member still lies inside its sequence.
template <class KeyT>
__device__ __forceinline__ uint2 walk_member_entry(const MatchArgs& a, uint32_t j, uint32_t sf, int32_t& lo, int32_t& hi) {
	const uint32_t val = a.vals[j];
	const uint32_t o = strand_of<KeyT>(a.keys, j) ^ sf;
	const SeqMeta m = a.meta[val >> a.pos_bits];
	const int32_t p = (int32_t)(val & a.pos_mask), last = (int32_t)m.n_seeds - 1;
	lo = o ? p - last : -p;
	hi = o ? p : last - p;
	const uint64_t base = m.word_off * 16ull + (uint64_t)p;
	return make_uint2((uint32_t)base, (uint32_t)(base >> 32) | (o ? 0x80000000u : 0u));
}

// GW lanes (a whole warp, or an aligned group of 4/8/16 lanes) walk one hit: every sync primitive below is scoped to the
// group's lanes, so the groups of a warp run their own control flow.
template <class KeyT, int GW = 32>
struct WalkHit {
	static constexpr int kWindows = 32 * GW;                             // windows per probe
	static constexpr int kTile = kMemberTile;  // members staged at once (groups only ever see hits of <= kGroupMembers)
	const MatchArgs& a;
	const SeedShape& shape;
	uint2* s_mem;  // this group's kTile slots: base index of the member's window 0 in the batch's base array
	               // (.x low 32 bits, .y high bits), bit 31 of .y = opposite orientation to the first member
	uint32_t s, len, sf;
	int32_t kmin, kmax;  // valid window offsets (all |k| < 2^31: positions are < 2^30)
	bool any_reverse;
	int lane;  // within the group
	uint32_t gmask = 0xffffffffu;  // the group's lanes in the warp
	int gshift = 0;                // its first lane
#ifdef MEMS_WALK_STATS
	uint32_t n_probes = 0;
#endif

	__device__ __forceinline__ uint32_t ballot(bool p) const {
		return GW == 32 ? __ballot_sync(0xffffffffu, p) : (__ballot_sync(gmask, p) & gmask) >> gshift;
	}
	__device__ __forceinline__ uint32_t shfl(uint32_t x, int src) const { return __shfl_sync(gmask, x, src, GW); }
	__device__ __forceinline__ void sync() const { __syncwarp(gmask); }
	__device__ __forceinline__ int32_t group_max(int32_t x) const {
		if constexpr (GW == 32) {
			return __reduce_max_sync(0xffffffffu, x);
		} else {
#pragma unroll
			for (int o = GW / 2; o > 0; o >>= 1) x = max(x, __shfl_xor_sync(gmask, x, o, GW));
			return x;
		}
	}
	__device__ __forceinline__ int32_t group_min(int32_t x) const {
		if constexpr (GW == 32) {
			return __reduce_min_sync(0xffffffffu, x);
		} else {
#pragma unroll
			for (int o = GW / 2; o > 0; o >>= 1) x = min(x, __shfl_xor_sync(gmask, x, o, GW));
			return x;
		}
	}

	__device__ uint2 member_entry(uint32_t j, int32_t& lo, int32_t& hi) const { return walk_member_entry<KeyT>(a, j, sf, lo, hi); }
	__device__ void load_tile(uint32_t first) {
		for (uint32_t t = lane; t < kTile && first + t < len; t += GW) {
			int32_t lo, hi;
			s_mem[t] = member_entry(s + first + t, lo, hi);
		}
		sync();
	}
	__device__ void init(uint32_t hit_s, uint32_t hit_len, uint32_t first_strand) {
		s = hit_s;
		len = hit_len;
		sf = first_strand;
		int32_t lo = INT32_MIN, hi = INT32_MAX;
		bool rev = false;
		for (uint32_t t = lane; t < len; t += GW) {  // one pass: valid range and the first member tile together
			int32_t l2, h2;
			const uint2 e = member_entry(s + t, l2, h2);
			if (t < kTile) s_mem[t] = e;
			rev |= (e.y >> 31) != 0u;
			lo = max(lo, l2);
			hi = min(hi, h2);
		}
		kmin = group_max(lo);
		kmax = group_min(hi);
		any_reverse = ballot(rev) != 0u;
		sync();
	}
	// A segment staged by segment_stage_kernel: members, valid range and orientations are already known.
	__device__ void adopt(const uint2* __restrict__ staged, uint32_t hit_s, uint32_t hit_len, uint32_t first_strand, int32_t lo,
	                      int32_t hi, bool rev, bool members_in_place) {
		s = hit_s;
		len = hit_len;
		sf = first_strand;
		kmin = lo;
		kmax = hi;
		any_reverse = rev;
		if (!members_in_place) {
			for (uint32_t t = lane; t < len; t += GW) s_mem[t] = staged[t];
			sync();
		}
	}
	// Probe the kWindows windows at distances 1.. from k0 in direction dir (+1/-1): this lane's 32 windows are
	// the distances 32*lane + 1 .. 32*lane + 32, bit j of the result = distance 32*lane + j + 1 matches.
	__device__ uint32_t probe(int32_t k0, int dir) {
#ifdef MEMS_WALK_STATS
		++n_probes;
#endif
		const int L = shape.L, w = shape.w;
		// the lane's windows in ascending window order: klo + j, j = 0..31
		const int32_t klo = dir > 0 ? k0 + 1 + 32 * lane : k0 - 32 * (lane + 1);
		uint32_t m = 0;  // valid windows
		{
			const int32_t j0 = max(kmin - klo, 0), j1 = min(kmax - klo, 31);
			if (j0 <= j1) m = (0xffffffffu >> (31 - j1)) & (0xffffffffu << j0);
		}
		const bool live = m != 0u;  // lanes without a valid window load nothing (their addresses may lie outside the buffer)
		if (!ballot(live)) return 0u;
		// Every member is read in the orientation of the hit's first member: members of the other orientation load the
		// mirrored chunk and reverse it (the complement is folded into the XOR).  The reference the others are compared
		// with is simply the first union entry, whichever orientation it has; cared base i of a window sits at offset
		// off[i] for members of the first orientation and at mirror[i] for the others (the same for palindromic patterns).
		const uint2 e0 = s_mem[0];
		const bool ref_rev = (e0.y >> 31) != 0u;
		const uint8_t* ref_off = ref_rev ? shape.mirror : shape.off;
		const uint8_t* oth_off = ref_rev ? shape.off : shape.mirror;
		Chunk ref{0, 0, 0, 0};
		if (live) {
			const int64_t base0 = (int64_t)(((uint64_t)(e0.y & 0x7fffffffu) << 32) | e0.x);
			ref = ref_rev ? reverse_chunk(load_chunk(a.planes, base0 - klo + (L - 1) - 63)) : load_chunk(a.planes, base0 + klo);
		}
		uint32_t d0 = 0, d1 = 0;  // disagreement per base of the chunk, all members
		for (uint32_t first = 0; first < len; first += kTile) {
			if (first) load_tile(first);
			const uint32_t cnt = len - first < kTile ? len - first : kTile;
			for (uint32_t t = first ? 0u : 1u; t < cnt; ++t) {
				const uint2 e = s_mem[t];
				const int64_t base = (int64_t)(((uint64_t)(e.y & 0x7fffffffu) << 32) | e.x);
				const bool rev = (e.y >> 31) != 0u;  // warp-uniform
				if (!live) continue;
				// window klo + j of a reverse member starts at base - (klo + j) and is read on the other strand: base t of
				// the reversed chunk that ENDS at base - klo + L - 1 is (the complement of) what window j holds at offset t - j
				const Chunk c = rev ? reverse_chunk(load_chunk(a.planes, base - klo + (L - 1) - 63)) : load_chunk(a.planes, base + klo);
				if (rev == ref_rev) {
					d0 |= (c.h0 ^ ref.h0) | (c.l0 ^ ref.l0);
					d1 |= (c.h1 ^ ref.h1) | (c.l1 ^ ref.l1);
				} else if (shape.palindromic) {
					d0 |= ~(c.h0 ^ ref.h0) | ~(c.l0 ^ ref.l0);
					d1 |= ~(c.h1 ^ ref.h1) | ~(c.l1 ^ ref.l1);
				} else {
					for (int i = 0; i < w; ++i) {
						const uint32_t o = ref_off[i], q = oth_off[i];
						const uint32_t x = ~(__funnelshift_r(c.h0, c.h1, q) ^ __funnelshift_r(ref.h0, ref.h1, o)) |
						                   ~(__funnelshift_r(c.l0, c.l1, q) ^ __funnelshift_r(ref.l0, ref.l1, o));
						m &= ~x;
					}
				}
			}
			if (len > kTile) sync();
		}
		if (len > kTile) load_tile(0);
		const uint32_t ok0 = ~d0, ok1 = ~d1;
		for (int i = 0; i < w; ++i) m &= __funnelshift_r(ok0, ok1, ref_off[i]);
		if (any_reverse && !(w & 1)) {
			// a mer equal to its own reverse complement carries the same strand flag on both strands: never a match
			// between opposite orientations.  Cared base i must be the complement of cared base w-1-i.
			uint32_t self = 0xffffffffu;
			for (int i = 0; i < w / 2; ++i) {
				const uint32_t o = ref_off[i], q = ref_off[w - 1 - i];
				self &= (__funnelshift_r(ref.h0, ref.h1, o) ^ __funnelshift_r(ref.h0, ref.h1, q)) &
				        (__funnelshift_r(ref.l0, ref.l1, o) ^ __funnelshift_r(ref.l0, ref.l1, q));
			}
			m &= ~self;
		}
		return dir > 0 ? m : __brev(m);
	}
	// distance (1-based) of the highest / lowest set bit over the warp's words, 0 if none
	__device__ int highest_set(uint32_t m) const {
		const uint32_t b = ballot(m != 0u);
		if (!b) return 0;
		const int z = 31 - __clz((int)b);
		return 32 * z + 32 - __clz((int)shfl(m, z));
	}
	__device__ int lowest_set(uint32_t m) const {
		const uint32_t b = ballot(m != 0u);
		if (!b) return 0;
		const int z = __ffs((int)b) - 1;
		return 32 * z + __ffs((int)shfl(m, z));
	}
	// Where a chain of matches at most L apart breaks inside one probe.  The chain starts at distance `from`
	// (0 = the window the walk stands on, else a set bit of m); matches before `from` are ignored.  A distance d is
	// "open" when none of the L windows before it (d-L .. d-1) matches; the chain's last match x is followed by L
	// mismatches, so the first open distance after `from` is x + L + 1 (whether or not that window itself matches).
	// Returns it (<= kWindows: the gap then lies completely inside the probed windows), or 0 when the chain runs
	// on to the end of the probe.  Lane-parallel: a lane smears its word and its predecessor's by 1..L bits.
	__device__ int chain_break(uint32_t m, int from) const {
		const int L = shape.L;
		// matches before `from` do not count; distance 0 sits in bit 31 of lane 0's predecessor word
		uint32_t cur = m;
		if (from > 0) {
			const int fl = (from - 1) >> 5, fb = (from - 1) & 31;
			if (lane < fl) cur = 0u;
			else if (lane == fl) cur &= 0xffffffffu << fb;
		}
		uint32_t prev = __shfl_up_sync(gmask, cur, 1, GW);
		if (lane == 0) prev = from == 0 ? 0x80000000u : 0u;
		// reach = bits of OR_{s = 1..L} ((cur:prev) << s) that fall into cur's word: shift by one, double, one more step
		uint32_t hi = __funnelshift_l(prev, cur, 1), lo = prev << 1;
		int have = 1;  // shifts 1 .. have are in
		while (2 * have <= L) {
			hi |= __funnelshift_l(lo, hi, have);
			lo |= lo << have;
			have *= 2;
		}
		if (have < L) hi |= __funnelshift_l(lo, hi, L - have);
		uint32_t open = ~hi;  // distances whose L predecessors hold no match ...
		{                      // ... after `from`
			const int fl = from >> 5, fb = from & 31;
			if (lane < fl) open = 0u;
			else if (lane == fl) open &= 0xffffffffu << fb;
		}
		return lowest_set(open);
	}
	// Walk from window k0 in direction dir over matching windows that start <= L apart, as far as they go
	// (the closure loop of MatchFinder::ExtendMatch, MatchFinder.h:259-355).  If stop_dist >= 0 the walk ends as
	// soon as it stands within L of that distance (the next segment of the diagonal) and *linked is set.
	// Returns the distance walked.  After max_probes probes the walk gives up with *exhausted set: such
	// walks (a handful per genome set, but up to hundreds of kbp long) are finished by whole CTAs, see cta_walk.
	__device__ int32_t walk(int32_t k0, int dir, int32_t stop_dist, bool* linked, int max_probes, bool* exhausted) {
		const int L = shape.L;
		int32_t walked = 0;
		*linked = false;
		*exhausted = false;
		for (int n = 0;; ++n) {
			if (stop_dist >= 0 && stop_dist <= walked + L) {
				*linked = true;
				return walked;
			}
			if (n == max_probes) {
				*exhausted = true;
				return walked;
			}
			uint32_t m = probe(k0 + dir * walked, dir);
			if (stop_dist >= 0 && stop_dist - walked <= kWindows) {
				// the next segment's first hit is a matching window of this diagonal: nothing beyond it matters
				const int t = stop_dist - walked - 1, tl = t >> 5, tb = t & 31;
				if (lane > tl) m = 0u;
				else if (lane == tl) m = (m & (0xffffffffu >> (31 - tb))) | (1u << tb);
			}
			const int open_at = chain_break(m, 0);
			if (open_at) {
				walked += open_at - L - 1;
				if (stop_dist >= 0 && stop_dist <= walked + L) *linked = true;
				return walked;
			}
			walked += highest_set(m);  // chain_break saw no gap: there is a match, and the chain goes on past the probe
		}
	}
};
template <class KeyT>
using WarpHit = WalkHit<KeyT, 32>;

#ifdef MEMS_WALK_STATS
__device__ unsigned long long g_walk_stats[48];  // [0] probes total, [1] max per walk, [2+i] walks with 2^i probes
#endif

struct SegView {
	const uint64_t* hkey;
	const uint32_t* hid;
	const uint32_t* hit_start;
	const uint16_t* hit_len;
	const uint8_t* flags;
	const uint32_t* seg_head;
	uint32_t n_hits, n_seg;
	const uint32_t* order;  // segment ids sorted by the first-member position of their first hit
	const uint8_t* suspect;  // per sorted hit: next to a foreign diagonal of the same hash bucket
};

// Right walk of every segment: from its last hit, follow matching windows until either the next segment of
// the same diagonal is within reach (link = 1: the two are connected, its own walk continues from there) or
// no window within L matches (link = 0: reach = last matching window, the component's right end).
// A warp takes kGroupsPerWarp consecutive segments of the walk order: first every group of kGroupLanes lanes walks its
// own segment (kGroupProbeBudget probes of 32 x kGroupLanes windows), then the whole warp finishes, one after the other,
// the walks that are still going (and the hits with more members than a group stages).
struct __align__(16) SegDesc {  // one segment as the walk kernels see it, staged in walk order by segment_stage_kernel
	uint32_t seg, x0;    // x0: first-member position of the segment's first hit
	uint32_t hit_start;  // union index of the first hit's members
	uint32_t info;       // member count | kSeg* flags
	int32_t c, next_at;  // last hit / first hit of the diagonal's next segment, as distances from x0
	int32_t kmin, kmax;  // valid window offsets (staged segments)
};
constexpr uint32_t kSegLenMask = 0xffffu, kSegStrand = 1u << 16, kSegHasNext = 1u << 17, kSegFirstOfDiagonal = 1u << 18,
                   kSegAnyReverse = 1u << 19, kSegStaged = 1u << 20;

// Everything a walk needs of its segment sits at the end of a chain of dependent loads (walk order -> segment -> hit ->
// union entries -> sequence table).  Followed by the walking warps themselves that chain is most of a short walk's
// time; here 8 threads per segment follow it once, with the whole grid's parallelism to hide it, and leave a 32-byte
// descriptor and the member entries (hits of <= kGroupMembers members) at the segment's slot of the walk order.
template <class KeyT>
__global__ void __launch_bounds__(256)
segment_stage_kernel(MatchArgs a, SegView v, SegDesc* __restrict__ desc, uint2* __restrict__ members, uint32_t* __restrict__ slot_of_seg) {
	const int sub = threadIdx.x & 7;
	const uint32_t slot = (blockIdx.x * 256u + threadIdx.x) >> 3;
	if (slot >= v.n_seg) return;  // whole groups leave
	const uint32_t gmask = 0xffu << ((threadIdx.x & 31) & ~7);
	const uint32_t seg = v.order[slot];  // segments are visited in order of genome position (L2 locality)
	const uint32_t hi = v.seg_head[seg];
	const uint32_t end = seg + 1 < v.n_seg ? v.seg_head[seg + 1] : v.n_hits;
	const uint32_t h = v.hid[hi];
	const int64_t x0 = (int64_t)(v.hkey[hi] & a.pos_mask);
	const uint32_t hit_start = v.hit_start[h], hl = v.hit_len[h];
	const uint32_t len = hl & ~kFirstStrandBit, strand = (hl & kFirstStrandBit) ? 1u : 0u;
	int32_t lo = INT32_MIN, hi_k = INT32_MAX;
	bool rev = false;
	const bool staged = len <= (uint32_t)kGroupMembers;
	if (staged) {
		for (uint32_t t = sub; t < len; t += 8) {
			int32_t l2, h2;
			const uint2 e = walk_member_entry<KeyT>(a, hit_start + t, strand, l2, h2);
			members[(size_t)slot * kGroupMembers + t] = e;
			rev |= (e.y >> 31) != 0u;
			lo = max(lo, l2);
			hi_k = min(hi_k, h2);
		}
#pragma unroll
		for (int o = 4; o > 0; o >>= 1) {
			lo = max(lo, __shfl_xor_sync(gmask, lo, o, 8));
			hi_k = min(hi_k, __shfl_xor_sync(gmask, hi_k, o, 8));
		}
		rev = (__ballot_sync(gmask, rev) & gmask) != 0u;
	}
	if (sub == 0) {
		const bool has_next = seg + 1 < v.n_seg && (v.flags[end] & kFlagSameDiag);
		SegDesc d;
		d.seg = seg;
		d.x0 = (uint32_t)x0;
		d.hit_start = hit_start;
		d.info = len | (strand ? kSegStrand : 0u) | (has_next ? kSegHasNext : 0u) | ((v.flags[hi] & kFlagSameDiag) ? 0u : kSegFirstOfDiagonal) |
		         (rev ? kSegAnyReverse : 0u) | (staged ? kSegStaged : 0u);
		d.c = (int32_t)((int64_t)(v.hkey[end - 1] & a.pos_mask) - x0);
		d.next_at = has_next ? (int32_t)((int64_t)(v.hkey[end] & a.pos_mask) - x0) : 0;
		d.kmin = lo;
		d.kmax = hi_k;
		desc[slot] = d;
		slot_of_seg[seg] = slot;
	}
}

__device__ __forceinline__ SegDesc load_seg_desc(const SegDesc* __restrict__ p) {
	const uint4 q0 = reinterpret_cast<const uint4*>(p)[0], q1 = reinterpret_cast<const uint4*>(p)[1];
	SegDesc d;
	d.seg = q0.x; d.x0 = q0.y; d.hit_start = q0.z; d.info = q0.w;
	d.c = (int32_t)q1.x; d.next_at = (int32_t)q1.y; d.kmin = (int32_t)q1.z; d.kmax = (int32_t)q1.w;
	return d;
}
__device__ __forceinline__ SegDesc broadcast_seg_desc(const SegDesc& w, int src) {
	SegDesc r;
	r.seg = __shfl_sync(0xffffffffu, w.seg, src);
	r.x0 = __shfl_sync(0xffffffffu, w.x0, src);
	r.hit_start = __shfl_sync(0xffffffffu, w.hit_start, src);
	r.info = __shfl_sync(0xffffffffu, w.info, src);
	r.c = __shfl_sync(0xffffffffu, w.c, src);
	r.next_at = __shfl_sync(0xffffffffu, w.next_at, src);
	r.kmin = __shfl_sync(0xffffffffu, w.kmin, src);
	r.kmax = __shfl_sync(0xffffffffu, w.kmax, src);
	return r;
}
// the whole warp takes over a segment one of its groups staged in shared memory (or sets a long hit up from scratch)
template <class KeyT>
__device__ __forceinline__ void warp_adopt(WarpHit<KeyT>& w, const SegDesc& d, uint2* group_mem, uint2* wide_mem) {
	const uint32_t len = d.info & kSegLenMask;
	if (d.info & kSegStaged) {
		w.s_mem = group_mem;
		w.adopt(nullptr, d.hit_start, len, (d.info & kSegStrand) ? 1u : 0u, d.kmin, d.kmax, (d.info & kSegAnyReverse) != 0u, true);
	} else {
		w.s_mem = wide_mem;
		w.init(d.hit_start, len, (d.info & kSegStrand) ? 1u : 0u);
	}
}
constexpr uint32_t kNeedRight = 1u, kNeedLeft = 2u;

template <class KeyT>
__global__ void __launch_bounds__(kExtendWarps * 32)
walk_right_kernel(MatchArgs a, const __grid_constant__ SeedDesc sd, SegView v, const SegDesc* __restrict__ desc,
                  const uint2* __restrict__ members, uint32_t* __restrict__ seg_link,
                  uint32_t* __restrict__ seg_reach, uint2* __restrict__ defer, uint32_t* __restrict__ defer_count,
                  uint32_t* __restrict__ seg_left, uint8_t* __restrict__ seg_left_state, uint2* __restrict__ defer_left,
                  uint32_t* __restrict__ defer_left_count) {
	__shared__ uint2 s_mem[kExtendWarps][kGroupsPerWarp * kGroupMembers], s_wide[kExtendWarps][kMemberTile];
	__shared__ SeedShape s_shape;
	load_seed_shape(&s_shape, sd);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / kGroupLanes, gl = lane % kGroupLanes;
	const uint32_t slot = (blockIdx.x * kExtendWarps + warp) * kGroupsPerWarp + grp;
	if (slot - grp >= v.n_seg) return;  // warp-uniform
	SegDesc d{};
	uint32_t need = 0;
	int32_t cl = 0;  // left walk so far
	if (slot < v.n_seg) {
		d = load_seg_desc(desc + slot);
		need = kNeedRight | ((d.info & kSegFirstOfDiagonal) ? kNeedLeft : 0u);
		if (d.info & kSegStaged) {
			WalkHit<KeyT, kGroupLanes> g{a, s_shape, s_mem[warp] + grp * kGroupMembers, 0, 0, 0, 0, 0, false, gl,
			                             kGroupLanes == 32 ? 0xffffffffu : ((1u << kGroupLanes) - 1u) << (grp * kGroupLanes), grp * kGroupLanes};
			g.adopt(members + (size_t)slot * kGroupMembers, d.hit_start, d.info & kSegLenMask, (d.info & kSegStrand) ? 1u : 0u, d.kmin,
			        d.kmax, (d.info & kSegAnyReverse) != 0u, false);
			bool linked, exhausted;
#ifdef MEMS_WALK_STATS
			const int32_t c_start = d.c;
#endif
			d.c += g.walk(d.c, +1, (d.info & kSegHasNext) ? d.next_at - d.c : -1, &linked, kGroupProbeBudget, &exhausted);
#ifdef MEMS_WALK_STATS
			if (gl == 0 && !exhausted) atomicAdd(&g_walk_stats[16 + (linked ? 0 : 16) + min(31 - __clz((d.c - c_start) | 1), 15)], 1ull);
#endif
			if (!exhausted) {
				need &= ~kNeedRight;
				if (gl == 0) {
					seg_link[d.seg] = linked ? 1u : 0u;
					seg_reach[d.seg] = d.x0 + (uint32_t)d.c;
				}
			}
			// The first segment of a DIAGONAL certainly starts a component (nothing on its diagonal can link to it), and its
			// members are already staged: walk left right away instead of setting all of this up again in walk_left_kernel.
			if (need & kNeedLeft) {
				bool l2, ex2;
				cl = -g.walk(0, -1, -1, &l2, kGroupProbeBudget, &ex2);
				if (!ex2) {
					need &= ~kNeedLeft;
					if (gl == 0) {
						seg_left[d.seg] = d.x0 + (uint32_t)cl;
						seg_left_state[d.seg] = 1;
					}
				}
			}
#ifdef MEMS_WALK_STATS
			if (gl == 0) {
				atomicAdd(&g_walk_stats[0], (unsigned long long)g.n_probes);
				atomicAdd(&g_walk_stats[2 + min(31 - __clz(g.n_probes | 1), 5)], 1ull);
				if (need & kNeedRight) atomicAdd(&g_walk_stats[8], 1ull);
				if (need & kNeedLeft) atomicAdd(&g_walk_stats[9], 1ull);
			}
#endif
		}
	}
	__syncwarp();
	// ---- the walks that go on: the whole warp, one segment at a time
	uint32_t todo = __ballot_sync(0xffffffffu, need != 0u && gl == 0);
	while (todo) {
		const int src = __ffs((int)todo) - 1;
		todo &= todo - 1u;
		SegDesc b = broadcast_seg_desc(d, src);
		const uint32_t b_need = __shfl_sync(0xffffffffu, need, src);
		int32_t b_cl = __shfl_sync(0xffffffffu, cl, src);
		WarpHit<KeyT> w{a, s_shape, nullptr, 0, 0, 0, 0, 0, false, lane};
		warp_adopt<KeyT>(w, b, s_mem[warp] + (src / kGroupLanes) * kGroupMembers, s_wide[warp]);
		if (b_need & kNeedRight) {
			bool linked, exhausted;
			b.c += w.walk(b.c, +1, (b.info & kSegHasNext) ? b.next_at - b.c : -1, &linked, a.warp_budget, &exhausted);
			if (lane == 0) {
				if (exhausted) {  // hand the rest of this walk to a whole CTA (long_walk_kernel)
					defer[atomicAdd(defer_count, 1u)] = make_uint2(b.seg, (uint32_t)b.c);
				} else {
					seg_link[b.seg] = linked ? 1u : 0u;
					seg_reach[b.seg] = b.x0 + (uint32_t)b.c;
				}
			}
		}
		if (b_need & kNeedLeft) {
			bool l2, ex2;
			b_cl -= w.walk(b_cl, -1, -1, &l2, a.warp_budget, &ex2);
			if (lane == 0) {
				if (ex2) {
					defer_left[atomicAdd(defer_left_count, 1u)] = make_uint2(b.seg, (uint32_t)b_cl);
					seg_left_state[b.seg] = 2;  // handed to the CTA-wide walker
				} else {
					seg_left[b.seg] = b.x0 + (uint32_t)b_cl;
					seg_left_state[b.seg] = 1;
				}
			}
		}
		__syncwarp();  // s_wide is rewritten by the next segment
	}
}

// a segment starts a component unless its predecessor linked to it
__global__ void chain_first_kernel(const uint32_t* __restrict__ seg_link, uint32_t n_seg, uint32_t* __restrict__ first) {
	const uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
	if (seg < n_seg) first[seg] = (seg == 0 || !seg_link[seg - 1]) ? 1u : 0u;
}

// Both ends of every component into the component arrays, one thread per segment; the first segments that still need
// their left walk (walk_right_kernel did it for the first segment of every diagonal) go on a list.
__global__ void finish_segments_kernel(SegView v, const uint32_t* __restrict__ seg_link, const uint32_t* __restrict__ seg_reach,
                                       const uint32_t* __restrict__ first, const uint32_t* __restrict__ first_excl,
                                       const uint32_t* __restrict__ seg_left, const uint8_t* __restrict__ seg_left_state,
                                       uint32_t* __restrict__ comp_rep, uint32_t* __restrict__ comp_left,
                                       uint32_t* __restrict__ comp_right, uint8_t* __restrict__ comp_suspect,
                                       const uint32_t* __restrict__ slot_of_seg, uint32_t* __restrict__ todo,
                                       uint32_t* __restrict__ todo_count) {
	const uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
	if (seg >= v.n_seg) return;
	const bool is_first = first[seg] != 0;
	const uint32_t comp = first_excl[seg] - (is_first ? 0u : 1u);
	if (!seg_link[seg]) comp_right[comp] = seg_reach[seg];
	const uint32_t e0 = v.seg_head[seg], e1 = (seg + 1 < v.n_seg ? v.seg_head[seg + 1] : v.n_hits) - 1;
	if (v.suspect[e0] | v.suspect[e1]) comp_suspect[comp] = 1;  // only ever set: racing writers agree
	if (!is_first) return;
	comp_rep[comp] = e0;
	const uint8_t done = seg_left_state[seg];
	if (done == 1) comp_left[comp] = seg_left[seg];
	else if (done == 0) todo[atomicAdd(todo_count, 1u)] = slot_of_seg[seg];  // (2: already with the CTA-wide walker)
}

// Left walk of the listed first segments (slots of the walk order; groups first, then the whole warp, as in
// walk_right_kernel).
template <class KeyT>
__global__ void __launch_bounds__(kExtendWarps * 32)
walk_left_kernel(MatchArgs a, const __grid_constant__ SeedDesc sd, const SegDesc* __restrict__ desc,
                 const uint2* __restrict__ members, const uint32_t* __restrict__ todo_list,
                 const uint32_t* __restrict__ todo_count, const uint32_t* __restrict__ first_excl,
                 uint32_t* __restrict__ comp_left, uint2* __restrict__ defer, uint32_t* __restrict__ defer_count) {
	__shared__ uint2 s_mem[kExtendWarps][kGroupsPerWarp * kGroupMembers], s_wide[kExtendWarps][kMemberTile];
	__shared__ SeedShape s_shape;
	load_seed_shape(&s_shape, sd);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / kGroupLanes, gl = lane % kGroupLanes;
	const uint32_t item = (blockIdx.x * kExtendWarps + warp) * kGroupsPerWarp + grp;
	const uint32_t n_todo = *todo_count;
	if (item - grp >= n_todo) return;  // warp-uniform
	SegDesc d{};
	uint32_t need = 0;
	int32_t cl = 0;
	if (item < n_todo) {
		const uint32_t slot = todo_list[item];
		d = load_seg_desc(desc + slot);
		need = 1;
		if (d.info & kSegStaged) {
			WalkHit<KeyT, kGroupLanes> g{a, s_shape, s_mem[warp] + grp * kGroupMembers, 0, 0, 0, 0, 0, false, gl,
			                             kGroupLanes == 32 ? 0xffffffffu : ((1u << kGroupLanes) - 1u) << (grp * kGroupLanes), grp * kGroupLanes};
			g.adopt(members + (size_t)slot * kGroupMembers, d.hit_start, d.info & kSegLenMask, (d.info & kSegStrand) ? 1u : 0u, d.kmin,
			        d.kmax, (d.info & kSegAnyReverse) != 0u, false);
			bool linked, exhausted;
			cl = -g.walk(0, -1, -1, &linked, kGroupProbeBudget, &exhausted);
			if (!exhausted) {
				need = 0;
				if (gl == 0) comp_left[first_excl[d.seg]] = d.x0 + (uint32_t)cl;
			}
		}
	}
	__syncwarp();
	uint32_t todo = __ballot_sync(0xffffffffu, need != 0u && gl == 0);
	while (todo) {
		const int src = __ffs((int)todo) - 1;
		todo &= todo - 1u;
		const SegDesc b = broadcast_seg_desc(d, src);
		int32_t b_cl = __shfl_sync(0xffffffffu, cl, src);
		WarpHit<KeyT> w{a, s_shape, nullptr, 0, 0, 0, 0, 0, false, lane};
		warp_adopt<KeyT>(w, b, s_mem[warp] + (src / kGroupLanes) * kGroupMembers, s_wide[warp]);
		bool linked, exhausted;
		b_cl -= w.walk(b_cl, -1, -1, &linked, a.warp_budget, &exhausted);
		if (lane == 0) {
			if (exhausted) defer[atomicAdd(defer_count, 1u)] = make_uint2(b.seg, (uint32_t)b_cl);
			else comp_left[first_excl[b.seg]] = b.x0 + (uint32_t)b_cl;
		}
		__syncwarp();
	}
}

// ---- long walks -------------------------------------------------------------------------------------
// A diagonal shared by few sequences has few hits but can match for hundreds of kbp (the sequences agree
// wherever the others carry a SNP), so a handful of walks are 10^2-10^4 probes long: as one warp each they
// would be the critical path of the whole call.  Walks that exhaust their warp budget are finished here by
// a 16-warp CTA: every round each warp probes its own kProbeWindows consecutive windows and summarises them
// (first match, end of the chain that starts there, last match); thread 0 stitches the 16 summaries.
constexpr int kLongWarps = 16;   // warps per CTA: two CTAs share an SM, so one's barrier waits overlap the other's probes
constexpr int kWarpSpan = kProbeWindows;
constexpr int kLongSpan = kLongWarps * kWarpSpan;  // 16384 windows between barriers
#ifndef MEMS_CTA_BUDGET
#define MEMS_CTA_BUDGET 24
#endif
constexpr int kCtaRoundBudget = MEMS_CTA_BUDGET; // rounds one CTA spends on a walk before the whole grid takes it over

struct ChainSummary {
	int first, chain_end, last;  // 1-based distances inside the summarised span, 0 = no match
};
__device__ inline ChainSummary combine_summaries(const int4* child, int n, int unit, int L) {
	ChainSummary r{0, 0, 0};
	int cur = 0;
	bool broken = false;
	for (int i = 0; i < n; ++i) {
		const int4 c = child[i];
		if (!c.x) continue;
		const int f = unit * i + c.x, ce = unit * i + c.y;
		if (!r.first) {
			r.first = f;
			cur = ce;
			broken = c.y != c.z;
		} else if (!broken) {
			if (f - cur > L) broken = true;
			else {
				cur = ce;
				if (c.y != c.z) broken = true;
			}
		}
		r.last = unit * i + c.z;
	}
	r.chain_end = cur;
	return r;
}

// Summary of the kWarpSpan windows after k_base (distances 1..kWarpSpan): first match, last match of the chain that
// starts at the first match, last match overall.
template <class KeyT>
__device__ __forceinline__ int4 warp_span_summary(WarpHit<KeyT>& w, int32_t k_base, int dir) {
	const uint32_t m = w.probe(k_base, dir);
	const int f = w.lowest_set(m);
	if (!f) return make_int4(0, 0, 0, 0);
	const int l = w.highest_set(m);
	const int open_at = w.chain_break(m, f);
	return make_int4(f, open_at ? open_at - w.shape.L - 1 : l, l, 0);
}

template <class KeyT>
__device__ int32_t cta_walk(WarpHit<KeyT>& w, int32_t k0, int dir, int32_t stop_dist, bool* linked, int4* s_sum,
                            int32_t* s_result, int max_rounds, bool* exhausted) {
	const int warp = threadIdx.x >> 5, L = w.shape.L;
	int32_t walked = 0;
	*exhausted = false;
	for (int round = 0;; ++round) {
		if (round == max_rounds) {  // still going: hand it to the whole grid (giant_walk_kernel)
			*exhausted = true;
			*linked = false;
			return walked;
		}
		const int4 mine = warp_span_summary<KeyT>(w, k0 + dir * (walked + kWarpSpan * warp), dir);
		if (w.lane == 0) s_sum[warp] = mine;
		__syncthreads();
		if (threadIdx.x == 0) {
			int32_t cur = 0;  // last confirmed match, as a distance from this round's base
			int state = 0;    // 0 = ran through the whole span, 1 = chain ended, 2 = linked
			for (int i = 0; i < kLongWarps && state == 0; ++i) {
				if (stop_dist >= 0 && stop_dist <= walked + cur + L) state = 2;
				const int4 sm = s_sum[i];
				if (state || !sm.x) continue;
				if (kWarpSpan * i + sm.x - cur > L) {
					state = 1;
				} else {
					cur = kWarpSpan * i + sm.y;
					if (stop_dist >= 0 && stop_dist <= walked + cur + L) state = 2;
					else if (sm.y != sm.z) state = 1;  // a gap > L inside this warp's windows
				}
			}
			if (state == 0 && stop_dist >= 0 && stop_dist <= walked + cur + L) state = 2;
			if (state == 0 && kLongSpan - cur >= L) state = 1;  // a full L-window without a match
			s_result[0] = cur;
			s_result[1] = state;
		}
		__syncthreads();
		walked += s_result[0];
		const int state = s_result[1];
		__syncthreads();
		if (state) {
			*linked = state == 2;
			return walked;
		}
	}
}

template <class KeyT>
__global__ void __launch_bounds__(kLongWarps * 32, 2)
long_walk_right_kernel(MatchArgs a, const __grid_constant__ SeedDesc sd, SegView v, const uint2* __restrict__ defer,
                       const uint32_t* __restrict__ defer_count, uint32_t* __restrict__ next_item,
                       uint32_t* __restrict__ seg_link, uint32_t* __restrict__ seg_reach, uint2* __restrict__ giant,
                       uint32_t* __restrict__ giant_count) {
	__shared__ uint2 s_mem[kLongWarps][kMemberTile];
	__shared__ SeedShape s_shape;
	__shared__ int4 s_sum[kLongWarps];
	__shared__ int32_t s_result[2];
	__shared__ uint32_t s_item;
	const uint32_t n_defer = *defer_count;
	if (n_defer == 0) return;
	load_seed_shape(&s_shape, sd);
	while (true) {  // walks differ in length by orders of magnitude: CTAs pull them from a queue
		if (threadIdx.x == 0) s_item = atomicAdd(next_item, 1u);
		__syncthreads();
		const uint32_t i = s_item;
		__syncthreads();
		if (i >= n_defer) break;
		const uint32_t seg = defer[i].x;
		int32_t c = (int32_t)defer[i].y;
		const uint32_t hi = v.seg_head[seg];
		const uint32_t end = seg + 1 < v.n_seg ? v.seg_head[seg + 1] : v.n_hits;
		const uint32_t h = v.hid[hi];
		const int64_t x0 = (int64_t)(v.hkey[hi] & a.pos_mask);
		WarpHit<KeyT> w{a, s_shape, s_mem[threadIdx.x >> 5], 0, 0, 0, 0, 0, false, (int)(threadIdx.x & 31)};
		w.init(v.hit_start[h], v.hit_len[h] & ~kFirstStrandBit, (v.hit_len[h] & kFirstStrandBit) ? 1u : 0u);
		const bool has_next = seg + 1 < v.n_seg && (v.flags[end] & kFlagSameDiag);
		const int32_t next_at = has_next ? (int32_t)((int64_t)(v.hkey[end] & a.pos_mask) - x0) : 0;
		bool linked, exhausted;
		c += cta_walk<KeyT>(w, c, +1, has_next ? next_at - c : -1, &linked, s_sum, s_result, a.cta_budget, &exhausted);
		if (threadIdx.x == 0) {
			if (exhausted) {
				giant[atomicAdd(giant_count, 1u)] = make_uint2(seg, (uint32_t)c);
			} else {
				seg_link[seg] = linked ? 1u : 0u;
				seg_reach[seg] = (uint32_t)(x0 + c);
			}
		}
	}
}

template <class KeyT>
__global__ void __launch_bounds__(kLongWarps * 32, 2)
long_walk_left_kernel(MatchArgs a, const __grid_constant__ SeedDesc sd, SegView v, const uint2* __restrict__ defer,
                      const uint32_t* __restrict__ defer_count, uint32_t* __restrict__ next_item,
                      const uint32_t* __restrict__ first_excl, uint32_t* __restrict__ comp_left, uint2* __restrict__ giant,
                      uint32_t* __restrict__ giant_count) {
	__shared__ uint2 s_mem[kLongWarps][kMemberTile];
	__shared__ SeedShape s_shape;
	__shared__ int4 s_sum[kLongWarps];
	__shared__ int32_t s_result[2];
	__shared__ uint32_t s_item;
	const uint32_t n_defer = *defer_count;
	if (n_defer == 0) return;
	load_seed_shape(&s_shape, sd);
	while (true) {
		if (threadIdx.x == 0) s_item = atomicAdd(next_item, 1u);
		__syncthreads();
		const uint32_t i = s_item;
		__syncthreads();
		if (i >= n_defer) break;
		const uint32_t seg = defer[i].x;  // always the first segment of its component
		int32_t c = (int32_t)defer[i].y;
		const uint32_t hi = v.seg_head[seg];
		const uint32_t h = v.hid[hi];
		const int64_t x0 = (int64_t)(v.hkey[hi] & a.pos_mask);
		WarpHit<KeyT> w{a, s_shape, s_mem[threadIdx.x >> 5], 0, 0, 0, 0, 0, false, (int)(threadIdx.x & 31)};
		w.init(v.hit_start[h], v.hit_len[h] & ~kFirstStrandBit, (v.hit_len[h] & kFirstStrandBit) ? 1u : 0u);
		bool linked, exhausted;
		c -= cta_walk<KeyT>(w, c, -1, -1, &linked, s_sum, s_result, a.cta_budget, &exhausted);
		if (threadIdx.x == 0) {
			if (exhausted) giant[atomicAdd(giant_count, 1u)] = make_uint2(seg, (uint32_t)c);
			else comp_left[first_excl[seg]] = (uint32_t)(x0 + c);
		}
	}
}


// ---- giant walks ------------------------------------------------------------------------------------
// The few walks that outlast a CTA's budget too (two sequences that agree for Mbp while the others differ) are
// finished by the WHOLE GRID, one walk at a time: every CTA probes its own span of kLongSpan windows, the
// per-CTA summaries meet in global memory, CTA 0 stitches them (same rule as inside a CTA) and a grid-wide
// barrier publishes the result.  One round covers gridDim x 16384 windows (~4.8 M on a B200), so the longest
// diagonal of a genome set costs a handful of rounds instead of being the critical path of the whole call.
template <class KeyT>
__global__ void __launch_bounds__(kLongWarps * 32, 2)
giant_walk_kernel(MatchArgs a, const __grid_constant__ SeedDesc sd, SegView v, const uint2* __restrict__ giant,
                  const uint32_t* __restrict__ giant_count, int dir, const uint32_t* __restrict__ first_excl,
                  uint32_t* __restrict__ seg_link, uint32_t* __restrict__ seg_reach, uint32_t* __restrict__ comp_left,
                  int4* __restrict__ g_sum, int32_t* __restrict__ g_state) {
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	__shared__ uint2 s_mem[kLongWarps][kMemberTile];
	__shared__ SeedShape s_shape;
	__shared__ int4 s_sum[kLongWarps];
	const int warp = threadIdx.x >> 5;
	const uint32_t n_items = *giant_count;
	if (n_items == 0) return;  // grid-uniform
	load_seed_shape(&s_shape, sd);
	const int L = s_shape.L;
	for (uint32_t it = 0; it < n_items; ++it) {
		const uint32_t seg = giant[it].x;
		const int32_t c0 = (int32_t)giant[it].y;
		const uint32_t hi = v.seg_head[seg];
		const uint32_t end = seg + 1 < v.n_seg ? v.seg_head[seg + 1] : v.n_hits;
		const uint32_t h = v.hid[hi];
		const int64_t x0 = (int64_t)(v.hkey[hi] & a.pos_mask);
		WarpHit<KeyT> w{a, s_shape, s_mem[warp], 0, 0, 0, 0, 0, false, (int)(threadIdx.x & 31)};
		w.init(v.hit_start[h], v.hit_len[h] & ~kFirstStrandBit, (v.hit_len[h] & kFirstStrandBit) ? 1u : 0u);
		const bool has_next = dir > 0 && seg + 1 < v.n_seg && (v.flags[end] & kFlagSameDiag);
		const int32_t stop_dist = has_next ? (int32_t)((int64_t)(v.hkey[end] & a.pos_mask) - x0) - c0 : -1;
		int32_t walked = 0;
		int state = 0;
		while (!state) {
			const int4 mine = warp_span_summary<KeyT>(w, c0 + dir * (walked + kLongSpan * (int32_t)blockIdx.x + kWarpSpan * warp), dir);
			if (w.lane == 0) s_sum[warp] = mine;
			__syncthreads();
			if (threadIdx.x == 0) {
				const ChainSummary cs = combine_summaries(s_sum, kLongWarps, kWarpSpan, L);
				g_sum[blockIdx.x] = make_int4(cs.first, cs.chain_end, cs.last, 0);
			}
			grid.sync();
			if (blockIdx.x == 0 && threadIdx.x == 0) {
				int32_t cur = 0;  // last confirmed match, as a distance from this round's base
				int st = 0;       // 0 = ran through the whole span, 1 = chain ended, 2 = linked
				const int n_cta = (int)gridDim.x;
				for (int i = 0; i < n_cta && st == 0; ++i) {
					if (stop_dist >= 0 && stop_dist <= walked + cur + L) st = 2;
					const int4 sm = g_sum[i];
					if (st || !sm.x) continue;
					if (kLongSpan * i + sm.x - cur > L) {
						st = 1;
					} else {
						cur = kLongSpan * i + sm.y;
						if (stop_dist >= 0 && stop_dist <= walked + cur + L) st = 2;
						else if (sm.y != sm.z) st = 1;
					}
				}
				if (st == 0 && stop_dist >= 0 && stop_dist <= walked + cur + L) st = 2;
				if (st == 0 && kLongSpan * n_cta - cur >= L) st = 1;
				g_state[0] = cur;
				g_state[1] = st;
			}
			grid.sync();
			walked += g_state[0];
			state = g_state[1];
		}
		if (blockIdx.x == 0 && threadIdx.x == 0) {
			if (dir > 0) {
				seg_link[seg] = state == 2 ? 1u : 0u;
				seg_reach[seg] = (uint32_t)(x0 + c0 + walked);
			} else {
				comp_left[first_excl[seg]] = (uint32_t)(x0 + c0 - walked);
			}
		}
		__syncthreads();  // s_mem is rewritten by the next item
	}
}

// ------------------------------------------------------------------------------------------------ 6. emit
__global__ void emit_size_kernel(MatchArgs a, const uint32_t* __restrict__ comp_rep, const uint32_t* __restrict__ hid,
                                 const uint32_t* __restrict__ hit_start, const uint16_t* __restrict__ hit_len, uint32_t n_comp,
                                 uint32_t* __restrict__ rec_size, uint32_t* __restrict__ rec_group) {
	const uint32_t comp = blockIdx.x * blockDim.x + threadIdx.x;
	if (comp >= n_comp) return;
	const uint32_t h = hid[comp_rep[comp]];
	const uint32_t len = hit_len[h] & ~kFirstStrandBit;
	uint32_t count = (uint32_t)a.n_seqs;
	if (a.max_group != a.n_seqs) {  // several problems in the batch: SeqCount is the problem's
		const SeqMeta m = a.meta[a.vals[hit_start[h]] >> a.pos_bits];
		count = m.group_count;
		if (rec_group) rec_group[comp] = m.group;
	} else if (rec_group) {
		rec_group[comp] = 0;
	}
	rec_size[comp] = 2u + (a.mode == MEMS_MODE_REPEAT ? len : count);
}

template <class KeyT>
__global__ void emit_kernel(MatchArgs a, int L, const uint64_t* __restrict__ hkey, const uint32_t* __restrict__ hid,
                            const uint32_t* __restrict__ hit_start, const uint16_t* __restrict__ hit_len,
                            const uint32_t* __restrict__ comp_rep, const uint32_t* __restrict__ comp_left,
                            const uint32_t* __restrict__ comp_right, const uint32_t* __restrict__ rec_off, uint32_t n_comp,
                            int64_t* __restrict__ flat) {
	const uint32_t comp = blockIdx.x * blockDim.x + threadIdx.x;
	if (comp >= n_comp) return;
	const uint32_t hi = comp_rep[comp];
	const uint32_t h = hid[hi];
	const uint32_t s = hit_start[h];
	const uint32_t len = hit_len[h] & ~kFirstStrandBit;
	const uint32_t sf = (hit_len[h] & kFirstStrandBit) ? 1u : 0u;
	const int64_t x0 = (int64_t)(hkey[hi] & a.pos_mask);
	const int64_t kl = (int64_t)comp_left[comp] - x0, kr = (int64_t)comp_right[comp] - x0;
	int64_t* rec = flat + rec_off[comp];
	uint32_t seqcount = a.mode == MEMS_MODE_REPEAT ? len : (uint32_t)a.n_seqs, first_seq = 0;
	if (a.max_group != a.n_seqs) {  // several problems in the batch: sequence numbers relative to the problem's first
		const SeqMeta m = a.meta[a.vals[s] >> a.pos_bits];
		seqcount = m.group_count;
		first_seq = m.group_first;
	}
	rec[0] = seqcount;
	rec[1] = kr - kl + L;
	if (a.mode != MEMS_MODE_REPEAT)
		for (uint32_t g = 0; g < seqcount; ++g) rec[2 + g] = 0;  // NO_MATCH
	MemberIter<KeyT> it(a, s, len);
	uint32_t val, st, idx = 0;
	while (it.next(val, st)) {
		const uint32_t o = st ^ sf;
		const int64_t p = val & a.pos_mask;
		// forward member: first covered base p+kl (1-based start p+kl+1); reverse member: covers p-kr .. p-kl+L-1
		const int64_t start = o ? -(p - kr + 1) : (p + kl + 1);
		const uint32_t slot = a.mode == MEMS_MODE_REPEAT ? idx : (val >> a.pos_bits) - first_seq;
		rec[2 + slot] = start;
		++idx;
	}
}

// the (few) components next to a foreign diagonal of their hash bucket, with the offset of their record: the host
// compares just these for duplicates
__global__ void suspect_list_kernel(const uint8_t* __restrict__ comp_suspect, const uint32_t* __restrict__ rec_off, uint32_t n_comp,
                                    uint2* __restrict__ list, uint32_t cap, uint32_t* __restrict__ count) {
	const uint32_t comp = blockIdx.x * blockDim.x + threadIdx.x;
	if (comp >= n_comp || !comp_suspect[comp]) return;
	const uint32_t at = atomicAdd(count, 1u);
	if (at < cap) list[at] = make_uint2(comp, rec_off[comp]);
}

// records of the listed (suspect) components, side by side (records of one size R)
__global__ void suspect_records_kernel(const int64_t* __restrict__ flat, const uint2* __restrict__ list, uint32_t n_sus, uint32_t R,
                                       int64_t* __restrict__ out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_sus * R) return;
	out[i] = flat[(size_t)list[i / R].y + i % R];
}
// the record list without the dropped components (ascending list of component numbers; records of one size R)
__global__ void drop_records_kernel(const int64_t* __restrict__ flat, uint32_t n_comp, uint32_t R, const uint32_t* __restrict__ drops,
                                    uint32_t n_drop, int64_t* __restrict__ out) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (uint64_t)n_comp * R) return;
	const uint32_t comp = (uint32_t)(i / R);
	uint32_t lo = 0, hi = n_drop;  // dropped components before (or at) comp
	while (lo < hi) {
		const uint32_t mid = (lo + hi) / 2;
		if (drops[mid] <= comp) lo = mid + 1;
		else hi = mid;
	}
	if (lo && drops[lo - 1] == comp) return;
	out[i - (uint64_t)lo * R] = flat[i];
}

// hit members in merged order, for the host-side table emulation (ORDER_REFERENCE)
template <class KeyT>
__global__ void gather_members_kernel(MatchArgs a, const uint32_t* __restrict__ hit_start, const uint16_t* __restrict__ hit_len,
                                      const uint32_t* __restrict__ mem_off, uint32_t n_hits, uint32_t* __restrict__ mem_val,
                                      uint8_t* __restrict__ mem_strand) {
	const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
	if (h >= n_hits) return;
	MemberIter<KeyT> it(a, hit_start[h], hit_len[h] & ~kFirstStrandBit);
	uint32_t val, st, at = mem_off[h];
	while (it.next(val, st)) {
		mem_val[at] = val;
		mem_strand[at] = (uint8_t)st;
		++at;
	}
}

__global__ void hit_len_widen_kernel(const uint16_t* __restrict__ hit_len, uint32_t n_hits, uint32_t* __restrict__ out) {
	const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
	if (h < n_hits) out[h] = hit_len[h] & ~kFirstStrandBit;
}

// record (= component) index of every hit: sorted entry -> segment -> component
__global__ void hit_record_kernel(const uint32_t* __restrict__ hid, const uint32_t* __restrict__ is_head,
                                  const uint32_t* __restrict__ seg_of, const uint32_t* __restrict__ first,
                                  const uint32_t* __restrict__ first_excl, uint32_t n_hits, uint32_t* __restrict__ rec_of_hit) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_hits) return;
	// seg_of is an exclusive scan of is_head: a head's own segment is seg_of[i], a follower's is seg_of[i]-1
	const uint32_t seg = is_head[i] ? seg_of[i] : seg_of[i] - 1;
	rec_of_hit[hid[i]] = first_excl[seg] - (first[seg] ? 0u : 1u);
}

// ================================================================================================ host side
namespace {

uint32_t d2h_u32(Ctx* c, const uint32_t* d) {
	uint32_t v = 0;
	c->fetch(&v, d, sizeof v);
	return v;
}

struct Rec {
	const int64_t* p;  // [SeqCount, Len, starts...]
	size_t size() const { return (size_t)p[0] + 2; }
};
bool rec_less(const Rec& x, const Rec& y) { return std::lexicographical_compare(x.p, x.p + x.size(), y.p, y.p + y.size()); }
bool rec_equal(const Rec& x, const Rec& y) { return x.size() == y.size() && std::equal(x.p, x.p + x.size(), y.p); }

std::vector<Rec> split_records(const int64_t* flat, size_t n) {
	std::vector<Rec> recs;
	for (size_t i = 0; i < n; i += (size_t)flat[i] + 2) recs.push_back({flat + i});
	return recs;
}

// ---- the reference's hash table, replayed on the host over (hit, extended match) pairs ------------------
// MemHash::AddHashEntry (MemHash.cpp:209-251) with MheCompare (MatchHashEntry.h:121-143),
// MatchHashEntry::Contains (MatchHashEntry.cpp:164-200), strict_start_lessthan_ptr (:48-67) and
// CalculateOffset (:141-160).  Needed because the table's lower_bound over a comparator that is not a
// strict weak order decides WHICH hits are dropped as collisions and the output order.
typedef TableEntry Entry;  // common.cuh
inline int64_t e_start(const Entry& e, uint32_t i) { return i < e.seqcount ? e.start[i] : 0; }
inline uint32_t e_first(const Entry& e) {
	for (uint32_t i = 0; i < e.seqcount; ++i)
		if (e.start[i] != 0) return i;
	return 0xffffffffu;
}
void e_calc_offset(Entry& e) {
	e.offset = 0;
	uint32_t i = e_first(e);
	if (i == 0xffffffffu) return;
	const int64_t ref = e.start[i];
	for (++i; i < e.seqcount; ++i)
		if (e.start[i] != 0) {
			int64_t t = e.start[i] - ref;
			if (e.start[i] < 0) t -= e.len;
			e.offset += t;
		}
}
bool e_contains(const Entry& A, const Entry& m) {
	if (A.seqcount != m.seqcount || A.offset != m.offset) return false;
	uint32_t i = e_first(m);
	const int64_t diff = e_start(m, i) - e_start(A, i);
	if (e_start(A, i) == 0) return false;
	if (diff < 0 || A.len < m.len + diff) return false;
	const int64_t diff_rc = m.len - A.len + diff;
	for (++i; i < m.seqcount; ++i) {
		const int64_t d = m.start[i] - A.start[i];
		if (m.start[i] == 0 && A.start[i] == 0) continue;
		else if (m.start[i] < 0 && diff_rc == d) continue;
		else if (diff != d) return false;
	}
	return true;
}
bool e_strict_less(const Entry& a, const Entry& b) {
	const int fa = (int)e_first(a), fb = (int)e_first(b);
	const int start_diff = fa - fb;
	if (start_diff == 0) {
		const uint32_t cnt = std::min(a.seqcount, b.seqcount);
		for (uint32_t s = 0; s < cnt; ++s) {
			int64_t x = a.start[s], y = b.start[s];
			if (x < 0) x = -x + a.len - a.mersize;
			if (y < 0) y = -y + b.len - b.mersize;
			if (x != y) return x < y;
		}
	}
	return start_diff < 0;
}
bool e_compare(const Entry& a, const Entry& b) {
	const uint32_t fa = e_first(a), fb = e_first(b);
	if (fa > fb) return true;
	if (fa == fb) {
		for (uint32_t i = fa; i < a.seqcount; ++i) {
			const int64_t as = e_start(a, i), bs = e_start(b, i);
			if (as == 0 && bs != 0) return true;
			else if (as != 0 && bs == 0) return false;
		}
		if (e_contains(a, b) || e_contains(b, a)) return false;
		return e_strict_less(a, b);
	}
	return false;
}
size_t bucket_lower_bound(const std::vector<Entry*>& v, const Entry& x) {
	size_t first = 0, len = v.size();
	while (len > 0) {
		size_t half = len >> 1, mid = first + half;
		if (e_compare(*v[mid], x)) {
			first = mid + 1;
			len = len - half - 1;
		} else
			len = half;
	}
	return first;
}

}  // namespace


// hits of one sorted union, in key order
struct HitSet {
	DevBuf<uint32_t> start;  // first union entry of the hit
	DevBuf<uint16_t> len;    // entries in the hit (bit 15 is set later: strand of the first member)
	uint32_t n = 0;
	uint32_t max_run = 0;
};

// ---- stage A: equal-seed runs of a sorted union -> hits in key order
template <class KeyT>
static void find_hits(Ctx* c, const MatchArgs& a, HitSet& hits) {
	const uint32_t n = a.n;
	const uint64_t n_tiles = ((uint64_t)n + kRunTile - 1) / kRunTile;
	// a hit has at least two entries: a tile's kRunTile / 2 staging slots hold whatever it finds
	DevBuf<uint32_t> stage_start(c, n_tiles * (kRunTile / 2)), tile_hits(c, n_tiles), tile_off(c, n_tiles), scalars(c, 2);
	DevBuf<uint16_t> stage_len(c, n_tiles * (kRunTile / 2));
	MEMS_CUDA(cudaMemsetAsync(scalars.p, 0, 2 * sizeof(uint32_t), c->stream));
	{
		KernelScope ks(c, "run_hits", (double)n * (sizeof(KeyT) + 4.0));
		run_hits_kernel<KeyT><<<(unsigned)n_tiles, kScanBlock, 0, c->stream>>>(a, stage_start.p, stage_len.p, tile_hits.p, scalars.p);
		MEMS_CUDA(cudaGetLastError());
	}
	exclusive_scan_u32(c, tile_hits.p, tile_off.p, n_tiles, scalars.p + 1);
	uint32_t h_scal[2];  // [0] longest run, [1] hits
	c->fetch(h_scal, scalars.p, sizeof h_scal);
	hits.max_run = h_scal[0];
	hits.n = h_scal[1];
	if (hits.n == 0) return;
	hits.start = DevBuf<uint32_t>(c, hits.n);
	hits.len = DevBuf<uint16_t>(c, hits.n);
	{
		KernelScope ks(c, "hit_gather", (double)hits.n * 12.0);
		hit_gather_kernel<<<(unsigned)n_tiles, kScanBlock, 0, c->stream>>>(stage_start.p, stage_len.p, tile_hits.p, tile_off.p, hits.start.p, hits.len.p);
		MEMS_CUDA(cudaGetLastError());
	}
}

// ------------------------------------------------------------------------------------------------ pairwise policy
// PairwiseMatchFinder::EnumerateMatches (PairwiseMatchFinder.cpp:37-71): of an equal-seed run keep the
// sequences that hold the seed exactly once (others may hold it any number of times) and hash every pair
// (i < j in sequence order) as a two-member hit.  Pairs are not contiguous in the union, so their members
// are written out as a synthetic union of two-entry "runs" (strand-0 member first, like a real run), which
// the extension stage reads exactly like the real one.
template <class KeyT>
__device__ uint64_t unique_sequences(const MatchArgs& a, uint32_t i, uint32_t* run_len) {
	const uint64_t mk = masked_of<KeyT>(a.keys, i);
	uint64_t once = 0, more = 0;
	uint32_t j = i;
	const uint32_t g0 = a.n_seqs > 64 ? a.meta[a.vals[i] >> a.pos_bits].group_first : 0u;  // bits relative to the run's problem
	while (j < a.n && masked_of<KeyT>(a.keys, j) == mk && j - i <= kRunCap) {
		const uint64_t bit = 1ull << ((a.vals[j] >> a.pos_bits) - g0);
		more |= once & bit;
		once |= bit;
		++j;
	}
	*run_len = j - i;
	return j - i > kRunCap ? 0ull : (once & ~more);
}

template <class KeyT>
__global__ void __launch_bounds__(kScanBlock)
pair_count_kernel(MatchArgs a, uint32_t* __restrict__ pair_count, uint32_t* __restrict__ max_run) {
	const uint32_t i = blockIdx.x * kScanBlock + threadIdx.x;
	uint32_t cnt = 0, run = 0;
	if (i < a.n) {
		const uint64_t mk = masked_of<KeyT>(a.keys, i);
		const bool head = i == 0 || masked_of<KeyT>(a.keys, i - 1) != mk;
		if (head && i + 1 < a.n && masked_of<KeyT>(a.keys, i + 1) == mk) {
			const uint32_t u = __popcll(unique_sequences<KeyT>(a, i, &run));
			cnt = u * (u - 1) / 2;
		}
		pair_count[i] = cnt;
	}
	run = __reduce_max_sync(0xffffffffu, run);
	if ((threadIdx.x & 31) == 0 && run > 1) atomicMax(max_run, run);
}

template <class KeyT>
__global__ void __launch_bounds__(kScanBlock)
pair_emit_kernel(MatchArgs a, const uint32_t* __restrict__ pair_count, const uint32_t* __restrict__ pair_off,
                 uint32_t* __restrict__ out_vals, KeyT* __restrict__ out_keys) {
	const uint32_t i = blockIdx.x * kScanBlock + threadIdx.x;
	if (i >= a.n || pair_count[i] == 0) return;
	uint32_t run;
	const uint64_t uniq = unique_sequences<KeyT>(a, i, &run);
	uint32_t at = 2 * pair_off[i];
	const uint32_t g0 = a.n_seqs > 64 ? a.meta[a.vals[i] >> a.pos_bits].group_first : 0u;
	// sequences in ascending order; a sequence's (single) entry is found by scanning the run
	for (uint64_t ra = uniq; ra; ra &= ra - 1) {
		const uint32_t ga = g0 + __ffsll((long long)ra) - 1;
		uint32_t ja = i;
		while ((a.vals[ja] >> a.pos_bits) != ga) ++ja;
		for (uint64_t rb = ra & (ra - 1); rb; rb &= rb - 1) {
			const uint32_t gb = g0 + __ffsll((long long)rb) - 1;
			uint32_t jb = i;
			while ((a.vals[jb] >> a.pos_bits) != gb) ++jb;
			const uint32_t lo = ja < jb ? ja : jb, hi = ja < jb ? jb : ja;  // union order: strand 0 first, then (seq,pos)
			out_vals[at] = a.vals[lo];
			out_keys[at] = (KeyT)strand_of<KeyT>(a.keys, lo);
			out_vals[at + 1] = a.vals[hi];
			out_keys[at + 1] = (KeyT)strand_of<KeyT>(a.keys, hi);
			at += 2;
		}
	}
}

__global__ void pair_hits_kernel(uint32_t n_hits, uint32_t* __restrict__ hit_start, uint16_t* __restrict__ hit_len) {
	const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
	if (h < n_hits) {
		hit_start[h] = 2 * h;
		hit_len[h] = 2;
	}
}

// stage A for the pairwise policy: hits + the synthetic union their members live in
template <class KeyT>
static void find_pair_hits(Ctx* c, const MatchArgs& a, HitSet& hits, DevBuf<uint32_t>& pvals, DevBuf<KeyT>& pkeys) {
	const uint32_t n = a.n, n_blocks = (n + kScanBlock - 1) / kScanBlock;
	DevBuf<uint32_t> cnt(c, n), off(c, n), scalars(c, 2);
	MEMS_CUDA(cudaMemsetAsync(scalars.p, 0, 2 * sizeof(uint32_t), c->stream));
	{
		KernelScope ks(c, "pair_scan", (double)n * (sizeof(KeyT) + 4.0));
		pair_count_kernel<KeyT><<<n_blocks, kScanBlock, 0, c->stream>>>(a, cnt.p, scalars.p + 0);
		MEMS_CUDA(cudaGetLastError());
	}
	exclusive_scan_u32(c, cnt.p, off.p, n, scalars.p + 1);
	uint32_t h_scal[2];
	c->fetch(h_scal, scalars.p, sizeof h_scal);
	hits.max_run = h_scal[0];
	hits.n = h_scal[1];
	if (hits.n == 0) return;
	if (hits.n >= (1u << 30)) throw Error(MEMS_ERR_UNSUPPORTED, "more than 2^30 pairwise hits");
	pvals = DevBuf<uint32_t>(c, 2 * (size_t)hits.n);
	pkeys = DevBuf<KeyT>(c, 2 * (size_t)hits.n);
	hits.start = DevBuf<uint32_t>(c, hits.n);
	hits.len = DevBuf<uint16_t>(c, hits.n);
	{
		KernelScope ks(c, "pair_emit");
		pair_emit_kernel<KeyT><<<n_blocks, kScanBlock, 0, c->stream>>>(a, cnt.p, off.p, pvals.p, pkeys.p);
		MEMS_CUDA(cudaGetLastError());
		pair_hits_kernel<<<(hits.n + 255) / 256, 256, 0, c->stream>>>(hits.n, hits.start.p, hits.len.p);
		MEMS_CUDA(cudaGetLastError());
	}
}

// several problems in one batch: per record (in device order) its problem, and which records are duplicates to drop
struct ManyOut {
	std::vector<uint32_t> group;
	std::vector<char> drop;
};

// ---- stage B: hits (members readable through a.keys / a.vals) -> extended, distinct matches
template <class KeyT>
static void extend_hits(std::shared_ptr<Ctx> ctx, const MatchArgs& a, const SeedDesc& sd, HitSet& hits, int order,
                        uint32_t table_size, MatchResult& out, HashTable* persistent = nullptr, uint64_t* given_hkey = nullptr,
                        ManyOut* many = nullptr) {
	Ctx* c = ctx.get();
	const int L = sd.L;
	const int mode = a.mode;
	const uint32_t n_hits = hits.n;
	DevBuf<uint32_t>& hit_start = hits.start;
	DevBuf<uint16_t>& hit_len = hits.len;
	DevBuf<uint32_t> scalars(c, 8);
	MEMS_CUDA(cudaMemsetAsync(scalars.p, 0, 8 * sizeof(uint32_t), c->stream));

	// ---- 2. describe + 3. sort by (diagonal hash, first-member position)
	// (given_hkey: the hits arrive described — sharded path, where the key travelled with the hit through the
	// diagonal exchange; only the identity permutation and the digit histograms are made here)
	SortPlan plan = make_sort_plan(a.hit_key_bits);
	DevBuf<uint64_t> hk_a(c, given_hkey ? 0 : n_hits), hk_b(c, n_hits);
	DevBuf<uint32_t> hid_a(c, n_hits), hid_b(c, n_hits), hist(c, (size_t)plan.n_passes * 256);
	DevBuf<HitSigRec> sig;  // signatures in hit order (hits that arrive described — sharded — are compared through their members)
	if (!given_hkey) sig = DevBuf<HitSigRec>(c, n_hits);
	MEMS_CUDA(cudaMemsetAsync(hist.p, 0, (size_t)plan.n_passes * 256 * sizeof(uint32_t), c->stream));
	const uint32_t hit_blocks = (n_hits + 255) / 256;
	const uint32_t describe_blocks = std::min(hit_blocks, (uint32_t)c->sm_count * 8u);
	if (given_hkey) {
		KernelScope ks(c, "hit_histogram", (double)n_hits * 8.0);
		hit_histogram_kernel<<<describe_blocks, 256, 0, c->stream>>>(given_hkey, n_hits, hid_a.p, hist.p, plan);
		MEMS_CUDA(cudaGetLastError());
	} else {
		KernelScope ks(c, "hit_describe");
		hit_describe_kernel<KeyT><<<describe_blocks, 256, 0, c->stream>>>(a, hit_start.p, hit_len.p, n_hits, hk_a.p, hid_a.p,
		                                                                  hist.p, plan, sig.p);
		MEMS_CUDA(cudaGetLastError());
	}
	void* kp[2] = {given_hkey ? given_hkey : hk_a.p, hk_b.p};
	uint32_t* vp[2] = {hid_a.p, hid_b.p};
	const int r = radix_sort_pairs(c, true, kp, vp, n_hits, plan, hist.p, "hit_sort_pass");
	const uint64_t* hkey = (const uint64_t*)kp[r];
	const uint32_t* hid = vp[r];

	// ---- 4. segments
	DevBuf<uint8_t> flags(c, n_hits), suspect(c, n_hits);
	DevBuf<uint32_t> is_head(c, n_hits), seg_of(c, n_hits);
	MEMS_CUDA(cudaMemsetAsync(suspect.p, 0, n_hits, c->stream));
	{
		KernelScope ks(c, "segment_flag");
		segment_flag_kernel<KeyT><<<hit_blocks, 256, 0, c->stream>>>(a, L, hkey, hid, hit_start.p, hit_len.p, n_hits,
		                                                             flags.p, is_head.p, scalars.p + 4, suspect.p, sig.p);
		MEMS_CUDA(cudaGetLastError());
	}
	exclusive_scan_u32(c, is_head.p, seg_of.p, n_hits, scalars.p + 2);
	const uint32_t n_seg = d2h_u32(c, scalars.p + 2);
	out.n_segments = n_seg;
	DevBuf<uint32_t> seg_head(c, n_seg), seg_x_a(c, n_seg), seg_x_b(c, n_seg), seg_id_a(c, n_seg), seg_id_b(c, n_seg);
	{
		KernelScope ks(c, "segment_compact");
		segment_compact_kernel<<<hit_blocks, 256, 0, c->stream>>>(is_head.p, seg_of.p, hkey, a.pos_mask, n_hits, seg_head.p,
		                                                          seg_x_a.p, seg_id_a.p);
		MEMS_CUDA(cudaGetLastError());
	}
	// Visit segments in order of genome position, not of diagonal hash: the walks of all diagonals that cross
	// one region then run together and find that region's keys in L2 instead of going to HBM one by one.
	const uint32_t* seg_order;
	{
		SortPlan splan = make_sort_plan(a.pos_bits);
		DevBuf<uint32_t> shist(c, (size_t)splan.n_passes * 256);
		launch_histogram(c, false, seg_x_a.p, n_seg, splan, shist.p);
		void* skp[2] = {seg_x_a.p, seg_x_b.p};
		uint32_t* svp[2] = {seg_id_a.p, seg_id_b.p};
		const int sr = radix_sort_pairs(c, false, skp, svp, n_seg, splan, shist.p, "segment_order_pass");
		seg_order = sr ? seg_id_b.p : seg_id_a.p;
	}

	// ---- 5. extend: right walks link segments into components, left walks finish each component
	SegView v{hkey, hid, hit_start.p, hit_len.p, flags.p, seg_head.p, n_hits, n_seg, seg_order, suspect.p};
	DevBuf<uint32_t> seg_link(c, n_seg), seg_reach(c, n_seg), first(c, n_seg), first_excl(c, n_seg);
	const SegView& v_for_giant = v;
	DevBuf<uint2> defer(c, n_seg), defer_left(c, n_seg);
	DevBuf<uint32_t> seg_left(c, n_seg);
	DevBuf<uint8_t> seg_left_state(c, n_seg);
	MEMS_CUDA(cudaMemsetAsync(seg_left_state.p, 0, n_seg, c->stream));
	uint32_t* defer_count = scalars.p + 6;  // [6] right walks, [7] left walks handed to whole CTAs
	const uint32_t long_grid = 2u * (uint32_t)c->sm_count;
	DevBuf<uint32_t> queue_heads(c, 4);  // [0],[1] work-queue heads; [2],[3] giant-walk counts (right, left)
	MEMS_CUDA(cudaMemsetAsync(queue_heads.p, 0, 4 * sizeof(uint32_t), c->stream));
	DevBuf<uint2> giant(c, n_seg);
	int giant_blocks_per_sm = 0;
	MEMS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&giant_blocks_per_sm, giant_walk_kernel<KeyT>, kLongWarps * 32, 0));
	const uint32_t giant_grid = (uint32_t)std::max(1, std::min(giant_blocks_per_sm, 2)) * (uint32_t)c->sm_count;
	DevBuf<int4> g_sum(c, giant_grid);
	DevBuf<int32_t> g_state(c, 2);
	auto launch_giant = [&](int dir, uint32_t* count, uint32_t* comp_left_p) {
		SeedDesc sd_arg = sd;
		int dir_arg = dir;
		MatchArgs a_arg = a;
		SegView v_arg = v_for_giant;
		const uint2* giant_p = giant.p;
		const uint32_t* count_p = count;
		const uint32_t* fe_p = first_excl.p;
		uint32_t* link_p = seg_link.p;
		uint32_t* reach_p = seg_reach.p;
		int4* sum_p = g_sum.p;
		int32_t* state_p = g_state.p;
		void* args[] = {&a_arg, &sd_arg, &v_arg, &giant_p, &count_p, &dir_arg, &fe_p, &link_p, &reach_p, &comp_left_p,
		                &sum_p, &state_p};
		KernelScope ks(c, dir > 0 ? "giant_walk_right" : "giant_walk_left");
		MEMS_CUDA(cudaLaunchCooperativeKernel((void*)giant_walk_kernel<KeyT>, dim3(giant_grid), dim3(kLongWarps * 32), args, 0,
		                                      c->stream));
	};
	constexpr uint32_t kSegsPerWalkBlock = kExtendWarps * kGroupsPerWarp;
	const uint32_t walk_blocks = (n_seg + kSegsPerWalkBlock - 1) / kSegsPerWalkBlock, seg_blocks = (n_seg + 255) / 256;
	DevBuf<SegDesc> seg_desc(c, n_seg);
	DevBuf<uint2> seg_members(c, (size_t)n_seg * kGroupMembers);
	DevBuf<uint32_t> slot_of_seg(c, n_seg);
	{
		KernelScope ks(c, "segment_stage");
		segment_stage_kernel<KeyT><<<(n_seg + 31) / 32, 256, 0, c->stream>>>(a, v, seg_desc.p, seg_members.p, slot_of_seg.p);
		MEMS_CUDA(cudaGetLastError());
	}
	{
		KernelScope ks(c, "walk_right");
		walk_right_kernel<KeyT><<<walk_blocks, kExtendWarps * 32, 0, c->stream>>>(a, sd, v, seg_desc.p, seg_members.p, seg_link.p, seg_reach.p,
		                                                                          defer.p, defer_count, seg_left.p, seg_left_state.p,
		                                                                          defer_left.p, defer_count + 1);
		MEMS_CUDA(cudaGetLastError());
	}
	{
		KernelScope ks(c, "long_walk_right");
		long_walk_right_kernel<KeyT><<<long_grid, kLongWarps * 32, 0, c->stream>>>(a, sd, v, defer.p, defer_count,
		                                                                           queue_heads.p, seg_link.p, seg_reach.p, giant.p, queue_heads.p + 2);
		MEMS_CUDA(cudaGetLastError());
	}
	launch_giant(+1, queue_heads.p + 2, nullptr);
	{
		KernelScope ks(c, "chain_first");
		chain_first_kernel<<<seg_blocks, 256, 0, c->stream>>>(seg_link.p, n_seg, first.p);
		MEMS_CUDA(cudaGetLastError());
	}
#ifdef MEMS_WALK_STATS
	{
		unsigned long long st[48];
		MEMS_CUDA(cudaMemcpyFromSymbol(st, g_walk_stats, sizeof st));
		fprintf(stderr, "walk_right (group phase): n_seg=%u probes=%llu; to the warp: right %llu left %llu; walks by log2(probes):", n_seg, st[0], st[8], st[9]);
		for (int i = 2; i < 8; ++i) fprintf(stderr, " %llu", st[i]);
		fprintf(stderr, "\n  finished by a group, by log2(distance): linked");
		for (int i = 16; i < 32; ++i) fprintf(stderr, " %llu", st[i]);
		fprintf(stderr, "\n  ended");
		for (int i = 32; i < 48; ++i) fprintf(stderr, " %llu", st[i]);
		fprintf(stderr, "\n");
		memset(st, 0, sizeof st);
		MEMS_CUDA(cudaMemcpyToSymbol(g_walk_stats, st, sizeof st));
	}
#endif
	exclusive_scan_u32(c, first.p, first_excl.p, n_seg, scalars.p + 3);
	uint32_t h_tail[2];  // [n_comp, diagonal-hash collision seen]
	c->fetch(h_tail, scalars.p + 3, sizeof h_tail);
	const uint32_t n_comp = h_tail[0];
	const bool collision_seen = h_tail[1] != 0;
	DevBuf<uint32_t> comp_rep(c, n_comp), comp_left(c, n_comp), comp_right(c, n_comp), rec_size(c, n_comp), rec_off(c, n_comp);
	DevBuf<uint32_t> rec_group(c, many ? n_comp : 1);
	DevBuf<uint8_t> comp_suspect(c, n_comp);
	if (collision_seen) MEMS_CUDA(cudaMemsetAsync(comp_suspect.p, 0, n_comp, c->stream));
	DevBuf<uint32_t> left_todo(c, n_comp), left_todo_count(c, 1);
	MEMS_CUDA(cudaMemsetAsync(left_todo_count.p, 0, sizeof(uint32_t), c->stream));
	{
		KernelScope ks(c, "finish_segments");
		finish_segments_kernel<<<seg_blocks, 256, 0, c->stream>>>(v, seg_link.p, seg_reach.p, first.p, first_excl.p, seg_left.p,
		                                                          seg_left_state.p, comp_rep.p, comp_left.p, comp_right.p,
		                                                          comp_suspect.p, slot_of_seg.p, left_todo.p, left_todo_count.p);
		MEMS_CUDA(cudaGetLastError());
	}
	{
		KernelScope ks(c, "walk_left");  // at most one listed segment per component: warps past the list's end leave at once
		walk_left_kernel<KeyT><<<(n_comp + kSegsPerWalkBlock - 1) / kSegsPerWalkBlock, kExtendWarps * 32, 0, c->stream>>>(
		    a, sd, seg_desc.p, seg_members.p, left_todo.p, left_todo_count.p, first_excl.p, comp_left.p, defer_left.p, defer_count + 1);
		MEMS_CUDA(cudaGetLastError());
	}
	{
		KernelScope ks(c, "long_walk_left");
		long_walk_left_kernel<KeyT><<<long_grid, kLongWarps * 32, 0, c->stream>>>(a, sd, v, defer_left.p, defer_count + 1,
		                                                                          queue_heads.p + 1, first_excl.p, comp_left.p, giant.p, queue_heads.p + 3);
		MEMS_CUDA(cudaGetLastError());
	}
	launch_giant(-1, queue_heads.p + 3, comp_left.p);

	// ---- 6. emit
	const uint32_t comp_blocks = (n_comp + 255) / 256;
	{
		KernelScope ks(c, "emit_size");
		emit_size_kernel<<<comp_blocks, 256, 0, c->stream>>>(a, comp_rep.p, hid, hit_start.p, hit_len.p, n_comp, rec_size.p,
		                                                     many ? rec_group.p : nullptr);
		MEMS_CUDA(cudaGetLastError());
	}
	exclusive_scan_u32(c, rec_size.p, rec_off.p, n_comp, scalars.p + 5);
	const uint32_t sus_cap = 1u << 16;
	DevBuf<uint2> sus_list(c, collision_seen ? sus_cap : 1);
	DevBuf<uint32_t> sus_count(c, 1);
	struct HostWord {  // a mapped page-locked word the device writes (Ctx::fetch_async), given back when the call is over
		Ctx* c;
		uint32_t* w;
		bool queued = false;
		explicit HostWord(Ctx* ctx) : c(ctx), w(ctx->host_words_get()) { w[0] = 0; }
		~HostWord() {
			if (queued) cudaStreamSynchronize(c->stream);  // the write into it must have landed before the word is reused
			c->host_words_put(w);
		}
	} sus_word(c);
	volatile uint32_t& h_sus_count = sus_word.w[0];
	if (collision_seen) {
		MEMS_CUDA(cudaMemsetAsync(sus_count.p, 0, sizeof(uint32_t), c->stream));
		KernelScope ks(c, "suspect_list");
		suspect_list_kernel<<<comp_blocks, 256, 0, c->stream>>>(comp_suspect.p, rec_off.p, n_comp, sus_list.p, sus_cap, sus_count.p);
		MEMS_CUDA(cudaGetLastError());
		c->fetch_async(sus_word.w, sus_count.p, 1);
		sus_word.queued = true;
	}
	if (mode != MEMS_MODE_REPEAT && (uint64_t)n_comp * (uint64_t)(a.max_group + 2) > 0xffffffffull)
		throw Error(MEMS_ERR_UNSUPPORTED, "match list larger than 2^32 values; search fewer sequences per call");
	// every record of a one-problem MemHash / Pairwise call has SeqCount + 2 values: the size is known without asking the device
	const bool fixed_records = mode != MEMS_MODE_REPEAT && a.max_group == a.n_seqs && !many;
	const uint32_t n_flat = fixed_records ? n_comp * (uint32_t)(a.n_seqs + 2) : d2h_u32(c, scalars.p + 5);
	DevBuf<int64_t> d_flat(c, n_flat);
	{
		KernelScope ks(c, "emit");
		emit_kernel<KeyT><<<comp_blocks, 256, 0, c->stream>>>(a, L, hkey, hid, hit_start.p, hit_len.p, comp_rep.p,
		                                                      comp_left.p, comp_right.p, rec_off.p, n_comp, d_flat.p);
		MEMS_CUDA(cudaGetLastError());
	}
	// Components are distinct by construction unless two diagonals shared a hash bucket: a foreign entry between two hits
	// of one diagonal hides them from each other and both may report the same component.  Every component that can be
	// involved was marked on the device (comp_suspect).  For records of one size the few marked records are compared on
	// the host BEFORE the list leaves the device and the duplicates are left out there: taking them out of a list of
	// hundreds of MB on the host (config 5: 633 MB per rank) cost more than the whole device part of the call.
	const int64_t* d_result = d_flat.p;
	uint32_t n_result = n_flat;
	size_t n_drop_device = 0;
	bool deduped_on_device = false;
	DevBuf<int64_t> d_flat_kept;
	if (order != MEMS_ORDER_REFERENCE && collision_seen && fixed_records) {
		MEMS_CUDA(cudaStreamSynchronize(c->stream));  // h_sus_count
		const uint32_t n_sus = h_sus_count, R = (uint32_t)a.n_seqs + 2u;
		if (n_sus <= sus_cap) {
			deduped_on_device = true;
			if (n_sus > 1) {
				DevBuf<int64_t> sus_rec(c, (size_t)n_sus * R);
				{
					KernelScope ks(c, "suspect_records");
					suspect_records_kernel<<<(n_sus * R + 255) / 256, 256, 0, c->stream>>>(d_flat.p, sus_list.p, n_sus, R, sus_rec.p);
					MEMS_CUDA(cudaGetLastError());
				}
				std::vector<uint2> h_list(n_sus);
				std::vector<int64_t> h_rec((size_t)n_sus * R);
				MEMS_CUDA(cudaMemcpyAsync(h_list.data(), sus_list.p, (size_t)n_sus * sizeof(uint2), cudaMemcpyDeviceToHost, c->stream));
				MEMS_CUDA(cudaMemcpyAsync(h_rec.data(), sus_rec.p, h_rec.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
				MEMS_CUDA(cudaStreamSynchronize(c->stream));
				std::vector<uint32_t> idx(n_sus);
				for (uint32_t k = 0; k < n_sus; ++k) idx[k] = k;
				auto rec_of = [&](uint32_t k) { return Rec{h_rec.data() + (size_t)k * R}; };
				std::sort(idx.begin(), idx.end(), [&](uint32_t x, uint32_t y) {  // by record, then by component: the first stays
					return rec_less(rec_of(x), rec_of(y)) || (!rec_less(rec_of(y), rec_of(x)) && h_list[x].x < h_list[y].x);
				});
				std::vector<uint32_t> drops;
				for (uint32_t k = 1; k < n_sus; ++k)
					if (rec_equal(rec_of(idx[k]), rec_of(idx[k - 1]))) drops.push_back(h_list[idx[k]].x);
				if (!drops.empty()) {
					std::sort(drops.begin(), drops.end());
					n_drop_device = drops.size();
					DevBuf<uint32_t> d_drops(c, drops.size());
					MEMS_CUDA(cudaMemcpyAsync(d_drops.p, drops.data(), drops.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
					n_result = n_flat - (uint32_t)drops.size() * R;
					d_flat_kept = DevBuf<int64_t>(c, n_result);
					KernelScope ks(c, "drop_records");
					drop_records_kernel<<<(unsigned)(((uint64_t)n_comp * R + 255) / 256), 256, 0, c->stream>>>(d_flat.p, n_comp, R, d_drops.p,
					                                                                                         (uint32_t)drops.size(), d_flat_kept.p);
					MEMS_CUDA(cudaGetLastError());
					MEMS_CUDA(cudaStreamSynchronize(c->stream));  // `drops` leaves scope
					d_result = d_flat_kept.p;
				}
			}
		}
	}
	// D2H straight into a page-locked buffer that the result object keeps (no pageable staging, no copy)
	out.flat.owner = ctx;
	out.flat.pinned = (int64_t*)c->pinned_get((size_t)n_result * sizeof(int64_t), &out.flat.pinned_cap);
	out.flat.pinned_n = n_result;
	const int64_t* raw = out.flat.pinned;
	// Nothing left for the host to do with the records (any order, no duplicate in doubt): they leave on the copy stream
	// behind the emit kernel and the call returns without waiting for them (FlatRecords::wait).
	const bool async_out = order == MEMS_ORDER_ANY && !many && !c->profiling && (!collision_seen || deduped_on_device);
	if (async_out) {
		if (!c->copy_stream) MEMS_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
		cudaEvent_t emitted = c->get_event();
		MEMS_CUDA(cudaEventRecord(emitted, c->stream));
		MEMS_CUDA(cudaStreamWaitEvent(c->copy_stream, emitted, 0));
		c->event_put(emitted);  // (the wait is already queued; the event may be recorded again)
		MEMS_CUDA(cudaMemcpyAsync(out.flat.pinned, d_result, (size_t)n_result * sizeof(int64_t), cudaMemcpyDeviceToHost, c->copy_stream));
		out.flat.ready = c->get_event();
		MEMS_CUDA(cudaEventRecord(out.flat.ready, c->copy_stream));
		// the source outlives this function: the result object frees it when the copy has arrived
		DevBuf<int64_t>& src = d_result == d_flat.p ? d_flat : d_flat_kept;
		out.flat.dev_keep = src.p;
		src.p = nullptr;
		out.n_matches = n_comp - n_drop_device;
		out.mem_count = out.n_matches;
		out.collisions = out.n_hits - out.n_matches;
		return;
	}
	{
		CopyScope cs(c, "copy_out_matches", (double)n_result * sizeof(int64_t));
		MEMS_CUDA(cudaMemcpyAsync(out.flat.pinned, d_result, (size_t)n_result * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
	}

	if (order != MEMS_ORDER_REFERENCE) {
		MEMS_CUDA(cudaStreamSynchronize(c->stream));
		if (deduped_on_device) {
			if (order == MEMS_ORDER_CANONICAL) {
				std::vector<Rec> recs = split_records(raw, n_result);
				std::sort(recs.begin(), recs.end(), rec_less);
				recs.erase(std::unique(recs.begin(), recs.end(), rec_equal), recs.end());
				out.flat.vec.reserve(n_result);
				for (const Rec& r2 : recs) out.flat.vec.insert(out.flat.vec.end(), r2.p, r2.p + r2.size());
				out.flat.release();
				out.n_matches = recs.size();
			} else {
				out.n_matches = n_comp - n_drop_device;
			}
			out.mem_count = out.n_matches;
			out.collisions = out.n_hits - out.n_matches;
			return;
		}
		// (records of several sizes — many problems per call — and overfull suspect lists: the duplicates are taken out
		// on the host)
		std::vector<char> drop;
		size_t n_drop = 0;
		if (collision_seen) {
			uint32_t n_sus = h_sus_count;
			std::vector<uint2> h_list(n_sus);
			if (n_sus > sus_cap) {  // more than the list holds (never seen outside the forced-collision tests): take them all
				std::vector<uint8_t> h_suspect(n_comp);
				MEMS_CUDA(cudaMemcpyAsync(h_suspect.data(), comp_suspect.p, n_comp, cudaMemcpyDeviceToHost, c->stream));
				MEMS_CUDA(cudaStreamSynchronize(c->stream));
				h_list.clear();
				size_t at = 0;
				for (uint32_t k = 0; k < n_comp; ++k) {
					if (h_suspect[k]) h_list.push_back(make_uint2(k, (uint32_t)at));
					at += (size_t)raw[at] + 2;
				}
			} else if (n_sus) {
				MEMS_CUDA(cudaMemcpyAsync(h_list.data(), sus_list.p, (size_t)n_sus * sizeof(uint2), cudaMemcpyDeviceToHost, c->stream));
				MEMS_CUDA(cudaStreamSynchronize(c->stream));
			}
			std::vector<std::pair<Rec, uint32_t>> sus;  // (record, component index)
			for (const uint2& e : h_list) sus.push_back({Rec{raw + e.y}, e.x});
			std::sort(sus.begin(), sus.end(), [](const std::pair<Rec, uint32_t>& x, const std::pair<Rec, uint32_t>& y) {
				return rec_less(x.first, y.first) || (!rec_less(y.first, x.first) && x.second < y.second);
			});
			drop.assign(n_comp, 0);
			for (size_t k = 1; k < sus.size(); ++k)
				if (rec_equal(sus[k].first, sus[k - 1].first)) {
					drop[sus[k].second] = 1;
					++n_drop;
				}
		}
		if (many) {  // the caller splits the records by problem (find_matches_many)
			many->group.resize(n_comp);
			MEMS_CUDA(cudaMemcpyAsync(many->group.data(), rec_group.p, (size_t)n_comp * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
			MEMS_CUDA(cudaStreamSynchronize(c->stream));
			many->drop = std::move(drop);
			out.n_matches = n_comp - n_drop;
			return;
		}
		if (order == MEMS_ORDER_CANONICAL) {
			std::vector<Rec> recs = split_records(raw, n_flat);
			std::sort(recs.begin(), recs.end(), rec_less);
			recs.erase(std::unique(recs.begin(), recs.end(), rec_equal), recs.end());
			out.flat.vec.reserve(n_flat);
			for (const Rec& r2 : recs) out.flat.vec.insert(out.flat.vec.end(), r2.p, r2.p + r2.size());
			out.flat.release();
			out.n_matches = recs.size();
		} else if (n_drop) {
			out.flat.vec.reserve(n_flat);
			size_t at = 0;
			for (uint32_t k = 0; k < n_comp; ++k) {
				const size_t sz = (size_t)raw[at] + 2;
				if (!drop[k]) out.flat.vec.insert(out.flat.vec.end(), raw + at, raw + at + sz);
				at += sz;
			}
			out.flat.release();
			out.n_matches = n_comp - n_drop;
		} else {
			out.n_matches = n_comp;
		}
		out.mem_count = out.n_matches;
		out.collisions = out.n_hits - out.n_matches;
		return;
	}

	// ---- ORDER_REFERENCE: replay the reference's hash table over (hit, extended match) on the host
	const auto t_replay = std::chrono::steady_clock::now();
	const bool trace = getenv("MEMS_TRACE") != nullptr;
	auto t_mark = std::chrono::steady_clock::now();
	auto mark = [&](const char* what) {
		if (!trace) return;
		cudaStreamSynchronize(c->stream);
		const auto now = std::chrono::steady_clock::now();
		fprintf(stderr, "[mems trace] replay %-24s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_mark).count());
		t_mark = now;
	};
	mark("device stages + records");
	DevBuf<uint32_t> rec_of_hit(c, n_hits), len32(c, n_hits), mem_off(c, n_hits);
	{
		KernelScope ks(c, "hit_record");
		hit_record_kernel<<<hit_blocks, 256, 0, c->stream>>>(hid, is_head.p, seg_of.p, first.p, first_excl.p, n_hits,
		                                                     rec_of_hit.p);
		MEMS_CUDA(cudaGetLastError());
		hit_len_widen_kernel<<<hit_blocks, 256, 0, c->stream>>>(hit_len.p, n_hits, len32.p);
		MEMS_CUDA(cudaGetLastError());
	}
	exclusive_scan_u32(c, len32.p, mem_off.p, n_hits, scalars.p + 5);
	const uint32_t n_mem = d2h_u32(c, scalars.p + 5);
	DevBuf<uint32_t> mem_val(c, n_mem);
	DevBuf<uint8_t> mem_strand(c, n_mem);
	{
		KernelScope ks(c, "gather_members");
		gather_members_kernel<KeyT><<<hit_blocks, 256, 0, c->stream>>>(a, hit_start.p, hit_len.p, mem_off.p, n_hits,
		                                                               mem_val.p, mem_strand.p);
		MEMS_CUDA(cudaGetLastError());
	}
	// (page-locked staging from the context's pool: pageable targets would make these copies several times slower)
	struct Staging {
		Ctx* c;
		void* p = nullptr;
		size_t cap = 0;
		~Staging() {
			if (p) c->pinned_put(p, cap);
		}
	} staging{c};
	const size_t stage_words = (size_t)n_hits * 2 + (size_t)n_mem + ((size_t)n_mem + 3) / 4;
	staging.p = c->pinned_get(stage_words * 4 + 64, &staging.cap);
	uint32_t* const h_rec = static_cast<uint32_t*>(staging.p);
	uint32_t* const h_off = h_rec + n_hits;
	uint32_t* const h_val = h_off + n_hits;
	uint8_t* const h_strand = reinterpret_cast<uint8_t*>(h_val + n_mem);
	MEMS_CUDA(cudaMemcpyAsync(h_rec, rec_of_hit.p, (size_t)n_hits * 4, cudaMemcpyDeviceToHost, c->stream));
	MEMS_CUDA(cudaMemcpyAsync(h_off, mem_off.p, (size_t)n_hits * 4, cudaMemcpyDeviceToHost, c->stream));
	MEMS_CUDA(cudaMemcpyAsync(h_val, mem_val.p, (size_t)n_mem * 4, cudaMemcpyDeviceToHost, c->stream));
	MEMS_CUDA(cudaMemcpyAsync(h_strand, mem_strand.p, (size_t)n_mem, cudaMemcpyDeviceToHost, c->stream));
	MEMS_CUDA(cudaStreamSynchronize(c->stream));

	mark("members to host");
	// emitted records in component order: record r starts at raw[rec_start[r]]
	std::vector<size_t> rec_start;
	rec_start.reserve(n_comp);
	for (size_t i = 0; i < n_flat; i += (size_t)raw[i] + 2) rec_start.push_back(i);

	// the table of this call, or the caller's persistent one (several FindMatches calls into one MemHash table)
	HashTable local_table;
	HashTable& T = persistent ? *persistent : local_table;
	if (T.buckets.empty()) {
		T.size = persistent && persistent->size ? persistent->size : table_size;
		T.buckets.resize(T.size);
	}
	table_size = T.size;
	std::vector<std::vector<Entry*>>& table = T.buckets;
	std::vector<std::unique_ptr<Entry>>& stored = T.stored;
	// A hit only ever touches the bucket its generalized offset selects, so the table is replayed bucket by bucket:
	// hits are grouped by bucket in their original order (stable counting sort) and disjoint bucket ranges go to
	// separate host threads.  The order of operations inside every bucket — all that the reference's result depends
	// on — is unchanged.
	auto build_probe = [&](uint32_t h, Entry& probe, std::vector<int64_t>& probe_start) {
		const uint32_t m0 = h_off[h], m1 = h + 1 < n_hits ? h_off[h + 1] : n_mem;
		probe.seqcount = mode == MEMS_MODE_REPEAT ? (m1 - m0) : (uint32_t)a.n_seqs;
		probe.len = L;
		probe.mersize = L;
		probe_start.assign(probe.seqcount, 0);
		const uint32_t sf = h_strand[m0];
		for (uint32_t j = m0; j < m1; ++j) {
			const uint32_t slot = mode == MEMS_MODE_REPEAT ? (j - m0) : (h_val[j] >> a.pos_bits);
			const int64_t st = (int64_t)(h_val[j] & a.pos_mask) + 1;
			probe_start[slot] = (h_strand[j] != sf) ? -st : st;  // SetDirection, MemHash.cpp:189-203
		}
		probe.start = probe_start.data();
		e_calc_offset(probe);
	};
	const int64_t ts = (int64_t)table_size;
	unsigned n_threads = 1;
	if (n_hits >= 20000) {
		n_threads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
		if (const char* e = getenv("MEMS_HOST_THREADS")) n_threads = (unsigned)std::max(1, atoi(e));
	}
	auto run_parallel = [&](const std::function<void(unsigned)>& body) {
		if (n_threads == 1) {
			body(0);
			return;
		}
		std::vector<std::thread> pool;
		std::vector<std::exception_ptr> errs(n_threads);
		for (unsigned t = 0; t < n_threads; ++t)
			pool.emplace_back([&, t] {
				try {
					body(t);
				} catch (...) {
					errs[t] = std::current_exception();
				}
			});
		for (auto& th : pool) th.join();
		for (auto& e : errs)
			if (e) std::rethrow_exception(e);
	};
	// 1. bucket of every hit
	std::vector<uint32_t> bucket_of(n_hits);
	run_parallel([&](unsigned t) {
		const uint32_t lo = (uint32_t)((uint64_t)n_hits * t / n_threads), hi = (uint32_t)((uint64_t)n_hits * (t + 1) / n_threads);
		Entry probe;
		std::vector<int64_t> probe_start;
		for (uint32_t h = lo; h < hi; ++h) {
			build_probe(h, probe, probe_start);
			bucket_of[h] = (uint32_t)(((probe.offset % ts) + ts) % ts);
		}
	});
	// 2. hits grouped by bucket, original order kept inside a bucket
	std::vector<uint32_t> bucket_first(table_size + 1, 0), by_bucket(n_hits);
	for (uint32_t h = 0; h < n_hits; ++h) ++bucket_first[bucket_of[h] + 1];
	for (uint32_t bkt = 0; bkt < table_size; ++bkt) bucket_first[bkt + 1] += bucket_first[bkt];
	{
		std::vector<uint32_t> at(bucket_first.begin(), bucket_first.end() - 1);
		for (uint32_t h = 0; h < n_hits; ++h) by_bucket[at[bucket_of[h]]++] = h;
	}
	// 3. bucket ranges holding about the same number of hits each
	std::vector<uint32_t> range(n_threads + 1, table_size);
	range[0] = 0;
	for (unsigned t = 1; t < n_threads; ++t) {
		const uint32_t want = (uint32_t)((uint64_t)n_hits * t / n_threads);
		range[t] = (uint32_t)(std::lower_bound(bucket_first.begin(), bucket_first.end(), want) - bucket_first.begin());
		if (range[t] > table_size) range[t] = table_size;
		if (range[t] < range[t - 1]) range[t] = range[t - 1];
	}
	std::vector<std::vector<std::unique_ptr<Entry>>> stored_by(n_threads);
	// a table that lives for this call only keeps no copies: its entries point at the records the device emitted
	std::vector<std::vector<Entry>> arena_by(n_threads);
	std::vector<uint64_t> collisions_by(n_threads, 0), words_by(n_threads, 0), count_by(n_threads, 0);
	run_parallel([&](unsigned t) {
		Entry probe;
		std::vector<int64_t> probe_start;
		if (!persistent) arena_by[t].reserve(bucket_first[range[t + 1]] - bucket_first[range[t]]);  // pointers stay valid
		for (uint32_t bkt = range[t]; bkt < range[t + 1]; ++bkt) {
			std::vector<Entry*>& bucket = table[bkt];
			for (uint32_t k = bucket_first[bkt]; k < bucket_first[bkt + 1]; ++k) {
				const uint32_t h = by_bucket[k];
				build_probe(h, probe, probe_start);
				size_t at = bucket_lower_bound(bucket, probe);
				if (at != bucket.size() && !e_compare(*bucket[at], probe) && !e_compare(probe, *bucket[at])) {
					++collisions_by[t];
					continue;
				}
				// "ExtendMatch": the extended form of this hit is the component the device computed for it
				const int64_t* rec = raw + rec_start[h_rec[h]];
				Entry* e;
				if (persistent) {
					stored_by[t].push_back(std::make_unique<Entry>());
					e = stored_by[t].back().get();
					e->own.assign(rec + 2, rec + 2 + rec[0]);
					e->start = e->own.data();
				} else {
					arena_by[t].emplace_back();
					e = &arena_by[t].back();
					e->start = rec + 2;
				}
				e->seqcount = (uint32_t)rec[0];
				e->len = rec[1];
				e->mersize = 0;  // stored copies lose m_mersize (MatchHashEntry.cpp:118-126)
				e_calc_offset(*e);
				at = bucket_lower_bound(bucket, *e);
				bucket.insert(bucket.begin() + at, e);
				words_by[t] += (uint64_t)e->seqcount + 2;
				++count_by[t];
			}
		}
	});
	for (unsigned t = 0; t < n_threads; ++t) {
		T.collisions += collisions_by[t];
		T.mem_count += count_by[t];
		for (auto& e : stored_by[t]) stored.push_back(std::move(e));
	}
	out.mem_count = T.mem_count;
	out.collisions = T.collisions;
	mark("table replay");
	if (!persistent) {
		// MemHash::GetMatchList (MemHash.h:183-203): buckets in order, front to back — every thread writes the records of
		// its own bucket range into a page-locked buffer of the pool (no copies of copies, no fresh pages to fault in)
		uint64_t total = 0, n_out = 0;
		std::vector<uint64_t> word_at(n_threads, 0);
		for (unsigned t = 0; t < n_threads; ++t) {
			word_at[t] = total;
			total += words_by[t];
			n_out += count_by[t];
		}
		size_t out_cap = 0;
		int64_t* dst = (int64_t*)c->pinned_get((size_t)total * sizeof(int64_t) + 64, &out_cap);
		run_parallel([&](unsigned t) {
			int64_t* w = dst + word_at[t];
			for (uint32_t bkt = range[t]; bkt < range[t + 1]; ++bkt)
				for (const Entry* e : table[bkt]) {
					*w++ = e->seqcount;
					*w++ = e->len;
					memcpy(w, e->start, sizeof(int64_t) * e->seqcount);
					w += e->seqcount;
				}
		});
		out.flat.release();  // the device's records (raw) are no longer needed
		out.flat.pinned = dst;
		out.flat.pinned_cap = out_cap;
		out.flat.pinned_n = (size_t)total;
		out.n_matches = n_out;
		mark("output list");
		out.host_replay_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_replay).count();
		return;
	}
	// MemHash::GetMatchList (MemHash.h:183-203): buckets in order, front to back
	{
		size_t total = 0;
		for (auto& e : stored) total += (size_t)e->seqcount + 2;
		out.flat.vec.reserve(total);
	}
	for (auto& bucket : table)
		for (Entry* e : bucket) {
			out.flat.vec.push_back(e->seqcount);
			out.flat.vec.push_back(e->len);
			out.flat.vec.insert(out.flat.vec.end(), e->start, e->start + e->seqcount);
		}
	out.flat.release();
	out.n_matches = stored.size();
	mark("output list");
	out.host_replay_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_replay).count();
}

static void emit_table(const HashTable& T, MatchResult& out) {
	for (auto& bucket : T.buckets)
		for (const TableEntry* e : bucket) {
			out.flat.vec.push_back(e->seqcount);
			out.flat.vec.push_back(e->len);
			out.flat.vec.insert(out.flat.vec.end(), e->start, e->start + e->seqcount);
		}
	out.n_matches = T.stored.size();
	out.mem_count = T.mem_count;
	out.collisions = T.collisions;
}

// ---- MemHash::FindMatchesFromPosition: sorted mer list g takes part from its entry start_points[g] on
// (MatchFinder.cpp:137-164 reads SML g from that index).  The union is a stable merge of the per-sequence lists, so
// "index >= start" of a sequence's list is "(key, position) >= the list's entry at start": one threshold per sequence.
template <class KeyT>
__global__ void start_thresholds_kernel(const uint32_t* __restrict__ positions, const KeyT* __restrict__ key_pos,
                                        const SeqMeta* __restrict__ meta, const uint64_t* __restrict__ start, int n_seqs,
                                        uint32_t pos_mask, KeyT* __restrict__ thr_key, uint32_t* __restrict__ thr_val,
                                        uint8_t* __restrict__ thr_mode) {
	const int g = threadIdx.x;
	if (g >= n_seqs) return;
	const SeqMeta m = meta[g];
	const uint64_t s = start[g];
	if (s == 0) {
		thr_mode[g] = 0;  // everything
	} else if (s >= m.n_seeds) {
		thr_mode[g] = 2;  // nothing
	} else {
		const uint32_t v = positions[m.seed_off + s];
		thr_mode[g] = 1;
		thr_val[g] = v;
		thr_key[g] = key_pos[m.seed_off + (v & pos_mask)];
	}
}

template <class KeyT>
__global__ void start_keep_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t n, int pos_bits,
                                  const KeyT* __restrict__ thr_key, const uint32_t* __restrict__ thr_val,
                                  const uint8_t* __restrict__ thr_mode, uint32_t* __restrict__ keep) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t v = vals[i], g = v >> pos_bits;
	const uint8_t mode = thr_mode[g];
	const KeyT k = keys[i];
	keep[i] = mode == 0 || (mode == 1 && (k > thr_key[g] || (k == thr_key[g] && v >= thr_val[g]))) ? 1u : 0u;
}

template <class KeyT>
__global__ void start_compact_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t n,
                                     const uint32_t* __restrict__ keep, const uint32_t* __restrict__ at,
                                     KeyT* __restrict__ out_keys, uint32_t* __restrict__ out_vals) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && keep[i]) {
		out_keys[at[i]] = keys[i];
		out_vals[at[i]] = vals[i];
	}
}

template <class KeyT>
static void find_matches_typed(Batch& b, int mode, int order, uint32_t table_size, uint64_t seq_mask, MatchResult& out,
                               HashTable* persistent, const uint64_t* start_points, ManyOut* many = nullptr) {
	Ctx* c = b.ctx.get();
	out.seq_count = (uint32_t)b.n_seqs;
	out.seed_length = (uint32_t)b.sd.L;
	if (b.n_total < 2) {
		if (persistent) emit_table(*persistent, out);
		return;
	}
	MatchArgs a;
	a.keys = b.keys.p;
	a.vals = b.vals.p;
	a.n = (uint32_t)b.n_total;
	a.pos_bits = b.pos_bits;
	a.pos_mask = b.pos_mask();
	a.n_seqs = b.n_seqs;
	a.max_group = b.max_group;
	a.mode = mode;
	// the reference's "match number" puts sequence 0 in the most significant of n_seqs bits; here bit g = sequence g
	a.hit_key_bits = b.pos_bits <= 26 ? 56 : 64;
	a.test_hash_bits = c->test_hash_bits;
	a.warp_budget = c->test_walk_budget ? c->test_walk_budget : kWarpProbeBudget;
	a.cta_budget = c->test_walk_budget ? c->test_walk_budget : kCtaRoundBudget;
	a.seq_set = 0;
	for (int g = 0; g < b.n_seqs; ++g)
		if ((seq_mask >> (b.n_seqs - 1 - g)) & 1) a.seq_set |= 1ull << g;
	a.planes = b.planes.p;
	a.meta = b.d_meta.p;
	DevBuf<KeyT> fkeys;
	DevBuf<uint32_t> fvals;
	bool from_position = false;
	for (int g = 0; start_points && g < b.n_seqs; ++g) from_position = from_position || start_points[g] != 0;
	if (from_position) {
		const uint32_t* positions = b.sorted_positions();
		DevBuf<uint64_t> d_start(c, b.n_seqs);
		DevBuf<KeyT> thr_key(c, b.n_seqs);
		DevBuf<uint32_t> thr_val(c, b.n_seqs), keep(c, a.n), at(c, a.n), total(c, 1);
		DevBuf<uint8_t> thr_mode(c, b.n_seqs);
		MEMS_CUDA(cudaMemcpyAsync(d_start.p, start_points, sizeof(uint64_t) * b.n_seqs, cudaMemcpyHostToDevice, c->stream));
		KernelScope ks(c, "start_points");
		start_thresholds_kernel<KeyT><<<1, 256, 0, c->stream>>>(positions, reinterpret_cast<const KeyT*>(b.keys_by_pos.p), b.d_meta.p,
		                                                          d_start.p, b.n_seqs, a.pos_mask, thr_key.p, thr_val.p, thr_mode.p);
		MEMS_CUDA(cudaGetLastError());
		const uint32_t blocks = (a.n + 255) / 256;
		start_keep_kernel<KeyT><<<blocks, 256, 0, c->stream>>>(reinterpret_cast<const KeyT*>(a.keys), a.vals, a.n, a.pos_bits, thr_key.p,
		                                                       thr_val.p, thr_mode.p, keep.p);
		MEMS_CUDA(cudaGetLastError());
		exclusive_scan_u32(c, keep.p, at.p, a.n, total.p);
		const uint32_t n_keep = d2h_u32(c, total.p);
		fkeys = DevBuf<KeyT>(c, n_keep);
		fvals = DevBuf<uint32_t>(c, n_keep);
		start_compact_kernel<KeyT><<<blocks, 256, 0, c->stream>>>(reinterpret_cast<const KeyT*>(a.keys), a.vals, a.n, keep.p, at.p,
		                                                          fkeys.p, fvals.p);
		MEMS_CUDA(cudaGetLastError());
		MEMS_CUDA(cudaStreamSynchronize(c->stream));  // d_start was filled from caller memory; the scratch dies with this scope
		a.keys = fkeys.p;
		a.vals = fvals.p;
		a.n = n_keep;
		if (n_keep < 2) {
			if (persistent) emit_table(*persistent, out);
			return;
		}
	}
	HitSet hits;
	DevBuf<uint32_t> pvals;
	DevBuf<KeyT> pkeys;
	if (mode == MEMS_MODE_PAIRWISE) {
		find_pair_hits<KeyT>(c, a, hits, pvals, pkeys);
		a.keys = pkeys.p;  // from here on a hit's members are the two entries of its pair
		a.vals = pvals.p;
		a.n = 2 * hits.n;
	} else {
		find_hits<KeyT>(c, a, hits);
	}
	out.max_run = hits.max_run;
	out.n_hits = hits.n;
	if (hits.n == 0) {
		if (persistent) emit_table(*persistent, out);  // nothing new: the result is still the whole table
		return;
	}
	extend_hits<KeyT>(b.ctx, a, b.sd, hits, order, table_size, out, persistent, nullptr, many);
}

void find_matches_on_batch(Batch& b, int mode, int order, uint32_t table_size, uint64_t seq_mask, MatchResult& out,
                           HashTable* persistent, const uint64_t* start_points) {
	if (persistent && order != MEMS_ORDER_REFERENCE) throw Error(MEMS_ERR_INVALID, "a persistent table needs MEMS_ORDER_REFERENCE");
	if (seq_mask && mode != MEMS_MODE_MEMHASH) throw Error(MEMS_ERR_INVALID, "seq_mask applies to MEMS_MODE_MEMHASH only");
	if (b.n_seqs > MEMS_MAX_SEQS) throw Error(MEMS_ERR_UNSUPPORTED, "more than MEMS_MAX_SEQS sequences in one match-finding call");
	if (b.n_groups > 1) throw Error(MEMS_ERR_INVALID, "a batch of several problems is searched with mems_find_matches_many");
	if (b.key64)
		find_matches_typed<uint64_t>(b, mode, order, table_size, seq_mask, out, persistent, start_points);
	else
		find_matches_typed<uint32_t>(b, mode, order, table_size, seq_mask, out, persistent, start_points);
}

// ---- many small problems in one launch set (the gap re-anchoring callers, ProgressiveAligner.cpp:589-678 under
// `omp parallel for` at :695): the batch holds the sequences of all problems, the problem number sits above the key
// bits, so one sort, one run scan and one extension serve them all; the records are split by problem on the host.
void find_matches_many(Batch& b, int mode, int order, std::vector<MatchResult>& out) {
	if (mode != MEMS_MODE_MEMHASH && mode != MEMS_MODE_PAIRWISE)
		throw Error(MEMS_ERR_UNSUPPORTED, "mems_find_matches_many: MEMS_MODE_MEMHASH or MEMS_MODE_PAIRWISE");
	if (order == MEMS_ORDER_REFERENCE)
		throw Error(MEMS_ERR_UNSUPPORTED, "mems_find_matches_many: MEMS_ORDER_ANY or MEMS_ORDER_CANONICAL (the reference's table order needs one table per problem)");
	if (b.max_group > MEMS_MAX_SEQS) throw Error(MEMS_ERR_UNSUPPORTED, "more than MEMS_MAX_SEQS sequences in one problem");
	out.clear();
	out.resize(b.n_groups);
	for (int g = 0; g < b.n_groups; ++g) out[g].seed_length = (uint32_t)b.sd.L;
	for (int i = 0; i < b.n_seqs; ++i) out[b.meta[i].group].seq_count = b.meta[i].group_count;
	MatchResult all;
	ManyOut many;
	if (b.key64)
		find_matches_typed<uint64_t>(b, mode, MEMS_ORDER_ANY, 40000u, 0, all, nullptr, nullptr, &many);
	else
		find_matches_typed<uint32_t>(b, mode, MEMS_ORDER_ANY, 40000u, 0, all, nullptr, nullptr, &many);
	const int64_t* raw = all.flat.data();
	const size_t n_flat = all.flat.size();
	size_t at = 0;
	for (size_t k = 0; at < n_flat; ++k) {
		const size_t sz = (size_t)raw[at] + 2;
		if (many.drop.empty() || !many.drop[k]) {
			MatchResult& r = out[many.group[k]];
			r.flat.vec.insert(r.flat.vec.end(), raw + at, raw + at + sz);
			++r.n_matches;
		}
		at += sz;
	}
	for (MatchResult& r : out) {
		if (order == MEMS_ORDER_CANONICAL && r.n_matches > 1) {
			std::vector<Rec> recs = split_records(r.flat.vec.data(), r.flat.vec.size());
			std::sort(recs.begin(), recs.end(), rec_less);
			std::vector<int64_t> sorted;
			sorted.reserve(r.flat.vec.size());
			for (const Rec& x : recs) sorted.insert(sorted.end(), x.p, x.p + x.size());
			r.flat.vec.swap(sorted);
		}
		r.mem_count = r.n_matches;
		r.max_run = all.max_run;  // (of the whole batch)
	}
}

bool table_add_entry(HashTable& T, uint32_t seq_count, int64_t length, const int64_t* starts, int64_t mersize) {
	if (T.buckets.empty()) {
		if (!T.size) T.size = 40000u;
		T.buckets.resize(T.size);
	}
	auto e = std::make_unique<Entry>();
	e->seqcount = seq_count;
	e->len = length;
	e->mersize = mersize;
	e->own.assign(starts, starts + seq_count);
	e->start = e->own.data();
	e_calc_offset(*e);
	const int64_t ts = (int64_t)T.size;
	std::vector<Entry*>& bucket = T.buckets[(size_t)(((e->offset % ts) + ts) % ts)];
	size_t at = bucket_lower_bound(bucket, *e);
	if (at != bucket.size() && !e_compare(*bucket[at], *e) && !e_compare(*e, *bucket[at])) {
		++T.collisions;
		return false;
	}
	// (the reference would call ExtendMatch here: a no-op for LoadFile, which runs before any sequence is added)
	e->mersize = 0;  // stored copies lose m_mersize (MatchHashEntry.cpp:118-126)
	at = bucket_lower_bound(bucket, *e);
	bucket.insert(bucket.begin() + at, e.get());
	T.stored.push_back(std::move(e));
	++T.mem_count;
	return true;
}

void table_list(const HashTable& T, MatchResult& out) { emit_table(T, out); }

// ================================================================================================ sharded (multi-GPU)
// One process per GPU.  The path shards in two exchanges (ParallelMemHash.cpp:42-121 is the
// reference's own, never-built, seed-range chunking):
//   * every rank packs + extracts a contiguous block of the sequences;
//   * seed space is range-partitioned on the top 8 key bits with owners balanced by the global histogram
//     (canonical keys are skewed ~15:1 across the key space), so one all-to-all delivers every occurrence
//     of a seed range to one GPU, which sorts it and finds the hits of its range;
//   * hits are then re-partitioned by the hash of their DIAGONAL (second, much smaller all-to-all), so that
//     all hits of a diagonal meet on one GPU: segments, walks and components are disjoint across ranks — no
//     duplicated extension work and nothing to de-duplicate between ranks;
//   * the position-ordered keys of all sequences are all-gathered (window tests read arbitrary sequences).
// Each rank returns its share of the (distinct) matches.
void shard_sequence_range(int n_seqs, int rank, int world, int* first, int* count) {
	const int base = n_seqs / world, extra = n_seqs % world;
	*first = rank * base + (rank < extra ? rank : extra);
	*count = base + (rank < extra ? 1 : 0);
}

void shard_bucket_owners(const uint64_t* hist256, int world, uint8_t* owner256) {
	uint64_t total = 0;
	for (int b = 0; b < 256; ++b) total += hist256[b];
	uint64_t acc = 0;
	int r = 0;
	if (total == 0) world = 1;  // nothing to balance: everything stays with rank 0
	for (int b = 0; b < 256; ++b) {
		// bucket b goes to the rank whose share [r*total/world, (r+1)*total/world) holds the bucket's midpoint
		const uint64_t mid = acc + hist256[b] / 2;
		while (r + 1 < world && mid >= (uint64_t)((__uint128_t)total * (uint64_t)(r + 1) / (uint64_t)world)) ++r;
		owner256[b] = (uint8_t)r;
		acc += hist256[b];
	}
}

// The record exchange as every rank derives it from the gathered top-digit histograms (hist_all[q * 256 + b] =
// records of bucket b on rank q): counts[q * world + p] = records rank q sends to rank p; for THIS rank,
// src_elem[p] = first element of its slice for rank p in its partitioned order (buckets of one owner are
// contiguous) and dst_elem[p] = element at which that slice starts inside rank p's receive region (slices lie in
// sender order); *max_recv = the largest receive region, which sizes the exchange windows identically everywhere.
void shard_exchange_plan(const uint32_t* hist_all, int world, int rank, const uint8_t* owner256, uint64_t* counts,
                         uint64_t* src_elem, uint64_t* dst_elem, uint64_t* max_recv) {
	for (int i = 0; i < world * world; ++i) counts[i] = 0;
	for (int q = 0; q < world; ++q)
		for (int b = 0; b < 256; ++b) counts[(size_t)q * world + owner256[b]] += hist_all[(size_t)256 * q + b];
	uint64_t mx = 0, at = 0;
	for (int p = 0; p < world; ++p) {
		uint64_t n = 0, before = 0;
		for (int q = 0; q < world; ++q) {
			if (q < rank) before += counts[(size_t)q * world + p];
			n += counts[(size_t)q * world + p];
		}
		mx = n > mx ? n : mx;
		src_elem[p] = at;
		dst_elem[p] = before;
		at += counts[(size_t)rank * world + p];
	}
	*max_recv = mx;
}

constexpr uint32_t kMemberStrandBit = 0x80000000u;  // exchanged members: (sequence << pos_bits | position) | strand << 31

// The hits in destination (diagonal-hash) order, packed for the exchange: length word of every hit (with the
// first-member strand flag hit_describe left in it) and its members as ONE word each — the union value with the
// strand in bit 31 — in union order (strand-0 members first, each part ascending).  A warp takes 32 consecutive
// hits; their member lists leave one after the other as contiguous pieces.
template <class KeyT>
__global__ void __launch_bounds__(256)
pack_hits_kernel(MatchArgs a, const uint32_t* __restrict__ hid, const uint32_t* __restrict__ hit_start,
                 const uint16_t* __restrict__ hit_len, const uint32_t* __restrict__ mem_off, uint32_t n_hits,
                 uint16_t* __restrict__ out_len, uint32_t* __restrict__ out_mem) {
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t i0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32u;
	if (i0 >= n_hits) return;
	const uint32_t i = i0 + lane;
	uint32_t s = 0, len = 0, off = 0;
	if (i < n_hits) {
		const uint32_t h = hid[i];
		const uint16_t l = hit_len[h];
		s = hit_start[h];
		len = l & ~kFirstStrandBit;
		off = mem_off[i];
		out_len[i] = l;
	}
	// (lanes past n_hits hold len = 0: their turns copy nothing)  Four hits per turn, eight lanes each (most hits have at
	// most eight members), all eight turns' loads in flight at once; the rare longer lists follow in the same shape.
	const uint32_t sub = lane & 7u, grp = lane >> 3;
#pragma unroll
	for (uint32_t k = 0; k < 8; ++k) {
		const uint32_t src = k * 4u + grp;
		const uint32_t sk = __shfl_sync(0xffffffffu, s, src), lk = __shfl_sync(0xffffffffu, len, src);
		const uint32_t ok = __shfl_sync(0xffffffffu, off, src);
		if (sub < lk) out_mem[ok + sub] = a.vals[sk + sub] | (strand_of<KeyT>(a.keys, sk + sub) << 31);
	}
	if (__any_sync(0xffffffffu, len > 8u)) {
		for (uint32_t k = 0; k < 8; ++k) {
			const uint32_t src = k * 4u + grp;
			const uint32_t sk = __shfl_sync(0xffffffffu, s, src), lk = __shfl_sync(0xffffffffu, len, src);
			const uint32_t ok = __shfl_sync(0xffffffffu, off, src);
			for (uint32_t t = 8u + sub; t < lk; t += 8u) out_mem[ok + t] = a.vals[sk + t] | (strand_of<KeyT>(a.keys, sk + t) << 31);
		}
	}
}

__global__ void sorted_len_kernel(const uint32_t* __restrict__ hid, const uint16_t* __restrict__ hit_len, uint32_t n_hits,
                                  uint32_t* __restrict__ out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n_hits) out[i] = hit_len[hid[i]] & ~kFirstStrandBit;
}

// What this rank sends to every rank d: hits [bound[d], bound[d+1]) of the hit keys grouped by their top 8 hash bits
// (rank d owns the buckets [ceil(256 d / world), ceil(256 (d+1) / world))) and the members of those hits.
// counts[d] = hits, counts[world + d] = members.
__global__ void hit_send_counts_kernel(const uint64_t* __restrict__ hkey, int top_shift, const uint32_t* __restrict__ mem_off,
                                       uint32_t n_hits, const uint32_t* __restrict__ last_len, int world,
                                       uint64_t* __restrict__ counts) {
	__shared__ uint32_t s_bound[257], s_mbound[257];
	const int d = threadIdx.x;
	if (d <= world) {
		uint32_t lo = n_hits;
		if (d < world) {
			const uint64_t t = (uint64_t)((256 * d + world - 1) / world);  // first bucket of rank d
			uint32_t hi = n_hits;
			lo = 0;
			while (lo < hi) {
				const uint32_t mid = (lo + hi) / 2;
				if ((hkey[mid] >> top_shift) < t) lo = mid + 1;
				else hi = mid;
			}
		}
		s_bound[d] = lo;
		// members before hit `lo`: its offset, or the total (offset of the last hit + its length) at the end
		s_mbound[d] = lo < n_hits ? mem_off[lo] : (n_hits ? mem_off[n_hits - 1] + *last_len : 0u);
	}
	__syncthreads();
	if (d < world) {
		counts[d] = s_bound[d + 1] - s_bound[d];
		counts[world + d] = s_mbound[d + 1] - s_mbound[d];
	}
}

// length of the last hit in sorted order (with mem_off of that hit: the member total)
__global__ void last_len_kernel(const uint32_t* __restrict__ slen, uint32_t n_hits, uint32_t* __restrict__ out) {
	if (threadIdx.x == 0 && blockIdx.x == 0) *out = n_hits ? slen[n_hits - 1] : 0u;
}

// received hits: length words -> u32 lengths (scanned into the hits' first member index) + the u16 words the
// extension reads; received members -> union value and a key array that carries just the strand bit
template <class KeyT>
__global__ void received_hits_kernel(const uint16_t* __restrict__ len16, uint32_t* __restrict__ mem, uint32_t n_hits, uint32_t n_mem,
                                     uint32_t* __restrict__ len32, uint16_t* __restrict__ hit_len, KeyT* __restrict__ keys) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n_hits) {
		const uint16_t l = len16[i];
		hit_len[i] = l;
		len32[i] = l & ~kFirstStrandBit;
	}
	if (i < n_mem) {
		const uint32_t v = mem[i];
		keys[i] = (KeyT)(v >> 31);  // the kernels downstream read only the strand bit of a member's key
		mem[i] = v & ~kMemberStrandBit;
	}
}

// A failure on one rank (a '-' in its sequences, a limit exceeded) must fail the call on EVERY rank, or the others
// would wait forever in the next collective: each collective step first agrees on a status word.
static void throw_agreed(int code, int rank_of_error, const char* what) {
	throw Error(code, std::string(what) + " (reported by rank " + std::to_string(rank_of_error) + "; every rank of the sharded call fails together)");
}

template <class KeyT>
static void find_matches_sharded_typed(std::shared_ptr<Ctx> ctx, Comm* comm, const SeedDesc& sd, int n_seqs,
                                       const char* const* seqs, const uint64_t* lens, uint64_t seed, int mode, int order,
                                       MatchResult& out) {
	Ctx* c = ctx.get();
	const int W = Comm_world(comm), R = Comm_rank(comm);
	const size_t K = sizeof(KeyT);
	const bool trace = getenv("MEMS_TRACE") != nullptr;
	auto t_last = std::chrono::steady_clock::now();
	auto mark = [&](const char* what) {
		if (!trace) return;
		cudaStreamSynchronize(c->stream);
		auto now = std::chrono::steady_clock::now();
		const double ms = std::chrono::duration<double, std::milli>(now - t_last).count();
		if (R == 0 || ms > 6.0) fprintf(stderr, "[mems trace r%d] %-28s %8.3f ms\n", R, what, ms);
		t_last = now;
	};
	// ---- global layout (identical on every rank; argument errors are therefore the same everywhere)
	std::vector<SeqMeta> gmeta(n_seqs);
	uint64_t seed_off = 0, word_off = kLeadWords;  // the same packed layout as one batch of all sequences would have
	uint32_t max_seeds = 0;
	for (int g = 0; g < n_seqs; ++g) {
		if (lens[g] > 0xffffffffull) throw Error(MEMS_ERR_UNSUPPORTED, "sequence longer than 2^32-1 bases");
		SeqMeta& m = gmeta[g];
		memset(&m, 0, sizeof m);
		m.n_bases = (uint32_t)lens[g];
		m.n_seeds = lens[g] >= (uint64_t)sd.L ? (uint32_t)(lens[g] - sd.L + 1) : 0u;
		m.seed_off = seed_off;
		m.word_off = word_off;
		m.tag = (uint32_t)g;
		m.group_count = (uint32_t)n_seqs;
		seed_off += m.n_seeds;
		word_off += seq_packed_words(lens[g]);
		max_seeds = std::max(max_seeds, m.n_seeds);
	}
	const uint64_t words_total = word_off + kTailWords;
	const uint64_t s_total = seed_off;
	const int pos_bits = bits_for(max_seeds ? max_seeds - 1 : 0);
	const int seq_bits = n_seqs > 1 ? bits_for((uint64_t)n_seqs - 1) : 0;
	if (pos_bits + seq_bits > 31)
		throw Error(MEMS_ERR_UNSUPPORTED, "sequence count x longest sequence exceeds 31 bits (the exchanged hit members carry their strand in bit 31)");
	if (s_total >= (1ull << 31)) throw Error(MEMS_ERR_UNSUPPORTED, "more than 2^31-1 seed positions in total");
	out.seq_count = (uint32_t)n_seqs;
	out.seed_length = (uint32_t)sd.L;

	// ---- 1. this rank's block of sequences: pack, extract, top-digit histogram.  Local failures are collected in
	// `my_error` and agreed on with the histogram all-gather below.
	int first, count;
	shard_sequence_range(n_seqs, R, W, &first, &count);
	int my_error = MEMS_OK;
	std::string my_error_text;
	std::shared_ptr<Batch> local;
	try {
		for (int g = first; g < first + count; ++g)
			if (lens[g] && !seqs[g]) throw Error(MEMS_ERR_INVALID, "a sequence of this rank's block is missing");
		local = prepare_batch_from_ascii(ctx, count, seqs + first, lens + first, seed, (uint32_t)first, pos_bits, seq_bits ? seq_bits : 0);
	} catch (const Error& e) {
		if (e.code == MEMS_ERR_CUDA || e.code == MEMS_ERR_NCCL) throw;  // the device is gone: nothing left to agree with
		my_error = e.code;
		my_error_text = e.what();
		local = prepare_batch_from_ascii(ctx, 0, seqs, lens, seed, 0, pos_bits, seq_bits ? seq_bits : 0);  // an empty block
		count = 0;
	}
	const uint64_t n_loc = local->n_total;
	SortPlan top;  // one digit: the top 8 key bits
	top.n_passes = 1;
	// the top bits of the MASKED key (never the strand flag in bit 0: both strands of a seed belong to one owner)
	top.bits[0] = sd.key_bits - 1 < 8 ? sd.key_bits - 1 : 8;
	top.shift[0] = sd.key_bits - top.bits[0];
	DevBuf<uint8_t> keys_loc(c, n_loc * K);
	DevBuf<uint32_t> vals_loc(c, n_loc), hist_top(c, 256 + 2);  // [256], [257] (as one u64 slot): this rank's status word
	MEMS_CUDA(cudaMemsetAsync(hist_top.p, 0, 258 * sizeof(uint32_t), c->stream));
	if (n_loc)
		launch_extract(c, local->packed.p, local->d_meta.p, local->meta.data(), count, sd, pos_bits, K == 8, keys_loc.p,
		               vals_loc.p, hist_top.p, 1, top.shift, top.bits);
	DevBuf<uint2> planes_own;  // (declared before the guard: released only after the side stream is done with it)
	const uint2* planes_all_p = nullptr;
	struct SideGuard {  // whatever way this function is left, the side stream must be done with local->planes / planes_all
		Comm* c;
		~SideGuard() { comm_side_synchronize(c); }
	} side_guard{comm};
	// the pack has long finished by now: a '-' in this rank's sequences joins the status word
	if (my_error == MEMS_OK && local->take_gap_flag()) {
		my_error = MEMS_ERR_GAP;
		my_error_text = "Gap in genome sequence: input sequences must be unaligned and ungapped (SortedMerList.cpp:433-437)";
	}
	{
		const uint32_t status[2] = {(uint32_t)my_error, 0u};
		MEMS_CUDA(cudaMemcpyAsync(hist_top.p + 256, status, sizeof status, cudaMemcpyHostToDevice, c->stream));
	}
	// every rank's top-digit histogram (+ status word) to every rank in one all-gather: the sum gives the owners of the
	// key ranges, row p gives what rank p will send here — no separate all-reduce and no count exchange, one host sync
	DevBuf<uint32_t> hist_all(c, (size_t)258 * W);
	comm_all_gather_u64(comm, reinterpret_cast<const uint64_t*>(hist_top.p), reinterpret_cast<uint64_t*>(hist_all.p), 129);
	std::vector<uint32_t> h_raw((size_t)258 * W), h_all((size_t)256 * W);
	c->fetch(h_raw.data(), hist_all.p, h_raw.size() * sizeof(uint32_t));
	for (int p = 0; p < W; ++p) {
		const int code = (int)h_raw[(size_t)258 * p + 256];
		if (code != MEMS_OK) {
			if (p == R) throw Error(my_error, my_error_text);
			throw_agreed(code, p, code == MEMS_ERR_GAP ? "Gap in genome sequence on another rank" : "another rank failed before the exchange");
		}
		memcpy(&h_all[(size_t)256 * p], &h_raw[(size_t)258 * p], 256 * sizeof(uint32_t));
	}
	const uint32_t* h_hist32 = h_all.data() + (size_t)256 * R;
	uint64_t g_hist[256];
	for (int b = 0; b < 256; ++b) {
		g_hist[b] = 0;
		for (int p = 0; p < W; ++p) g_hist[b] += h_all[(size_t)256 * p + b];
	}
	// ---- the packed sequences (as bit planes, 0.25 B per base) of ALL sequences on every rank: the window tests of the
	// extension read any sequence.  Started on the side stream, needed only at the end; written straight into every
	// peer's window over NVLink (NCCL send/recv where CUDA IPC is not available).
	// Only now: a peer may write this rank's window once this rank has left the previous call, and the histogram
	// all-gather above is the first point that proves it.
	const bool direct_planes = !getenv("MEMS_NO_PEER_WINDOWS") && comm_window_reserve(comm, 2, words_total * 4);
	{
		std::vector<uint64_t> bytes(W), offs(W);
		for (int p = 0; p < W; ++p) {
			int f, n;
			shard_sequence_range(n_seqs, p, W, &f, &n);
			uint64_t words = 0;
			for (int g = f; g < f + n; ++g) words += seq_packed_words(lens[g]);
			bytes[p] = words * 4;
			offs[p] = (n ? gmeta[f].word_off : kLeadWords) * 4;
		}
		// this rank's block has the global layout shifted to start behind its own lead pad
		const uint2* mine = count ? local->planes.p + kLeadWords / 2 : nullptr;
		KernelScope ks(c, direct_planes ? "peer_all_gather_planes" : "nccl_all_gather_planes", (double)words_total * 4);
		uint2* dst;
		if (direct_planes) {
			dst = static_cast<uint2*>(comm_window_local(comm, 2));
			comm_window_all_gather(comm, 2, mine, bytes[R], offs[R]);
		} else {
			planes_own = DevBuf<uint2>(c, words_total / 2);
			dst = planes_own.p;
			comm_all_gather_v(comm, mine, dst, bytes.data(), offs.data());
		}
		// lead and tail pad: chunks of the window test may reach into them (their content never decides a window)
		MEMS_CUDA(cudaMemsetAsync(dst, 0, kLeadWords * 4, c->stream));
		MEMS_CUDA(cudaMemsetAsync(reinterpret_cast<uint32_t*>(dst) + (words_total - kTailWords), 0, kTailWords * 4, c->stream));
		planes_all_p = dst;
	}

	mark("pack+extract+histogram");
	// ---- 2. owners of the key ranges; every rank derives the whole count matrix from the gathered histograms
	uint8_t owner[256];
	shard_bucket_owners(g_hist, W, owner);
	std::vector<uint64_t> send_counts(W, 0), recv_counts(W, 0);
	for (int b = 0; b < 256; ++b) send_counts[owner[b]] += h_hist32[b];
	std::vector<uint64_t> rec_counts((size_t)W * W, 0);  // [sender][receiver], identical on every rank
	std::vector<uint64_t> slice_src(W), slice_dst(W);
	uint64_t max_recv = 0, n_recv = 0;
	shard_exchange_plan(h_all.data(), W, R, owner, rec_counts.data(), slice_src.data(), slice_dst.data(), &max_recv);
	for (int p = 0; p < W; ++p) {
		recv_counts[p] = rec_counts[(size_t)p * W + R];
		n_recv += recv_counts[p];
	}
	// a limit every rank evaluates identically (max_recv is the largest receive region of ANY rank): all throw together
	if (max_recv > radix_max_items()) throw Error(MEMS_ERR_UNSUPPORTED, "more than 2^30-1 seed records on one rank");
	auto align256 = [](size_t x) { return (x + 255) & ~(size_t)255; };
	// ---- 3. partition the local records by top digit and deliver every key range to its owner.
	// With exchange windows (CUDA IPC mappings of the peers' receive buffers) the delivery is peer-to-peer over
	// NVLink, in one of two forms:
	//   copy    (default) partition locally, then one DMA copy per peer straight into its window + a barrier;
	//   scatter (MEMS_PEER_SCATTER=1) ONE kernel: the counting pass scatters each digit's run into the owner's window,
	//           i.e. the partition pass is the send side of the all-to-all.
	// Measured on 2 x B200, 40 M records per rank: scatter 0.90 ms, partition 0.27 + copies 0.35 = 0.62 ms — a tile
	// leaves only ~64 B per digit, and NVLink moves such small writes far below its bulk rate, so the copy form is
	// the default.  Without IPC: partition locally, then NCCL send/recv (1.4 ms for the same exchange).
	const size_t rec_region[2] = {0, align256(max_recv * K)};
	const bool direct = !getenv("MEMS_NO_PEER_WINDOWS") && comm_window_reserve(comm, 0, rec_region[1] + align256(max_recv * 4));
	const bool scatter = direct && getenv("MEMS_PEER_SCATTER") != nullptr;
	// (all buffers of the exchanges live to the end of the call: nothing has to wait for the stream just to free them)
	DevBuf<uint8_t> rk_own, rk_b(c, n_recv * K), keys_part;
	DevBuf<uint32_t> rv_own, rv_b(c, n_recv), vals_part;
	DevBuf<uint64_t> d_dst;
	uint64_t h_dst[512];
	uint8_t* rk_a_p;
	uint32_t* rv_a_p;
	if (scatter) {
		// digit b's run starts at index first_idx[b] of this rank's partitioned order and belongs at element
		// (records of lower ranks for that owner) + (index inside this rank's slice for that owner) of the owner's region
		for (int b = 0; b < 256; ++b) {
			const int p = owner[b];
			const int64_t shift_elems = (int64_t)slice_dst[p] - (int64_t)slice_src[p];  // destination index - partitioned index
			const uint64_t win = reinterpret_cast<uint64_t>(comm_window_peer(comm, 0, p));
			h_dst[b] = win + rec_region[0] + (uint64_t)(shift_elems * (int64_t)K);
			h_dst[256 + b] = win + rec_region[1] + (uint64_t)(shift_elems * 4);
		}
		d_dst = DevBuf<uint64_t>(c, 512);
		MEMS_CUDA(cudaMemcpyAsync(d_dst.p, h_dst, sizeof h_dst, cudaMemcpyHostToDevice, c->stream));
		void* kp[2] = {keys_loc.p, nullptr};
		uint32_t* vp[2] = {vals_loc.p, nullptr};
		radix_sort_pairs(c, K == 8, kp, vp, n_loc, top, hist_top.p, "peer_scatter_records", nullptr, d_dst.p, d_dst.p + 256);
		{
			KernelScope ks(c, "peer_barrier");
			comm_window_barrier(comm);
		}
		rk_a_p = static_cast<uint8_t*>(comm_window_local(comm, 0)) + rec_region[0];
		rv_a_p = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(comm_window_local(comm, 0)) + rec_region[1]);
	} else {
		keys_part = DevBuf<uint8_t>(c, n_loc * K);
		vals_part = DevBuf<uint32_t>(c, n_loc);
		void* kp[2] = {keys_loc.p, keys_part.p};
		uint32_t* vp[2] = {vals_loc.p, vals_part.p};
		radix_sort_pairs(c, K == 8, kp, vp, n_loc, top, hist_top.p, "shard_partition_pass");
		KernelScope ks(c, direct ? "peer_all_to_all_records" : "nccl_all_to_all_records", (double)n_loc * (K + 4));
		const void* snd[2] = {keys_part.p, vals_part.p};
		const size_t eb[2] = {K, 4};
		if (direct) {
			comm_window_all_to_all(comm, 0, 2, snd, eb, rec_region, rec_counts.data(), true);
			rk_a_p = static_cast<uint8_t*>(comm_window_local(comm, 0)) + rec_region[0];
			rv_a_p = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(comm_window_local(comm, 0)) + rec_region[1]);
		} else {
			rk_own = DevBuf<uint8_t>(c, n_recv * K);
			rv_own = DevBuf<uint32_t>(c, n_recv);
			rk_a_p = rk_own.p;
			rv_a_p = rv_own.p;
			void* rcv[2] = {rk_a_p, rv_a_p};
			comm_all_to_all_v_multi(comm, 2, snd, rcv, eb, send_counts.data(), recv_counts.data());
		}
	}
	mark("all-to-all records");
	DevBuf<SeqMeta> d_gmeta(c, n_seqs);
	MEMS_CUDA(cudaMemcpyAsync(d_gmeta.p, gmeta.data(), sizeof(SeqMeta) * n_seqs, cudaMemcpyHostToDevice, c->stream));

	// ---- 5. sort the received range; hits of this rank's seed range
	const void* u_keys = rk_a_p;
	const uint32_t* u_vals = rv_a_p;
	if (n_recv) {
		SortPlan plan = make_sort_plan(sd.key_bits);
		DevBuf<uint32_t> hist(c, (size_t)plan.n_passes * 256);
		launch_histogram(c, K == 8, rk_a_p, n_recv, plan, hist.p);
		void* kp[2] = {rk_a_p, rk_b.p};
		uint32_t* vp[2] = {rv_a_p, rv_b.p};
		const int r = radix_sort_pairs(c, K == 8, kp, vp, n_recv, plan, hist.p, "radix_pass");
		u_keys = kp[r];
		u_vals = vp[r];
	}
	MatchArgs a1;
	a1.keys = u_keys;
	a1.vals = u_vals;
	a1.n = (uint32_t)n_recv;
	a1.pos_bits = pos_bits;
	a1.pos_mask = pos_bits >= 32 ? 0xffffffffu : ((1u << pos_bits) - 1u);
	a1.n_seqs = n_seqs;
	a1.max_group = n_seqs;
	a1.mode = mode;
	a1.seq_set = 0;
	a1.hit_key_bits = pos_bits <= 26 ? 56 : 64;
	a1.test_hash_bits = c->test_hash_bits;
	a1.warp_budget = c->test_walk_budget ? c->test_walk_budget : kWarpProbeBudget;
	a1.cta_budget = c->test_walk_budget ? c->test_walk_budget : kCtaRoundBudget;
	a1.planes = nullptr;
	a1.meta = d_gmeta.p;
	HitSet hits1;
	if (n_recv >= 2) find_hits<KeyT>(c, a1, hits1);
	out.max_run = hits1.max_run;

	mark("sort + run scan");
	// ---- 6. describe the hits and group them by the top 8 bits of their diagonal hash (one counting pass: the owner
	// sorts what it receives anyway); contiguous ranges of those 256 buckets go to their owner rank.  The hash key
	// travels with the hit, so the owner does not describe it again.
	const uint32_t n1 = hits1.n;
	SortPlan hplan;
	hplan.n_passes = 1;
	hplan.shift[0] = a1.hit_key_bits - 8;
	hplan.bits[0] = 8;
	DevBuf<uint64_t> hk_a(c, n1), hk_b(c, n1), d_counts(c, (size_t)2 * W * (W + 1));
	DevBuf<uint32_t> hid_a(c, n1), hid_b(c, n1), slen(c, n1), moff(c, n1), scal(c, 2);
	const uint64_t* hkey = hk_a.p;
	const uint32_t* hid = hid_a.p;
	MEMS_CUDA(cudaMemsetAsync(d_counts.p, 0, (size_t)2 * W * sizeof(uint64_t), c->stream));
	if (n1) {
		DevBuf<uint32_t> hist(c, (size_t)hplan.n_passes * 256);
		MEMS_CUDA(cudaMemsetAsync(hist.p, 0, (size_t)hplan.n_passes * 256 * sizeof(uint32_t), c->stream));
		const uint32_t hb = (n1 + 255) / 256, describe_blocks = std::min(hb, (uint32_t)c->sm_count * 8u);
		{
			KernelScope ks(c, "hit_describe");
			hit_describe_kernel<KeyT><<<describe_blocks, 256, 0, c->stream>>>(a1, hits1.start.p, hits1.len.p, n1, hk_a.p, hid_a.p, hist.p, hplan, nullptr);
			MEMS_CUDA(cudaGetLastError());
		}
		void* kp[2] = {hk_a.p, hk_b.p};
		uint32_t* vp[2] = {hid_a.p, hid_b.p};
		const int r = radix_sort_pairs(c, true, kp, vp, n1, hplan, hist.p, "hit_partition_pass");
		hkey = r ? hk_b.p : hk_a.p;
		hid = r ? hid_b.p : hid_a.p;
		{
			KernelScope ks(c, "shard_hit_lengths", (double)n1 * 10.0);
			sorted_len_kernel<<<hb, 256, 0, c->stream>>>(hid, hits1.len.p, n1, slen.p);
			MEMS_CUDA(cudaGetLastError());
		}
		exclusive_scan_u32(c, slen.p, moff.p, n1, scal.p);  // scal[0] = members in total
	}
	// what every rank sends to every rank, as one all-gather of 2 W counts per rank and ONE host sync
	if (n1) {
		KernelScope ks(c, "shard_hit_counts");
		last_len_kernel<<<1, 32, 0, c->stream>>>(slen.p, n1, scal.p + 1);
		MEMS_CUDA(cudaGetLastError());
		hit_send_counts_kernel<<<1, 32 * ((W + 32) / 32), 0, c->stream>>>(hkey, a1.hit_key_bits - 8, moff.p, n1, scal.p + 1, W, d_counts.p);
		MEMS_CUDA(cudaGetLastError());
	}
	comm_all_gather_u64(comm, d_counts.p, d_counts.p + 2 * W, (size_t)2 * W);
	std::vector<uint64_t> h_counts((size_t)2 * W * W);
	c->fetch(h_counts.data(), d_counts.p + 2 * W, h_counts.size() * sizeof(uint64_t));
	std::vector<uint64_t> hs(W), hr(W), ms(W), mr(W);
	std::vector<uint64_t> hit_counts((size_t)W * W), mem_counts((size_t)W * W);
	for (int p = 0; p < W; ++p)
		for (int d = 0; d < W; ++d) {
			hit_counts[(size_t)p * W + d] = h_counts[(size_t)p * 2 * W + d];
			mem_counts[(size_t)p * W + d] = h_counts[(size_t)p * 2 * W + W + d];
		}
	uint64_t n2 = 0, n_mem2 = 0, max_n2 = 0, max_mem2 = 0, n_mem1 = 0;
	for (int d = 0; d < W; ++d) {
		uint64_t a = 0, b = 0;
		for (int p = 0; p < W; ++p) {
			a += hit_counts[(size_t)p * W + d];
			b += mem_counts[(size_t)p * W + d];
		}
		max_n2 = std::max(max_n2, a);
		max_mem2 = std::max(max_mem2, b);
		hs[d] = hit_counts[(size_t)R * W + d];
		ms[d] = mem_counts[(size_t)R * W + d];
		n_mem1 += ms[d];
	}
	for (int p = 0; p < W; ++p) {
		hr[p] = hit_counts[(size_t)p * W + R];
		mr[p] = mem_counts[(size_t)p * W + R];
		n2 += hr[p];
		n_mem2 += mr[p];
	}
	// evaluated identically on every rank (the largest receiver of ANY rank): all throw together
	if (max_mem2 >= (1ull << 31) || max_n2 > radix_max_items()) throw Error(MEMS_ERR_UNSUPPORTED, "too many hits or hit members on one rank");
	DevBuf<uint16_t> s_len16(c, n1);
	DevBuf<uint32_t> s_mem(c, n_mem1);
	if (n1) {
		KernelScope ks(c, "shard_pack_hits", (double)n1 * 12.0 + (double)n_mem1 * 12.0);
		const uint32_t warps = (n1 + 31) / 32;
		pack_hits_kernel<KeyT><<<(warps + 7) / 8, 256, 0, c->stream>>>(a1, hid, hits1.start.p, hits1.len.p, moff.p, n1, s_len16.p, s_mem.p);
		MEMS_CUDA(cudaGetLastError());
	}
	mark("hit describe/sort/pack");
	// window 1: [hash keys u64][length words u16][members u32], every region sized for the largest receiver
	const size_t hit_region[3] = {0, align256(max_n2 * 8), align256(max_n2 * 8) + align256(max_n2 * 2)};
	const bool direct_hits = direct && comm_window_reserve(comm, 1, hit_region[2] + align256(max_mem2 * 4));
	DevBuf<uint64_t> r_key_own;
	DevBuf<uint16_t> r_len_own;
	DevBuf<uint32_t> r_mem_own;
	uint64_t* r_key_p;
	uint16_t* r_len_p;
	uint32_t* r_mem_p;
	{
		KernelScope ks(c, direct_hits ? "peer_all_to_all_hits" : "nccl_all_to_all_hits", (double)n1 * 10 + (double)n_mem1 * 4);
		if (direct_hits) {
			uint8_t* base = static_cast<uint8_t*>(comm_window_local(comm, 1));
			r_key_p = reinterpret_cast<uint64_t*>(base + hit_region[0]);
			r_len_p = reinterpret_cast<uint16_t*>(base + hit_region[1]);
			r_mem_p = reinterpret_cast<uint32_t*>(base + hit_region[2]);
			const void* s1[2] = {hkey, s_len16.p};
			const size_t e1[2] = {8, 2};
			comm_window_all_to_all(comm, 1, 2, s1, e1, hit_region, hit_counts.data(), false);
			const void* s2[1] = {s_mem.p};
			const size_t e2[1] = {4};
			comm_window_all_to_all(comm, 1, 1, s2, e2, hit_region + 2, mem_counts.data(), true);
		} else {
			r_key_own = DevBuf<uint64_t>(c, n2);
			r_len_own = DevBuf<uint16_t>(c, n2);
			r_mem_own = DevBuf<uint32_t>(c, n_mem2);
			r_key_p = r_key_own.p;
			r_len_p = r_len_own.p;
			r_mem_p = r_mem_own.p;
			const void* snd[2] = {hkey, s_len16.p};
			void* rcv[2] = {r_key_p, r_len_p};
			const size_t eb[2] = {8, 2};
			comm_all_to_all_v_multi(comm, 2, snd, rcv, eb, hs.data(), hr.data());
			comm_all_to_all_v(comm, s_mem.p, ms.data(), r_mem_p, mr.data(), 4);
		}
	}
	mark("all-to-all hits");
	// ---- 7. this rank's diagonals: segments, walks, components
	comm_all_gather_v_wait(comm);  // the gathered planes are needed from here on
	if (direct_planes) comm_window_barrier(comm);  // ... on every rank: the peers' copies into this rank's window are done
	out.n_hits = n2;
	if (n2 == 0) {
		MEMS_CUDA(cudaStreamSynchronize(c->stream));  // the send buffers go out of scope
		return;
	}
	HitSet hits2;
	hits2.n = (uint32_t)n2;
	hits2.start = DevBuf<uint32_t>(c, n2);
	hits2.len = DevBuf<uint16_t>(c, n2);
	DevBuf<KeyT> keys2(c, n_mem2);
	DevBuf<uint32_t> len32(c, n2);
	{
		KernelScope ks(c, "shard_received_hits", (double)n2 * 8.0 + (double)n_mem2 * (8.0 + sizeof(KeyT)));
		const uint32_t nmax = (uint32_t)std::max<uint64_t>(n2, n_mem2);
		received_hits_kernel<KeyT><<<(nmax + 255) / 256, 256, 0, c->stream>>>(r_len_p, r_mem_p, (uint32_t)n2, (uint32_t)n_mem2, len32.p,
		                                                                        hits2.len.p, keys2.p);
		MEMS_CUDA(cudaGetLastError());
	}
	exclusive_scan_u32(c, len32.p, hits2.start.p, n2, nullptr);
	MatchArgs a2 = a1;
	a2.keys = keys2.p;
	a2.vals = r_mem_p;
	a2.n = (uint32_t)n_mem2;
	a2.planes = planes_all_p;
	extend_hits<KeyT>(ctx, a2, sd, hits2, order, 40000u, out, nullptr, r_key_p);
	MEMS_CUDA(cudaStreamSynchronize(c->stream));
	mark("extend + emit + D2H");
}

void find_matches_sharded(std::shared_ptr<Ctx> ctx, Comm* comm, int n_seqs, const char* const* seqs, const uint64_t* lens,
                          uint64_t seed, int mode, int order, MatchResult& out) {
	if (mode != MEMS_MODE_MEMHASH)
		throw Error(MEMS_ERR_UNSUPPORTED, "only MemHash is sharded (RepeatHash needs the reference's table order: run it on one GPU)");
	if (order == MEMS_ORDER_REFERENCE)
		throw Error(MEMS_ERR_UNSUPPORTED, "the reference's table order needs all hits in one place; use ORDER_ANY or ORDER_CANONICAL when sharded");
	if (n_seqs < 1 || n_seqs > MEMS_MAX_SEQS) throw Error(MEMS_ERR_UNSUPPORTED, "1..MEMS_MAX_SEQS sequences");
	MEMS_CUDA(cudaSetDevice(ctx->device));
	const SeedDesc sd = make_seed_desc(seed);
	if (sd.key_bits > 32)
		find_matches_sharded_typed<uint64_t>(ctx, comm, sd, n_seqs, seqs, lens, seed, mode, order, out);
	else
		find_matches_sharded_typed<uint32_t>(ctx, comm, sd, n_seqs, seqs, lens, seed, mode, order, out);
}

}  // namespace mems
