// kernels_sml.cu — 2-bit packing and spaced-seed mer extraction (SML build, stage 1).
//
// Replaces, for DNA: SortedMerList::translate32 (SortedMerList.cpp:425-460), GetMer (:321-342),
// GetSeedMer (:726-762), RevCompMer (:597-614), GetDnaSeedMer (:764-769), FillDnaSeedSML (:771-783)
// and the rolling FillDnaSML (:617-723), which yields the same keys for solid seeds.
#include "common.cuh"
#include "seed_dev.cuh"

namespace mems {

// ------------------------------------------------------------------------------------------------
// pack: one thread per output word = 16 bases, one 128-bit load of ASCII per thread.
// grid.y = sequence, grid.x covers the longest sequence's words (incl. the two zero pad words).
constexpr int kPackThreads = 256;

__global__ void __launch_bounds__(kPackThreads)
pack_kernel(const uint8_t* __restrict__ ascii, uint32_t* __restrict__ packed, const SeqMeta* __restrict__ meta,
            uint32_t* __restrict__ gap_flag) {
	// BasicDNATable (SortedMerList.cpp:29-47): c,b,y -> 1; g,s,k -> 2; t -> 3; everything else (incl. N) -> 0
	__shared__ uint8_t lut[256];
	for (int i = threadIdx.x; i < 256; i += kPackThreads) {
		int c = i | 0x20;  // fold case
		uint8_t v = 0;
		bool alpha = (i >= 'A' && i <= 'Z') || (i >= 'a' && i <= 'z');
		if (alpha) {
			if (c == 'c' || c == 'b' || c == 'y') v = 1;
			else if (c == 'g' || c == 's' || c == 'k') v = 2;
			else if (c == 't') v = 3;
		}
		if (i == '-') v = 4;  // gap marker: reported, packed as 0
		lut[i] = v;
	}
	__syncthreads();
	const SeqMeta m = meta[blockIdx.y];
	const uint64_t n_words = ((uint64_t)m.n_bases + 15) / 16 + 2;
	uint64_t wi = (uint64_t)blockIdx.x * kPackThreads + threadIdx.x;
	if (wi >= n_words) return;
	uint32_t out = 0;
	uint64_t base0 = wi * 16;
	if (base0 < m.n_bases) {
		// byte_off is relative to the staging buffer; a sequence that already lives in device memory is read where it
		// lies (its byte_off then is the distance from the staging buffer to it).  16-byte aligned either way.
		const uint8_t* src = reinterpret_cast<const uint8_t*>(reinterpret_cast<uint64_t>(ascii) + m.byte_off) + base0;
		uint32_t n_here = m.n_bases - base0 < 16 ? (uint32_t)(m.n_bases - base0) : 16u;
		uint32_t r[4] = {0, 0, 0, 0};
		if (n_here == 16u) {
			const uint4 raw = *reinterpret_cast<const uint4*>(src);
			r[0] = raw.x; r[1] = raw.y; r[2] = raw.z; r[3] = raw.w;
		} else {  // the sequence's last bases: never read past its end (it may be the end of a caller's allocation)
			for (uint32_t k = 0; k < n_here; ++k) r[k >> 2] |= (uint32_t)src[k] << ((k & 3) * 8);
		}
		bool gap = false;
#pragma unroll
		for (int k = 0; k < 16; ++k) {
			uint32_t ch = (r[k >> 2] >> ((k & 3) * 8)) & 0xffu;
			uint32_t v = k < (int)n_here ? lut[ch] : 0u;
			gap |= (v == 4u);
			out |= (v & 3u) << (30 - 2 * k);
		}
		if (gap) atomicOr(gap_flag, 1u);
	}
	packed[m.word_off + wi] = out;
}

void launch_pack(Ctx* c, const uint8_t* d_ascii, uint32_t* d_packed, const SeqMeta* d_meta, const SeqMeta* h_meta,
                 int n_seqs, uint32_t* d_gap_flag) {
	uint64_t max_words = 0, total_bases = 0;
	for (int g = 0; g < n_seqs; ++g) {
		uint64_t nw = ((uint64_t)h_meta[g].n_bases + 15) / 16 + 2;
		if (nw > max_words) max_words = nw;
		total_bases += h_meta[g].n_bases;
	}
	dim3 grid((unsigned)((max_words + kPackThreads - 1) / kPackThreads), (unsigned)n_seqs);
	KernelScope ks(c, "pack", 1.25 * (double)total_bases);
	pack_kernel<<<grid, kPackThreads, 0, c->stream>>>(d_ascii, d_packed, d_meta, d_gap_flag);
	MEMS_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// bit planes: the 2-bit codes of 32 consecutive bases as one word of high bits and one of low bits, base j of the
// group in bit j.  Same bytes as the packed words, re-laid for the window test of match extension: XOR-ing two
// members' planes gives one disagreement bit per base without any 2-bit fold (kernels_match.cu, WarpHit::probe).
__device__ __forceinline__ uint32_t plane16(uint32_t w) {  // bits 0,2,..,30 of w (base 15 first) -> bit b = base b
	w &= 0x55555555u;
	w = (w | (w >> 1)) & 0x33333333u;
	w = (w | (w >> 2)) & 0x0f0f0f0fu;
	w = (w | (w >> 4)) & 0x00ff00ffu;
	w = (w | (w >> 8)) & 0x0000ffffu;
	return __brev(w) >> 16;
}

__global__ void __launch_bounds__(256)
planes_kernel(const uint2* __restrict__ packed2, uint2* __restrict__ planes, uint64_t n) {
	const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	if (i >= n) return;
	const uint2 w = packed2[i];  // bases 32i..32i+15 in w.x, 32i+16..32i+31 in w.y (MSB-first pairs)
	planes[i] = make_uint2(plane16(w.x >> 1) | (plane16(w.y >> 1) << 16), plane16(w.x) | (plane16(w.y) << 16));
}

void launch_planes(Ctx* c, const uint32_t* d_packed, uint2* d_planes, uint64_t n_words) {
	const uint64_t n = n_words / 2;
	if (n == 0) return;
	KernelScope ks(c, "planes", (double)n_words * 8.0);
	planes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<const uint2*>(d_packed), d_planes, n);
	MEMS_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// extract: one thread per 8 consecutive seed positions of one sequence, chosen so that the thread's 8 slots of
// the union arrays start on a 16-byte boundary: keys and values leave as 128-bit stores.  The thread reads the
// four packed words under its positions once (coalesced, L1 serves the overlap between neighbours), shifts them
// into place once, and derives each window with two funnel shifts by a constant.  Loops run over the pattern's
// runs / the radix passes on the outside and over the 8 items on the inside, so each run descriptor is fetched
// once per thread.  For weights <= 16 the mer fits 32 bits and every step works on single registers.
// The digit histograms of ALL radix passes are accumulated here (shared-memory atomics, one flush per CTA),
// which removes the separate histogram pre-pass of a classic onesweep sort.
// Global traffic: 0.25 B/base read + one key and one value written per seed.
constexpr int kExtractThreads = 256;
constexpr int kExtractItems = 8;
constexpr int kExtractTile = kExtractThreads * kExtractItems;
constexpr int kMaxPasses = 8;

struct PassDesc {
	int n_passes;
	int shift[kMaxPasses];
	int bits[kMaxPasses];
};

template <class KeyT>
__device__ __forceinline__ void store8(KeyT* dst, const KeyT (&v)[8]);
template <>
__device__ __forceinline__ void store8<uint32_t>(uint32_t* dst, const uint32_t (&v)[8]) {
	reinterpret_cast<uint4*>(dst)[0] = make_uint4(v[0], v[1], v[2], v[3]);
	reinterpret_cast<uint4*>(dst)[1] = make_uint4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<uint64_t>(uint64_t* dst, const uint64_t (&v)[8]) {
#pragma unroll
	for (int j = 0; j < 4; ++j) reinterpret_cast<ulonglong2*>(dst)[j] = make_ulonglong2(v[2 * j], v[2 * j + 1]);
}

template <class KeyT, bool W32>
__global__ void __launch_bounds__(kExtractThreads)
extract_kernel(const uint32_t* __restrict__ packed, const SeqMeta* __restrict__ meta, SeedDesc sd, int pos_bits,
               KeyT* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ hist, PassDesc pd,
               uint32_t tiles_per_cta) {
	__shared__ uint32_t s_hist[kMaxPasses * 256];
	const SeqMeta m = meta[blockIdx.y];
	const uint32_t lead = (uint32_t)(m.seed_off & 3u);  // slots before the sequence's first seed in its 4-slot group
	const uint64_t span = (uint64_t)m.n_seeds + lead;
	if ((uint64_t)blockIdx.x * tiles_per_cta * kExtractTile >= span) return;
	for (int i = threadIdx.x; i < pd.n_passes * 256; i += kExtractThreads) s_hist[i] = 0;
	__syncthreads();
	const uint32_t seq_tag = m.tag << pos_bits;
	const KeyT group_key = m.group ? (KeyT)((KeyT)m.group << sd.key_bits) : (KeyT)0;  // problem number above the key bits
	const uint32_t* src = packed + m.word_off;
	// a CTA walks tiles_per_cta consecutive tiles and flushes its histograms once: one flush per tile would be
	// ~1000 same-address global atomics per 2048 seeds, which bounds the kernel at the L2 atomic units
	for (uint32_t t = 0; t < tiles_per_cta; ++t) {
	const uint64_t tile0 = ((uint64_t)blockIdx.x * tiles_per_cta + t) * kExtractTile;
	if (tile0 >= span) break;
	// first position of this thread; negative (as int64) only for the first thread of a sequence with lead > 0
	const int64_t p0 = (int64_t)tile0 + (int64_t)threadIdx.x * kExtractItems - (int64_t)lead;
	KeyT key[kExtractItems];
	bool all = p0 >= 0 && p0 + kExtractItems <= (int64_t)m.n_seeds;
	if (all) {
		// words wi .. wi+3 hold bits [0, 128) of the thread's span; the last valid window ends before bit
		// 2*15 + 2*7 + 62.  The buffer carries two zero pad words; reads past it are clamped.
		const uint64_t n_words = ((uint64_t)m.n_bases + 15) / 16 + 2;
		const uint32_t wi = (uint32_t)p0 >> 4, sh = ((uint32_t)p0 & 15u) * 2u;
		const uint32_t w0 = src[wi], w1 = src[wi + 1], w2 = src[wi + 2];
		const uint32_t w3 = (uint64_t)wi + 3 < n_words ? src[wi + 3] : 0u;
		const uint32_t v0 = __funnelshift_l(w1, w0, sh), v1 = __funnelshift_l(w2, w1, sh), v2 = __funnelshift_l(w3, w2, sh);
		uint32_t hi[kExtractItems], lo[kExtractItems];
#pragma unroll
		for (int k = 0; k < kExtractItems; ++k) {
			hi[k] = __funnelshift_l(v1, v0, 2 * k);
			lo[k] = __funnelshift_l(v2, v1, 2 * k);
		}
		if (W32) {
			uint32_t f[kExtractItems];
#pragma unroll
			for (int k = 0; k < kExtractItems; ++k) f[k] = 0;
			// fully unrolled with a uniform early exit: run descriptors are then read from the kernel-parameter
			// constant bank at fixed offsets (a runtime-indexed parameter array would be copied to local memory)
#pragma unroll
			for (int r = 0; r < kMaxSeedRuns; ++r) {
				if (r >= sd.n_runs) break;
				const uint32_t net = sd.run_net[r], mask = (uint32_t)sd.run_mask[r];
				if (net >= 32u) {
#pragma unroll
					for (int k = 0; k < kExtractItems; ++k) f[k] |= (hi[k] >> (net - 32u)) & mask;
				} else {
#pragma unroll
					for (int k = 0; k < kExtractItems; ++k) f[k] |= __funnelshift_r(lo[k], hi[k], net) & mask;
				}
			}
#pragma unroll
			for (int k = 0; k < kExtractItems; ++k) key[k] = (KeyT)canonical_key((uint64_t)f[k], sd.w) | group_key;
		} else {
			uint64_t f[kExtractItems];
#pragma unroll
			for (int k = 0; k < kExtractItems; ++k) f[k] = 0;
#pragma unroll
			for (int r = 0; r < kMaxSeedRuns; ++r) {
				if (r >= sd.n_runs) break;
				const uint32_t net = sd.run_net[r];
				const uint64_t mask = sd.run_mask[r];
#pragma unroll
				for (int k = 0; k < kExtractItems; ++k) f[k] |= ((((uint64_t)hi[k] << 32) | lo[k]) >> net) & mask;
			}
#pragma unroll
			for (int k = 0; k < kExtractItems; ++k) key[k] = (KeyT)canonical_key(f[k], sd.w) | group_key;
		}
		const uint64_t slot = m.seed_off + (uint64_t)p0;  // a multiple of 4
		store8<KeyT>(keys + slot, key);
		uint32_t val[kExtractItems];
#pragma unroll
		for (int k = 0; k < kExtractItems; ++k) val[k] = seq_tag | ((uint32_t)p0 + k);
		store8<uint32_t>(vals + slot, val);
#pragma unroll
		for (int q = 0; q < kMaxPasses; ++q) {
			if (q >= pd.n_passes) break;
			const int shift = pd.shift[q];
			const uint32_t mask = (1u << pd.bits[q]) - 1u;
#pragma unroll
			for (int k = 0; k < kExtractItems; ++k) atomicAdd(&s_hist[q * 256 + ((uint32_t)(key[k] >> shift) & mask)], 1u);
		}
	} else {
		// the (at most two) threads per sequence whose span crosses the sequence's first or last seed
		for (int k = 0; k < kExtractItems; ++k) {
			const int64_t p = p0 + k;
			if (p < 0 || p >= (int64_t)m.n_seeds) continue;
			const uint64_t ck = canonical_key(extract_fwd(window64(src, (uint32_t)p), sd), sd.w) | (uint64_t)group_key;
			keys[m.seed_off + p] = (KeyT)ck;
			vals[m.seed_off + p] = seq_tag | (uint32_t)p;
#pragma unroll
			for (int q = 0; q < kMaxPasses; ++q) {
				if (q >= pd.n_passes) break;
				atomicAdd(&s_hist[q * 256 + ((uint32_t)(ck >> pd.shift[q]) & ((1u << pd.bits[q]) - 1u))], 1u);
			}
		}
	}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < pd.n_passes * 256; i += kExtractThreads) {
		const uint32_t v = s_hist[i];
		if (v) atomicAdd(&hist[i], v);
	}
}

void launch_extract(Ctx* c, const uint32_t* d_packed, const SeqMeta* d_meta, const SeqMeta* h_meta, int n_seqs,
                    const SeedDesc& sd, int pos_bits, bool key64, void* d_keys, uint32_t* d_vals, uint32_t* d_hist,
                    int n_passes, const int* pass_shift, const int* pass_bits) {
	uint32_t max_seeds = 0;
	uint64_t total = 0;
	for (int g = 0; g < n_seqs; ++g) {
		if (h_meta[g].n_seeds > max_seeds) max_seeds = h_meta[g].n_seeds;
		total += h_meta[g].n_seeds;
	}
	if (max_seeds == 0) return;
	if (n_passes > kMaxPasses) throw Error(4, "too many radix passes");
	PassDesc pd;
	pd.n_passes = n_passes;
	for (int q = 0; q < n_passes; ++q) {
		pd.shift[q] = pass_shift[q];
		pd.bits[q] = pass_bits[q];
	}
	const uint64_t tiles_max = ((uint64_t)max_seeds + 3 + kExtractTile - 1) / kExtractTile;
	const uint64_t tiles_all = (total + 3 * (uint64_t)n_seqs + kExtractTile - 1) / kExtractTile;
	uint64_t tpc = (tiles_all + (uint64_t)c->sm_count * 8 - 1) / ((uint64_t)c->sm_count * 8);  // ~8 CTAs per SM in total
	if (tpc < 1) tpc = 1;
	dim3 grid((unsigned)((tiles_max + tpc - 1) / tpc), (unsigned)n_seqs);
	double bytes = (double)total * (0.25 + (key64 ? 8.0 : 4.0) + 4.0);
	KernelScope ks(c, "extract", bytes);
	const bool w32 = sd.w <= 16;
	if (!key64)
		extract_kernel<uint32_t, true><<<grid, kExtractThreads, 0, c->stream>>>(d_packed, d_meta, sd, pos_bits,
		                                                                         (uint32_t*)d_keys, d_vals, d_hist, pd, (uint32_t)tpc);
	else if (w32)
		extract_kernel<uint64_t, true><<<grid, kExtractThreads, 0, c->stream>>>(d_packed, d_meta, sd, pos_bits,
		                                                                         (uint64_t*)d_keys, d_vals, d_hist, pd, (uint32_t)tpc);
	else
		extract_kernel<uint64_t, false><<<grid, kExtractThreads, 0, c->stream>>>(d_packed, d_meta, sd, pos_bits,
		                                                                          (uint64_t*)d_keys, d_vals, d_hist, pd, (uint32_t)tpc);
	MEMS_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// accessors (SortedMerList::GetSeedMer / GetDnaSeedMer / MemorySML::Read): gather-style, one thread per query.
__global__ void seed_mers_kernel(const uint32_t* __restrict__ words, uint32_t n_seeds, SeedDesc sd,
                                 const uint64_t* __restrict__ pos, uint64_t n, uint64_t* __restrict__ fwd_out,
                                 uint64_t* __restrict__ dna_out) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint64_t p = pos[i];
	uint64_t f = 0, d = 0;
	if (p < n_seeds) {
		uint64_t fw = extract_fwd(window64(words, (uint32_t)p), sd);
		f = fw << (64 - 2 * sd.w);
		d = to_reference_mer(canonical_key(fw, sd.w), sd.w);
	}
	if (fwd_out) fwd_out[i] = f;
	if (dna_out) dna_out[i] = d;
}

void launch_seed_mers(Ctx* c, const uint32_t* d_words, uint32_t n_seeds, const SeedDesc& sd, const uint64_t* d_pos,
                      uint64_t n, uint64_t* d_fwd, uint64_t* d_dna) {
	if (n == 0) return;
	KernelScope ks(c, "seed_mers");
	seed_mers_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_words, n_seeds, sd, d_pos, n, d_fwd, d_dna);
	MEMS_CUDA(cudaGetLastError());
}

__global__ void sml_read_kernel(const uint32_t* __restrict__ words, SeedDesc sd, const uint32_t* __restrict__ positions,
                                uint32_t pos_mask, uint64_t count, uint64_t* __restrict__ mers) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	uint64_t fw = extract_fwd(window64(words, positions[i] & pos_mask), sd);
	mers[i] = to_reference_mer(canonical_key(fw, sd.w), sd.w);
}

// SortedMerList::FindMer/bsearch (SortedMerList.cpp:170-179,380-394): the same bisection (middle = (start+end)/2,
// inclusive bounds), so "not found" reports the same index the reference would.  One thread; a few dozen probes.
__global__ void find_mer_kernel(const uint32_t* __restrict__ words, SeedDesc sd, const uint32_t* __restrict__ positions,
                                uint32_t pos_mask, uint64_t n, uint64_t query, uint64_t* __restrict__ result) {
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	uint64_t start = 0, end = n - 1, middle = 0;
	bool found = false;
	while (true) {
		middle = (start + end) / 2;
		uint64_t fw = extract_fwd(window64(words, positions[middle] & pos_mask), sd);
		uint64_t mer = to_reference_mer(canonical_key(fw, sd.w), sd.w);
		if (mer == query) {
			found = true;
			break;
		} else if (mer < query && middle < end)
			start = middle + 1;
		else if (mer > query && start < middle)
			end = middle - 1;
		else
			break;
	}
	result[0] = middle;
	result[1] = found ? 1 : 0;
}

void launch_find_mer(Ctx* c, const uint32_t* d_words, const SeedDesc& sd, const uint32_t* d_positions, uint32_t pos_mask,
                     uint64_t n, uint64_t query_mer, uint64_t* d_result) {
	KernelScope ks(c, "find_mer");
	find_mer_kernel<<<1, 32, 0, c->stream>>>(d_words, sd, d_positions, pos_mask, n, query_mer, d_result);
	MEMS_CUDA(cudaGetLastError());
}

void launch_sml_read(Ctx* c, const uint32_t* d_words, const SeedDesc& sd, const uint32_t* d_positions, uint32_t pos_mask,
                     uint64_t count, uint64_t* d_mers) {
	if (count == 0) return;
	KernelScope ks(c, "sml_read");
	sml_read_kernel<<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(d_words, sd, d_positions, pos_mask, count, d_mers);
	MEMS_CUDA(cudaGetLastError());
}

}  // namespace mems

// ------------------------------------------------------------------------------------------------ seed occurrence
// SeedOccurrenceList::construct (SeedOccurrenceList.h:21-92): for every base position the number of times the
// seed starting there occurs in the sequence (the length of its equal-masked-key run in the sorted list),
// positions past the last seed count 1, then each value becomes the mean over the L seeds that contain the
// position (window sum / L in double, stored as float; the first L-1 positions see "1" for seeds before the start).
namespace mems {

template <class KeyT>
__global__ void occ_heads_kernel(const uint32_t* __restrict__ positions, uint32_t pos_mask, const KeyT* __restrict__ key_pos,
                                 uint32_t n, uint32_t* __restrict__ is_head) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const KeyT k = key_pos[positions[i] & pos_mask] >> 1;
	is_head[i] = (i == 0 || (KeyT)(key_pos[positions[i - 1] & pos_mask] >> 1) != k) ? 1u : 0u;
}

__global__ void occ_run_starts_kernel(const uint32_t* __restrict__ is_head, const uint32_t* __restrict__ run_of, uint32_t n,
                                      uint32_t* __restrict__ run_start) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && is_head[i]) run_start[run_of[i]] = i;
}

__global__ void occ_counts_kernel(const uint32_t* __restrict__ positions, uint32_t pos_mask, const uint32_t* __restrict__ is_head,
                                  const uint32_t* __restrict__ run_of, const uint32_t* __restrict__ run_start, uint32_t n_runs,
                                  uint32_t n, uint32_t* __restrict__ count) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t r = is_head[i] ? run_of[i] : run_of[i] - 1;  // run_of is an exclusive scan of is_head
	const uint32_t end = r + 1 < n_runs ? run_start[r + 1] : n;
	count[positions[i] & pos_mask] = end - run_start[r];
}

__global__ void occ_smooth_kernel(const uint32_t* __restrict__ count, uint32_t n_bases, uint32_t n_seeds, int L,
                                  float* __restrict__ out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_bases) return;
	// raw value of position j: its run length, 1 past the last seed (and for "positions" before the start)
	auto raw = [&](int64_t j) -> uint64_t {
		if (j < 0) return 1;
		const uint32_t tail_from = n_seeds > 1 ? n_seeds : 1;
		return (uint32_t)j >= tail_from ? 1u : (uint64_t)count[j];
	};
	float v;
	if (i + 1 < n_bases) {
		uint64_t sum = 0;
		for (int64_t j = (int64_t)i - L + 1; j <= (int64_t)i; ++j) sum += raw(j);
		v = (float)((double)sum / (double)L);
	} else {
		v = (float)raw(i);  // the reference's loop never smooths the last position
	}
	out[i] = v == 0.f ? 1.f : v;
}

void launch_seed_occurrence(Ctx* c, const uint32_t* d_positions, uint32_t pos_mask, const void* d_key_pos, bool key64,
                            uint32_t n_seeds, uint32_t n_bases, int L, float* d_out) {
	if (n_bases == 0) return;
	DevBuf<uint32_t> count(c, n_bases);
	MEMS_CUDA(cudaMemsetAsync(count.p, 0, (size_t)n_bases * sizeof(uint32_t), c->stream));
	if (n_seeds) {
		DevBuf<uint32_t> is_head(c, n_seeds), run_of(c, n_seeds), total(c, 1);
		const unsigned blocks = (n_seeds + 255) / 256;
		KernelScope ks(c, "seed_occurrence");
		if (key64)
			occ_heads_kernel<uint64_t><<<blocks, 256, 0, c->stream>>>(d_positions, pos_mask, (const uint64_t*)d_key_pos, n_seeds, is_head.p);
		else
			occ_heads_kernel<uint32_t><<<blocks, 256, 0, c->stream>>>(d_positions, pos_mask, (const uint32_t*)d_key_pos, n_seeds, is_head.p);
		MEMS_CUDA(cudaGetLastError());
		exclusive_scan_u32(c, is_head.p, run_of.p, n_seeds, total.p);
		uint32_t n_runs = 0;
		MEMS_CUDA(cudaMemcpyAsync(&n_runs, total.p, 4, cudaMemcpyDeviceToHost, c->stream));
		MEMS_CUDA(cudaStreamSynchronize(c->stream));
		DevBuf<uint32_t> run_start(c, n_runs);
		occ_run_starts_kernel<<<blocks, 256, 0, c->stream>>>(is_head.p, run_of.p, n_seeds, run_start.p);
		MEMS_CUDA(cudaGetLastError());
		occ_counts_kernel<<<blocks, 256, 0, c->stream>>>(d_positions, pos_mask, is_head.p, run_of.p, run_start.p, n_runs, n_seeds,
		                                                 count.p);
		MEMS_CUDA(cudaGetLastError());
		occ_smooth_kernel<<<(n_bases + 255) / 256, 256, 0, c->stream>>>(count.p, n_bases, n_seeds, L, d_out);
		MEMS_CUDA(cudaGetLastError());
		MEMS_CUDA(cudaStreamSynchronize(c->stream));  // scratch buffers die with this scope
	} else {
		KernelScope ks(c, "seed_occurrence");
		occ_smooth_kernel<<<(n_bases + 255) / 256, 256, 0, c->stream>>>(count.p, n_bases, 0, L, d_out);
		MEMS_CUDA(cudaGetLastError());
		MEMS_CUDA(cudaStreamSynchronize(c->stream));
	}
}

}  // namespace mems
