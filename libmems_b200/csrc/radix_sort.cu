// radix_sort.cu — stable LSD radix sort of (key, u32 value) pairs, one "onesweep" kernel per digit.
//
// Replaces std::sort over vector<bmer> in MemorySML::Create (MemorySML.cpp:54, comparator
// SortedMerList.h:312-314).  The reference sorts 16-byte bmer records by the full 64-bit mer; here the
// key is the compact canonical key (2w+1 bits, see seed_dev.cuh) whose order is identical, so only
// ceil((2w+1)/8) digit passes run.  Stability makes ties (which std::sort leaves unspecified) come out
// in ascending position order.
//
// Per pass, one CTA per tile of kTile pairs:
//   1. coalesced warp-striped load of the keys (each warp owns a contiguous 32*kItems chunk),
//   2. per-warp digit ranking from ballots (no atomics, order preserving) against per-warp counters in
//      shared memory,
//   3. the tile's digit counts are published, the keys are reordered through shared memory (values go
//      straight from global memory to their reordered slot with cp.async),
//   4. per-digit decoupled look-back across tiles (thread d resolves digit d, a few predecessors per round)
//      against the digit histogram that extraction already produced — a tile never waits for more than its
//      predecessors' 256 counters, so the whole pass is a single read and a single write of the data,
//   5. coalesced store: every digit's run leaves the CTA as one contiguous segment — into the second
//      buffers, or (PEER) into per-digit destinations that may be other GPUs' exchange windows.
// Tiles take their index from an atomic ticket, which guarantees that all predecessors of a running
// tile are themselves running or finished (forward progress of the look-back).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace mems {

constexpr int kRadix = 256;
#ifndef MEMS_RADIX_TILE
#define MEMS_RADIX_TILE 4096
#endif
#ifndef MEMS_RADIX_MINB32
#define MEMS_RADIX_MINB32 5
#endif
#ifndef MEMS_RADIX_MINB64
#define MEMS_RADIX_MINB64 4
#endif
constexpr int kTile = MEMS_RADIX_TILE;  // pairs per CTA
constexpr int kMinB32 = MEMS_RADIX_MINB32, kMinB64 = MEMS_RADIX_MINB64;  // resident CTAs per SM the kernels are compiled for

constexpr uint32_t kFlagPartial = 0x40000000u;
constexpr uint32_t kFlagInclusive = 0x80000000u;
constexpr uint32_t kFlagMask = 0xC0000000u;
constexpr uint32_t kValueMask = 0x3FFFFFFFu;

size_t radix_max_items() { return (size_t)kValueMask; }

SortPlan make_sort_plan(int key_bits, int begin_bit) {
	SortPlan p;
	if (key_bits < 1) key_bits = 1;
	p.n_passes = (key_bits + 7) / 8;
	// spread the bits evenly so no pass is narrower than it has to be
	int base = key_bits / p.n_passes, extra = key_bits % p.n_passes, at = begin_bit;
	for (int q = 0; q < p.n_passes; ++q) {
		p.bits[q] = base + (q < extra ? 1 : 0);
		p.shift[q] = at;
		at += p.bits[q];
	}
	return p;
}

// tile status words carry flag and value together, so relaxed device-scope accesses are enough
__device__ __forceinline__ uint32_t ld_status(const uint32_t* p) {
	uint32_t v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_status(uint32_t* p, uint32_t v) {
	asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Lanes of the warp that hold the same digit as this lane, from NBITS ballots (MATCH.ANY is one instruction but
// occupies its pipe for tens of cycles; a ballot costs test + VOTE + select + AND at full issue rate).
// written in PTX: the C++ form makes the compiler derive the vote predicate and the select mask separately
// (7 instructions per bit instead of 4)
template <int B>
__device__ __forceinline__ void peers_bit(uint32_t& peers, uint32_t d) {
	asm volatile(
	    "{\n\t"
	    ".reg .pred p;\n\t"
	    ".reg .b32 t, m;\n\t"
	    "and.b32 t, %1, %2;\n\t"
	    "setp.ne.u32 p, t, 0;\n\t"
	    "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
	    "@!p not.b32 m, m;\n\t"
	    "and.b32 %0, %0, m;\n\t"
	    "}"
	    : "+r"(peers)
	    : "r"(d), "n"(1 << B));
}
template <int NBITS>
__device__ __forceinline__ uint32_t digit_peers(uint32_t d) {
	uint32_t peers = 0xffffffffu;
	peers_bit<0>(peers, d);
	peers_bit<1>(peers, d);
	peers_bit<2>(peers, d);
	peers_bit<3>(peers, d);
	peers_bit<4>(peers, d);
	peers_bit<5>(peers, d);
	peers_bit<6>(peers, d);
	if (NBITS > 7) peers_bit<7>(peers, d);
	return peers;
}

__device__ __forceinline__ void cp_async_u32(void* smem_dst, const uint32_t* gmem_src) {
	const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// One tile of one pass.  THREADS x ITEMS = kTile, THREADS == kRadix (thread d also owns digit d).  The pass is
// bound by instruction issue, not by HBM (ncu, profiles/r01_*), so the body is written for instruction count:
// FULL tiles carry no bounds predicates; every lane reads its digit's running counter itself (one LDS) instead
// of a leader read + shuffle; the scan over digits folds the tile-local digit offset into the per-warp counters
// so the reorder step is one LDS + add per item; values never pass through registers (cp.async straight into
// their reordered shared-memory slot); the tile's partial counts are published before the reorder and the
// look-back runs after it, so predecessors have usually finished by the time they are polled.
template <class KeyT, int THREADS, bool FULL, int NBITS, int kLookBack, bool PEER>
__device__ __forceinline__ void onesweep_tile(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                              KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n_valid,
                                              uint32_t tile, int shift, uint32_t digit_mask,
                                              const uint32_t* __restrict__ bin_base, uint32_t* __restrict__ status, KeyT* s_keys,
                                              uint32_t* s_vals, uint16_t (*s_warp_cnt)[kRadix], uint32_t* s_global_base,
                                              uint32_t* s_warp_tot, const uint64_t* __restrict__ key_dst,
                                              const uint64_t* __restrict__ val_dst, uint64_t* s_kaddr, uint64_t* s_vaddr) {
	constexpr int WARPS = THREADS / 32;
	constexpr int ITEMS = kTile / THREADS;
	static_assert(THREADS >= kRadix, "thread d < 256 owns digit d");
	const bool owner = THREADS == kRadix || threadIdx.x < kRadix;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t tile_base = tile * (uint32_t)kTile;
	const uint32_t warp_base = warp * (32 * ITEMS);
	const KeyT* kin = keys_in + tile_base + warp_base + lane;
	const uint32_t* vin = vals_in + tile_base + warp_base + lane;

	KeyT key[ITEMS];
	uint16_t rank[ITEMS];
#pragma unroll
	for (int i = 0; i < ITEMS; ++i) {
		if (FULL) key[i] = kin[i * 32];
		else key[i] = warp_base + i * 32 + lane < n_valid ? kin[i * 32] : ~(KeyT)0;
	}
	// ---- rank inside the warp, in memory order
	const uint32_t lanemask_lt = (1u << lane) - 1u;
	uint16_t* my_cnt = s_warp_cnt[warp];
#pragma unroll
	for (int i = 0; i < ITEMS; ++i) {
		uint32_t d = (uint32_t)(key[i] >> shift) & digit_mask;
		if (!FULL && !(warp_base + i * 32 + lane < n_valid)) d = kRadix - 1;  // padding ranks after every real item
		const uint32_t peers = digit_peers<FULL ? NBITS : 8>(d);
		const uint32_t base = my_cnt[d];
		__syncwarp();
		const uint32_t below = peers & lanemask_lt;
		if (below == 0u) my_cnt[d] = (uint16_t)(base + __popc(peers));
		rank[i] = (uint16_t)(base + __popc(below));
		__syncwarp();
	}
	__syncthreads();
	// ---- thread d owns digit d: counts per warp -> exclusive over warps; tile count; scan over digits
	uint32_t count = 0, incl = 0;
	uint32_t* my_status = status + (size_t)tile * kRadix + tid;
	if (owner) {
#pragma unroll
		for (int w = 0; w < WARPS; ++w) count += s_warp_cnt[w][tid];
		if (!FULL && tid == kRadix - 1) count -= (uint32_t)kTile - n_valid;
		st_status(my_status, count | (tile == 0 ? kFlagInclusive : kFlagPartial));
		incl = count;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += t;
		}
		if (lane == 31) s_warp_tot[warp] = incl;
	}
	__syncthreads();
	uint32_t digit_excl = incl - count;  // first tile-local sorted index of digit tid
	if (owner) {
#pragma unroll
		for (int w = 0; w < kRadix / 32; ++w)
			if (w < warp) digit_excl += s_warp_tot[w];
		uint32_t run = digit_excl;
#pragma unroll
		for (int w = 0; w < WARPS; ++w) {
			const uint32_t c = s_warp_cnt[w][tid];
			s_warp_cnt[w][tid] = (uint16_t)run;  // tile-local sorted index of warp w's first item with digit tid
			run += c;
		}
	}
	__syncthreads();
	// ---- reorder through shared memory; values go straight from global to their slot
#pragma unroll
	for (int i = 0; i < ITEMS; ++i) {
		const bool valid = FULL || warp_base + i * 32 + lane < n_valid;
		const uint32_t d = valid ? ((uint32_t)(key[i] >> shift) & digit_mask) : (uint32_t)(kRadix - 1);
		const uint32_t pos = (uint32_t)my_cnt[d] + rank[i];
		s_keys[pos] = key[i];
		if (valid) cp_async_u32(s_vals + pos, vin + i * 32);
	}
	// ---- decoupled look-back for digit tid
	if (owner) {
		uint32_t excl = 0;
		if (tile != 0) {
			// predecessors are polled kLookBack at a time (independent loads in flight): a serial walk costs one
			// L2 round trip per predecessor and was a third of all stall samples (profiles/r01_ncu_onesweep_*)
			const uint32_t* look = my_status - kRadix;
			uint32_t left = tile;  // predecessors not yet consumed
			bool done = false;
			while (!done) {
				uint32_t s[kLookBack];
#pragma unroll
				for (int j = 0; j < kLookBack; ++j)
					s[j] = (uint32_t)j < left ? ld_status(look - j * kRadix) : kFlagInclusive;
				int used = 0;
#pragma unroll
				for (int j = 0; j < kLookBack; ++j) {
					if (!done && used == j && (s[j] & kFlagMask) != 0u) {
						excl += s[j] & kValueMask;
						used = j + 1;
						if (s[j] & kFlagInclusive) done = true;
					}
				}
				look -= used * kRadix;
				left -= used;
			}
			st_status(my_status, (excl + count) | kFlagInclusive);
		}
		const uint32_t gbase = bin_base[tid] + excl - digit_excl;  // + tile-local sorted index = global index
		s_global_base[tid] = gbase;
		if (PEER) {  // every digit has its own destination array (a peer's exchange window): byte address of index gbase
			s_kaddr[tid] = key_dst[tid] + (uint64_t)gbase * sizeof(KeyT);
			s_vaddr[tid] = val_dst[tid] + (uint64_t)gbase * sizeof(uint32_t);
		}
	}
	cp_async_wait_all();
	__syncthreads();
#pragma unroll
	for (int k = 0; k < ITEMS; ++k) {
		const uint32_t idx = k * THREADS + tid;
		if (FULL || idx < n_valid) {
			const KeyT kk = s_keys[idx];
			const uint32_t d = (uint32_t)(kk >> shift) & digit_mask;
			if (PEER) {
				*reinterpret_cast<KeyT*>(s_kaddr[d] + (uint64_t)idx * sizeof(KeyT)) = kk;
				*reinterpret_cast<uint32_t*>(s_vaddr[d] + (uint64_t)idx * sizeof(uint32_t)) = s_vals[idx];
			} else {
				const uint32_t g = s_global_base[d] + idx;
				keys_out[g] = kk;
				vals_out[g] = s_vals[idx];
			}
		}
	}
}

// PEER: the pass is also the send side of an all-to-all — digit d's run is written to key_dst[d] / val_dst[d]
// (byte addresses such that index g of this rank's partitioned order lands at address + g * element size), which
// point into the exchange windows of the ranks that own the digits: the scatter goes straight over NVLink.
template <class KeyT, int THREADS, int MINB, int NBITS, int LB = 4, bool PEER = false>
__global__ void __launch_bounds__(THREADS, MINB)
onesweep_kernel(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, KeyT* __restrict__ keys_out,
                uint32_t* __restrict__ vals_out, uint32_t n, int shift, uint32_t digit_mask,
                const uint32_t* __restrict__ bin_base, uint32_t* __restrict__ status, uint32_t* __restrict__ ticket,
                const uint64_t* __restrict__ key_dst, const uint64_t* __restrict__ val_dst) {
	constexpr int WARPS = THREADS / 32;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	KeyT* s_keys = reinterpret_cast<KeyT*>(smem_raw);
	uint32_t* s_vals = reinterpret_cast<uint32_t*>(smem_raw + sizeof(KeyT) * kTile);
	__shared__ uint16_t s_warp_cnt[WARPS][kRadix];  // per-warp digit counts, then tile-local offsets
	__shared__ uint32_t s_global_base[kRadix];
	__shared__ uint32_t s_warp_tot[kRadix / 32];
	__shared__ uint32_t s_tile;
	__shared__ uint64_t s_kaddr[PEER ? kRadix : 1], s_vaddr[PEER ? kRadix : 1];

	const int tid = threadIdx.x;
	if (tid == 0) s_tile = atomicAdd(ticket, 1u);
	{
		uint32_t* z = reinterpret_cast<uint32_t*>(&s_warp_cnt[0][0]);
		for (int i = tid; i < WARPS * kRadix / 2; i += THREADS) z[i] = 0;
	}
	__syncthreads();
	const uint32_t tile = s_tile;
	const uint32_t tile_base = tile * (uint32_t)kTile;
	const uint32_t n_valid = n - tile_base < (uint32_t)kTile ? n - tile_base : (uint32_t)kTile;
	if (n_valid == (uint32_t)kTile)
		onesweep_tile<KeyT, THREADS, true, NBITS, LB, PEER>(keys_in, vals_in, keys_out, vals_out, n_valid, tile, shift, digit_mask,
		                                                bin_base, status, s_keys, s_vals, s_warp_cnt, s_global_base, s_warp_tot,
		                                                key_dst, val_dst, s_kaddr, s_vaddr);
	else
		onesweep_tile<KeyT, THREADS, false, NBITS, LB, PEER>(keys_in, vals_in, keys_out, vals_out, n_valid, tile, shift, digit_mask,
		                                                 bin_base, status, s_keys, s_vals, s_warp_cnt, s_global_base, s_warp_tot,
		                                                 key_dst, val_dst, s_kaddr, s_vaddr);
}

// exclusive scan of each pass's 256 digit counts -> first output index of each digit
__global__ void scan_bins_kernel(const uint32_t* __restrict__ hist, uint32_t* __restrict__ bin_base) {
	__shared__ uint32_t s_tot[8];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	uint32_t c = hist[blockIdx.x * kRadix + tid];
	uint32_t incl = c;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += t;
	}
	if (lane == 31) s_tot[warp] = incl;
	__syncthreads();
	uint32_t woff = 0;
	for (int w = 0; w < warp; ++w) woff += s_tot[w];
	bin_base[blockIdx.x * kRadix + tid] = woff + incl - c;
}

template <class K>
static const K* first_arg_of(void (*)(const K*, const uint32_t*, K*, uint32_t*, uint32_t, int, uint32_t, const uint32_t*, uint32_t*,
                                      uint32_t*, const uint64_t*, const uint64_t*)) {
	return nullptr;
}

int radix_sort_pairs(Ctx* c, bool key64, void* d_keys[2], uint32_t* d_vals[2], uint64_t n, const SortPlan& plan,
                     uint32_t* d_hist, const char* prof_name, const void* first_keys_in, const uint64_t* d_key_dst,
                     const uint64_t* d_val_dst) {
	if (n == 0) return 0;
	if (n > radix_max_items()) throw Error(4, "radix sort: more than 2^30-1 items in one device sort");
	const uint32_t n_tiles = (uint32_t)((n + kTile - 1) / kTile);
	const int P = plan.n_passes;
	DevBuf<uint32_t> bin_base(c, (size_t)P * kRadix);
	const size_t status_words = (size_t)n_tiles * kRadix;
	DevBuf<uint32_t> status(c, status_words * P + P);  // + one ticket per pass
	{
		CopyScope zs(c, "copy_zero_sort_state", (double)(status_words * P + P) * 4);
		MEMS_CUDA(cudaMemsetAsync(status.p, 0, (status_words * P + P) * sizeof(uint32_t), c->stream));
	}
	{
		KernelScope ks(c, "scan_bins");
		scan_bins_kernel<<<P, kRadix, 0, c->stream>>>(d_hist, bin_base.p);
		MEMS_CUDA(cudaGetLastError());
	}
	const size_t key_bytes = key64 ? 8 : 4;
	const size_t smem = (key_bytes + 4) * kTile;
	// launch configuration (measured on B200, tools/sort_bench.py): 256 threads x 16 pairs, 5 CTAs per SM for
	// 32-bit keys and 4 for 64-bit keys
	auto launch = [&](auto kern, int threads, int q, int cur) {
		MEMS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		uint32_t* st = status.p + status_words * q;
		uint32_t* ticket = status.p + status_words * P + q;
		const uint32_t mask = (1u << plan.bits[q]) - 1u;
		using K = std::remove_const_t<std::remove_pointer_t<decltype(first_arg_of(kern))>>;
		kern<<<n_tiles, threads, smem, c->stream>>>((const K*)(q == 0 && first_keys_in ? first_keys_in : d_keys[cur]), d_vals[cur],
		                                            (K*)d_keys[cur ^ 1], d_vals[cur ^ 1], (uint32_t)n, plan.shift[q], mask,
		                                            bin_base.p + q * kRadix, st, ticket, d_key_dst, d_val_dst);
		MEMS_CUDA(cudaGetLastError());
	};
	int cur = 0;
	for (int q = 0; q < P; ++q) {
		KernelScope ks(c, prof_name, 2.0 * (double)n * (double)(key_bytes + 4));
		const bool eight = plan.bits[q] > 7;
		if (d_key_dst) {  // single-pass partition whose output goes to the peers' exchange windows
			if (key64) {
				if (eight) launch(onesweep_kernel<uint64_t, 256, kMinB64, 8, 4, true>, 256, q, cur);
				else launch(onesweep_kernel<uint64_t, 256, kMinB64, 7, 4, true>, 256, q, cur);
			} else {
				if (eight) launch(onesweep_kernel<uint32_t, 256, kMinB32, 8, 4, true>, 256, q, cur);
				else launch(onesweep_kernel<uint32_t, 256, kMinB32, 7, 4, true>, 256, q, cur);
			}
		} else if (key64) {
			if (eight) launch(onesweep_kernel<uint64_t, 256, kMinB64, 8>, 256, q, cur);
			else launch(onesweep_kernel<uint64_t, 256, kMinB64, 7>, 256, q, cur);
		} else {
			if (eight) launch(onesweep_kernel<uint32_t, 256, kMinB32, 8>, 256, q, cur);
			else launch(onesweep_kernel<uint32_t, 256, kMinB32, 7>, 256, q, cur);
		}
		cur ^= 1;
	}
	return cur;
}

// ------------------------------------------------------------------------------------------------
// standalone digit histograms (inputs that did not come out of launch_extract)
template <class KeyT>
__global__ void __launch_bounds__(256)
histogram_kernel(const KeyT* __restrict__ keys, uint64_t n, SortPlan plan, uint32_t* __restrict__ hist) {
	__shared__ uint32_t s_hist[10 * kRadix];
	for (int i = threadIdx.x; i < plan.n_passes * kRadix; i += blockDim.x) s_hist[i] = 0;
	__syncthreads();
	uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		KeyT k = keys[i];
		for (int q = 0; q < plan.n_passes; ++q)
			atomicAdd(&s_hist[q * kRadix + ((uint32_t)(k >> plan.shift[q]) & ((1u << plan.bits[q]) - 1u))], 1u);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < plan.n_passes * kRadix; i += blockDim.x) {
		uint32_t v = s_hist[i];
		if (v) atomicAdd(&hist[i], v);
	}
}

void launch_histogram(Ctx* c, bool key64, const void* d_keys, uint64_t n, const SortPlan& plan, uint32_t* d_hist) {
	MEMS_CUDA(cudaMemsetAsync(d_hist, 0, (size_t)plan.n_passes * kRadix * sizeof(uint32_t), c->stream));
	if (n == 0) return;
	uint64_t want = (n + 256 * 16 - 1) / (256 * 16);
	unsigned grid = (unsigned)(want < (uint64_t)c->sm_count * 8 ? want : (uint64_t)c->sm_count * 8);
	KernelScope ks(c, "histogram", (double)n * (key64 ? 8.0 : 4.0));
	if (key64)
		histogram_kernel<uint64_t><<<grid, 256, 0, c->stream>>>((const uint64_t*)d_keys, n, plan, d_hist);
	else
		histogram_kernel<uint32_t><<<grid, 256, 0, c->stream>>>((const uint32_t*)d_keys, n, plan, d_hist);
	MEMS_CUDA(cudaGetLastError());
}

}  // namespace mems
