// radix_sort.cu — stable LSD radix sort of (key, u32 value) pairs, one "onesweep" kernel per digit.
//
// Replaces std::sort over vector<bmer> in MemorySML::Create (MemorySML.cpp:54, comparator
// SortedMerList.h:312-314).  The reference sorts 16-byte bmer records by the full 64-bit mer; here the
// key is the compact canonical key (2w+1 bits, see seed_dev.cuh) whose order is identical, so only
// ceil((2w+1)/8) digit passes run.  Stability makes ties (which std::sort leaves unspecified) come out
// in ascending position order.
//
// Per pass, one CTA per tile of kTile pairs:
//   1. coalesced warp-striped load of the tile (each warp owns a contiguous 32*kItems chunk),
//   2. per-warp digit ranking with __match_any_sync (no atomics, order preserving),
//   3. per-digit decoupled look-back across tiles (thread d resolves digit d) against the digit
//      histogram that extraction already produced — a tile never waits for more than its predecessors'
//      256 counters, so the whole pass is a single read and a single write of the data,
//   4. reorder through shared memory so every digit's run leaves the CTA as one contiguous segment.
// Tiles take their index from an atomic ticket, which guarantees that all predecessors of a running
// tile are themselves running or finished (forward progress of the look-back).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace mems {

constexpr int kRadix = 256;
constexpr int kTile = 4096;  // pairs per CTA

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

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
	uint32_t v;
	asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
	asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// THREADS x ITEMS = kTile.  The kernel is bound by the latency of its shared-memory ranking chain, not by
// HBM (ncu, profiles/r01_*): what buys throughput is resident warps, so the register footprint is kept
// small — ranks are 16-bit, values are not loaded until the reorder step — and MINB CTAs share an SM.
// lanes of the warp holding the same 8-bit digit.  MATCH.ANY is a single instruction but occupies its pipe for
// tens of cycles; eight VOTEs plus a little logic (the digit is <= 8 bits) issue at full rate.
template <bool BALLOT>
__device__ __forceinline__ uint32_t digit_peers(uint32_t d) {
	if (!BALLOT) return __match_any_sync(0xffffffffu, d);
	uint32_t peers = 0xffffffffu;
#pragma unroll
	for (int b = 0; b < 8; ++b) {
		const bool bit = (d >> b) & 1u;
		const uint32_t m = __ballot_sync(0xffffffffu, bit);
		peers &= bit ? m : ~m;
	}
	return peers;
}

template <class KeyT, int THREADS, int MINB, bool BALLOT = false>
__global__ void __launch_bounds__(THREADS, MINB)
onesweep_kernel(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, KeyT* __restrict__ keys_out,
                uint32_t* __restrict__ vals_out, uint32_t n, int shift, uint32_t digit_mask,
                const uint32_t* __restrict__ bin_base, uint32_t* __restrict__ status, uint32_t* __restrict__ ticket) {
	constexpr int WARPS = THREADS / 32;
	constexpr int ITEMS = kTile / THREADS;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	KeyT* s_keys = reinterpret_cast<KeyT*>(smem_raw);
	uint32_t* s_vals = reinterpret_cast<uint32_t*>(smem_raw + sizeof(KeyT) * kTile);
	__shared__ uint16_t s_warp_cnt[WARPS][kRadix];  // per-warp digit counts (<= 32*ITEMS), then exclusive over warps
	__shared__ uint32_t s_digit_excl[kRadix];
	__shared__ uint32_t s_global_base[kRadix];
	__shared__ uint32_t s_warp_tot[8];
	__shared__ uint32_t s_tile;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) s_tile = atomicAdd(ticket, 1u);
	for (int i = lane; i < kRadix; i += 32) s_warp_cnt[warp][i] = 0;
	__syncthreads();
	const uint32_t tile = s_tile;
	const uint32_t tile_base = tile * (uint32_t)kTile;
	const uint32_t n_valid = n - tile_base < (uint32_t)kTile ? n - tile_base : (uint32_t)kTile;

	KeyT key[ITEMS];
	uint16_t rank[ITEMS];
	const uint32_t warp_base = warp * (32 * ITEMS);
#pragma unroll
	for (int i = 0; i < ITEMS; ++i) {
		const uint32_t local = warp_base + i * 32 + lane;
		key[i] = local < n_valid ? keys_in[tile_base + local] : ~(KeyT)0;
	}
	// ---- rank inside the warp, in memory order
	const uint32_t lanemask_lt = (1u << lane) - 1u;
#pragma unroll
	for (int i = 0; i < ITEMS; ++i) {
		const uint32_t local = warp_base + i * 32 + lane;
		const uint32_t d = local < n_valid ? ((uint32_t)(key[i] >> shift) & digit_mask) : (uint32_t)(kRadix - 1);
		const uint32_t peers = digit_peers<BALLOT>(d);
		const int leader = __ffs(peers) - 1;
		uint32_t base = 0;
		if (lane == leader) {
			base = s_warp_cnt[warp][d];
			s_warp_cnt[warp][d] = (uint16_t)(base + __popc(peers));
		}
		base = __shfl_sync(0xffffffffu, base, leader);
		rank[i] = (uint16_t)(base + __popc(peers & lanemask_lt));
		__syncwarp();
	}
	__syncthreads();
	// ---- thread d owns digit d: exclusive scan over warps, tile count, look-back
	uint32_t count = 0;
	if (tid < kRadix) {
		uint32_t sum = 0;
#pragma unroll
		for (int w = 0; w < WARPS; ++w) {
			const uint32_t c = s_warp_cnt[w][tid];
			s_warp_cnt[w][tid] = (uint16_t)sum;
			sum += c;
		}
		// padding items of the last tile were ranked as digit 255, after every real item
		count = sum - (tid == kRadix - 1 ? (uint32_t)kTile - n_valid : 0u);
		uint32_t incl = count;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += t;
		}
		if (lane == 31) s_warp_tot[warp] = incl;
		s_digit_excl[tid] = incl - count;  // exclusive inside this warp's 32 digits; warp offsets added below
	}
	__syncthreads();
	if (tid < kRadix) {
		uint32_t woff = 0;
#pragma unroll
		for (int w = 0; w < kRadix / 32; ++w)
			if (w < warp) woff += s_warp_tot[w];
		const uint32_t digit_excl = s_digit_excl[tid] + woff;
		uint32_t excl = 0;
		uint32_t* my_status = status + (size_t)tile * kRadix + tid;
		if (tile == 0) {
			st_volatile_u32(my_status, count | kFlagInclusive);
		} else {
			st_volatile_u32(my_status, count | kFlagPartial);
			const uint32_t* look = my_status - kRadix;
			while (true) {
				uint32_t s;
				do {
					s = ld_volatile_u32(look);
				} while ((s & kFlagMask) == 0u);
				excl += s & kValueMask;
				if (s & kFlagInclusive) break;
				look -= kRadix;
			}
			st_volatile_u32(my_status, (excl + count) | kFlagInclusive);
		}
		s_digit_excl[tid] = digit_excl;
		s_global_base[tid] = bin_base[tid] + excl - digit_excl;  // + tile-local sorted index = global index
	}
	__syncthreads();
	// ---- reorder through shared memory (values are loaded only now: one register live per value)
#pragma unroll
	for (int i = 0; i < ITEMS; ++i) {
		const uint32_t local = warp_base + i * 32 + lane;
		const bool valid = local < n_valid;
		const uint32_t d = valid ? ((uint32_t)(key[i] >> shift) & digit_mask) : (uint32_t)(kRadix - 1);
		const uint32_t pos = s_digit_excl[d] + s_warp_cnt[warp][d] + rank[i];
		s_keys[pos] = key[i];
		s_vals[pos] = valid ? vals_in[tile_base + local] : 0u;
	}
	__syncthreads();
#pragma unroll
	for (int k = 0; k < ITEMS; ++k) {
		const uint32_t idx = k * THREADS + tid;
		if (idx < n_valid) {
			const KeyT kk = s_keys[idx];
			const uint32_t d = (uint32_t)(kk >> shift) & digit_mask;
			const uint32_t g = s_global_base[d] + idx;
			keys_out[g] = kk;
			vals_out[g] = s_vals[idx];
		}
	}
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
                                      uint32_t*)) {
	return nullptr;
}

int radix_sort_pairs(Ctx* c, bool key64, void* d_keys[2], uint32_t* d_vals[2], uint64_t n, const SortPlan& plan,
                     uint32_t* d_hist, const char* prof_name, const void* first_keys_in) {
	if (n == 0) return 0;
	if (n > radix_max_items()) throw Error(4, "radix sort: more than 2^30-1 items in one device sort");
	const uint32_t n_tiles = (uint32_t)((n + kTile - 1) / kTile);
	const int P = plan.n_passes;
	DevBuf<uint32_t> bin_base(c, (size_t)P * kRadix);
	const size_t status_words = (size_t)n_tiles * kRadix;
	DevBuf<uint32_t> status(c, status_words * P + P);  // + one ticket per pass
	MEMS_CUDA(cudaMemsetAsync(status.p, 0, (status_words * P + P) * sizeof(uint32_t), c->stream));
	{
		KernelScope ks(c, "scan_bins");
		scan_bins_kernel<<<P, kRadix, 0, c->stream>>>(d_hist, bin_base.p);
		MEMS_CUDA(cudaGetLastError());
	}
	const size_t key_bytes = key64 ? 8 : 4;
	const size_t smem = (key_bytes + 4) * kTile;
	// launch configuration (measured on B200, tools/sort_bench.py): 256 threads x 16 pairs, 5 CTAs per SM for
	// 32-bit keys and 4 for 64-bit keys; MEMS_SORT_VARIANT selects the alternatives for experiments
	static int variant = -1;
	if (variant < 0) {
		const char* e = getenv("MEMS_SORT_VARIANT");
		variant = e ? atoi(e) : 0;
	}
	auto launch = [&](auto kern, int threads, int q, int cur) {
		MEMS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		uint32_t* st = status.p + status_words * q;
		uint32_t* ticket = status.p + status_words * P + q;
		const uint32_t mask = (1u << plan.bits[q]) - 1u;
		using K = std::remove_const_t<std::remove_pointer_t<decltype(first_arg_of(kern))>>;
		kern<<<n_tiles, threads, smem, c->stream>>>((const K*)(q == 0 && first_keys_in ? first_keys_in : d_keys[cur]), d_vals[cur],
		                                            (K*)d_keys[cur ^ 1], d_vals[cur ^ 1], (uint32_t)n, plan.shift[q], mask,
		                                            bin_base.p + q * kRadix, st, ticket);
		MEMS_CUDA(cudaGetLastError());
	};
	int cur = 0;
	for (int q = 0; q < P; ++q) {
		KernelScope ks(c, prof_name, 2.0 * (double)n * (double)(key_bytes + 4));
		if (key64) {
			if (variant == 1) launch(onesweep_kernel<uint64_t, 512, 3, true>, 512, q, cur);
			else if (variant == 2) launch(onesweep_kernel<uint64_t, 256, 3, true>, 256, q, cur);
			else launch(onesweep_kernel<uint64_t, 256, 4, true>, 256, q, cur);
		} else {
			if (variant == 1) launch(onesweep_kernel<uint32_t, 512, 3, true>, 512, q, cur);
			else if (variant == 2) launch(onesweep_kernel<uint32_t, 256, 4, true>, 256, q, cur);
			else launch(onesweep_kernel<uint32_t, 256, 5, true>, 256, q, cur);
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

// ------------------------------------------------------------------------------------------------
// exclusive scan (used for stream compaction of hits / segments / output records)
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_tot, uint32_t* total) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t incl = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += t;
	}
	if (lane == 31) s_tot[warp] = incl;
	__syncthreads();
	uint32_t woff = 0, tot = 0;
#pragma unroll
	for (int w = 0; w < kScanThreads / 32; ++w) {
		uint32_t t = s_tot[w];
		if (w < warp) woff += t;
		tot += t;
	}
	*total = tot;
	__syncthreads();
	return woff + incl - v;
}

__global__ void __launch_bounds__(kScanThreads)
scan_reduce_kernel(const uint32_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ partial) {
	__shared__ uint32_t s_tot[kScanThreads / 32];
	uint64_t base = (uint64_t)blockIdx.x * kScanTile;
	uint32_t sum = 0;
#pragma unroll
	for (int k = 0; k < kScanItems; ++k) {
		uint64_t i = base + (uint64_t)k * kScanThreads + threadIdx.x;
		if (i < n) sum += in[i];
	}
	uint32_t tot;
	block_exclusive_scan(sum, s_tot, &tot);
	if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(const uint32_t* in, uint32_t* out, uint64_t n,  // in may alias out (in-place scan)
                  const uint32_t* partial_excl, uint32_t* total_out) {
	__shared__ uint32_t s_tot[kScanThreads / 32];
	uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;  // blocked: thread owns 8 in a row
	uint32_t v[kScanItems];
	uint32_t sum = 0;
#pragma unroll
	for (int k = 0; k < kScanItems; ++k) {
		uint64_t i = base + k;
		v[k] = i < n ? in[i] : 0u;
		sum += v[k];
	}
	uint32_t tot;
	uint32_t excl = block_exclusive_scan(sum, s_tot, &tot);
	uint32_t off = (partial_excl ? partial_excl[blockIdx.x] : 0u) + excl;
#pragma unroll
	for (int k = 0; k < kScanItems; ++k) {
		uint64_t i = base + k;
		if (i < n) out[i] = off;
		off += v[k];
	}
	if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) *total_out = off;
}

void exclusive_scan_u32(Ctx* c, const uint32_t* d_in, uint32_t* d_out, uint64_t n, uint32_t* d_total) {
	if (n == 0) {
		if (d_total) MEMS_CUDA(cudaMemsetAsync(d_total, 0, sizeof(uint32_t), c->stream));
		return;
	}
	uint64_t n_blocks = (n + kScanTile - 1) / kScanTile;
	if (n_blocks == 1) {
		KernelScope ks(c, "scan");
		scan_apply_kernel<<<1, kScanThreads, 0, c->stream>>>(d_in, d_out, n, nullptr, d_total);
		MEMS_CUDA(cudaGetLastError());
		return;
	}
	DevBuf<uint32_t> partial(c, n_blocks);
	{
		KernelScope ks(c, "scan");
		scan_reduce_kernel<<<(unsigned)n_blocks, kScanThreads, 0, c->stream>>>(d_in, n, partial.p);
		MEMS_CUDA(cudaGetLastError());
	}
	exclusive_scan_u32(c, partial.p, partial.p, n_blocks, nullptr);
	{
		KernelScope ks(c, "scan");
		scan_apply_kernel<<<(unsigned)n_blocks, kScanThreads, 0, c->stream>>>(d_in, d_out, n, partial.p, d_total);
		MEMS_CUDA(cudaGetLastError());
	}
}

}  // namespace mems
