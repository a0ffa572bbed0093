// scan.cu — exclusive prefix sums of 32-bit counts, one launch per scan (stream compaction of hits, segments, components,
// output records).  No counterpart in the reference: its table inserts matches one by one (MemHash.cpp:209-251); here
// every data-parallel stage sizes and places its output with a scan.
#include <algorithm>

#include "common.cuh"

namespace mems {

// ------------------------------------------------------------------------------------------------
// exclusive scan (stream compaction of hits / segments / output records): ONE launch per scan.
// Chained tiles with decoupled look-back: a tile publishes its sum, warp 0 polls 32 predecessors at a time until it
// meets an inclusive prefix.  The tile states live in a buffer the context keeps across calls; every scan stamps its
// states with a fresh epoch, so stale words of earlier scans read as "not ready" and nothing has to be zeroed, and tiles
// take their index from a ticket counter whose start value the host tracks (predecessors of a running tile are
// themselves running or done).  The three-kernel form this replaces (reduce / scan the sums / apply) cost 3-5
// launches per scan, 12 of the ~50 launches of a match-finding call.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;
__global__ void __launch_bounds__(kScanThreads)
scan_chained_kernel(const uint32_t* in, uint32_t* out, uint64_t n,  // in may alias out (in-place scan)
                    uint64_t* state, uint32_t ticket_base, uint32_t epoch, uint32_t* total_out) {
	__shared__ uint32_t s_tot[kScanThreads / 32];
	__shared__ uint32_t s_tile, s_excl;
	const int tid = threadIdx.x, lane = tid & 31;
	if (tid == 0) s_tile = atomicAdd(reinterpret_cast<uint32_t*>(state), 1u) - ticket_base;
	__syncthreads();
	const uint32_t tile = s_tile;
	uint64_t* status = state + 1;
	const uint64_t base = (uint64_t)tile * kScanTile + (uint64_t)tid * kScanItems;  // blocked: a thread owns 16 in a row
	uint32_t v[kScanItems];
	if (base + kScanItems <= n && (reinterpret_cast<uintptr_t>(in) & 15u) == 0) {
#pragma unroll
		for (int k = 0; k < kScanItems; k += 4) {
			const uint4 q = *reinterpret_cast<const uint4*>(in + base + k);
			v[k] = q.x; v[k + 1] = q.y; v[k + 2] = q.z; v[k + 3] = q.w;
		}
	} else {
#pragma unroll
		for (int k = 0; k < kScanItems; ++k) v[k] = base + k < n ? in[base + k] : 0u;
	}
	uint32_t sum = 0;
#pragma unroll
	for (int k = 0; k < kScanItems; ++k) sum += v[k];
	// block-wide exclusive scan of the thread sums
	uint32_t incl = sum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += t;
	}
	if (lane == 31) s_tot[tid >> 5] = incl;
	__syncthreads();
	uint32_t woff = 0, tot = 0;
#pragma unroll
	for (int w = 0; w < kScanThreads / 32; ++w) {
		const uint32_t t = s_tot[w];
		if (w < (tid >> 5)) woff += t;
		tot += t;
	}
	if (tid < 32) {  // warp 0: publish, look back
		const uint32_t excl = chain_lookback(status, tile, epoch, tot, lane);
		if (lane == 0) s_excl = excl;
	}
	__syncthreads();
	uint32_t off = s_excl + woff + incl - sum;
	if (base + kScanItems <= n && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
#pragma unroll
		for (int k = 0; k < kScanItems; k += 4) {
			uint4 q;
			q.x = off; off += v[k];
			q.y = off; off += v[k + 1];
			q.z = off; off += v[k + 2];
			q.w = off; off += v[k + 3];
			*reinterpret_cast<uint4*>(out + base + k) = q;
		}
	} else {
#pragma unroll
		for (int k = 0; k < kScanItems; ++k) {
			if (base + k < n) out[base + k] = off;
			off += v[k];
		}
	}
	if (total_out && tid == kScanThreads - 1 && (uint64_t)(tile + 1) * kScanTile >= n) *total_out = off;
}

ChainTicket reserve_chain_tiles(Ctx* c, uint64_t n_tiles) {
	if (n_tiles + 1 > c->scan_cap || c->scan_epoch >= (1u << 30) - 1u) {  // (re)create the tile states: zero = no epoch
		const size_t cap = std::max<size_t>(n_tiles + 1, std::max<size_t>(c->scan_cap, 1u << 16));
		MEMS_CUDA(cudaStreamSynchronize(c->stream));
		if (c->scan_state) cudaFree(c->scan_state);
		c->scan_state = nullptr;
		MEMS_CUDA(cudaMalloc((void**)&c->scan_state, cap * sizeof(uint64_t)));
		MEMS_CUDA(cudaMemsetAsync(c->scan_state, 0, cap * sizeof(uint64_t), c->stream));
		c->scan_cap = cap;
		c->scan_epoch = 0;
		c->scan_ticket_base = 0;
	}
	ChainTicket t{c->scan_state, c->scan_ticket_base, ++c->scan_epoch};
	c->scan_ticket_base += (uint32_t)n_tiles;
	return t;
}

void exclusive_scan_u32(Ctx* c, const uint32_t* d_in, uint32_t* d_out, uint64_t n, uint32_t* d_total) {
	if (n == 0) {
		if (d_total) MEMS_CUDA(cudaMemsetAsync(d_total, 0, sizeof(uint32_t), c->stream));
		return;
	}
	const uint64_t n_tiles = (n + kScanTile - 1) / kScanTile;
	const ChainTicket t = reserve_chain_tiles(c, n_tiles);
	KernelScope ks(c, "scan");
	scan_chained_kernel<<<(unsigned)n_tiles, kScanThreads, 0, c->stream>>>(d_in, d_out, n, t.state, t.ticket_base, t.epoch, d_total);
	MEMS_CUDA(cudaGetLastError());
}

}  // namespace mems
