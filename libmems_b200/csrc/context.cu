// context.cu — per-thread execution context: stream, stream-ordered device-memory arena, launch profiler.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "mems_b200.h"

namespace mems {

namespace {
// MEMS_TRACE_SLOW=<ms>: report host-side calls of the allocator that take longer (diagnosis of step-time outliers)
double slow_ms() {
	static const double v = getenv("MEMS_TRACE_SLOW") ? atof(getenv("MEMS_TRACE_SLOW")) : 0.0;
	return v;
}
struct SlowCall {
	const char* what;
	size_t bytes;
	std::chrono::steady_clock::time_point t0;
	SlowCall(const char* w, size_t b) : what(w), bytes(b) {
		if (slow_ms() > 0) t0 = std::chrono::steady_clock::now();
	}
	~SlowCall() {
		if (slow_ms() > 0) {
			const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
			if (ms >= slow_ms()) fprintf(stderr, "[mems slow] %s(%zu bytes) took %.3f ms\n", what, bytes, ms);
		}
	}
};
}  // namespace

// Device memory: an arena owned by the context.  Every buffer of a call lives and dies in stream order on the context's
// one stream, so a host-side best-fit allocator over a few cudaMalloc'ed slabs is exact (a block freed here is reused only
// by work enqueued later on the same stream) and a repeated workload makes no driver call at all in the steady state.
// cudaMallocAsync's pool went back to the driver for the large blocks of every step on the pool's boxes — milliseconds each,
// and hundreds of milliseconds whenever another tenant of the node held the kernel driver's lock.
static constexpr size_t kSlabMin = 32u << 20, kSlabRound = 2u << 20;

void Arena::insert_free(char* p, size_t bytes) {
	free_by_addr[p] = bytes;
	free_by_size.insert({bytes, p});
}

void Arena::erase_free(std::map<char*, size_t>::iterator it) {
	auto range = free_by_size.equal_range(it->second);
	for (auto j = range.first; j != range.second; ++j)
		if (j->second == it->first) {
			free_by_size.erase(j);
			break;
		}
	free_by_addr.erase(it);
}

void Arena::add_slab(char* base, size_t bytes) {
	slabs.push_back({base, bytes});
	reserved += bytes;
	insert_free(base, bytes);
}

void* Arena::take(size_t bytes) {
	auto fit = free_by_size.lower_bound(bytes);  // best fit
	if (fit == free_by_size.end()) return nullptr;
	char* p = fit->second;
	const size_t have = fit->first;
	free_by_size.erase(fit);
	free_by_addr.erase(p);
	if (have > bytes) insert_free(p + bytes, have - bytes);
	used[p] = bytes;
	return p;
}

bool Arena::give(void* ptr) {
	auto u = used.find((char*)ptr);
	if (u == used.end()) return false;
	char* p = u->first;
	size_t bytes = u->second;
	used.erase(u);
	// merge with free neighbours of the same slab
	size_t si = 0;
	while (si < slabs.size() && !(p >= slabs[si].first && p < slabs[si].first + slabs[si].second)) ++si;
	char* lo = slabs[si].first;
	char* hi = lo + slabs[si].second;
	auto next = free_by_addr.lower_bound(p);
	if (next != free_by_addr.end() && next->first == p + bytes && next->first < hi) {
		bytes += next->second;
		erase_free(next);
	}
	auto prev = free_by_addr.lower_bound(p);
	if (prev != free_by_addr.begin()) {
		--prev;
		if (prev->first >= lo && prev->first + prev->second == p) {
			p = prev->first;
			bytes += prev->second;
			erase_free(prev);
		}
	}
	insert_free(p, bytes);
	return true;
}

std::vector<std::pair<char*, size_t>> Arena::drop_idle_slabs() {
	std::vector<std::pair<char*, size_t>> idle;
	for (size_t i = 0; i < slabs.size();) {
		auto it = free_by_addr.find(slabs[i].first);
		if (it != free_by_addr.end() && it->second == slabs[i].second) {
			erase_free(it);
			idle.push_back(slabs[i]);
			reserved -= slabs[i].second;
			slabs.erase(slabs.begin() + i);
		} else {
			++i;
		}
	}
	return idle;
}

void* Ctx::alloc(size_t bytes) {
	bytes = Arena::round_up(bytes);
	std::lock_guard<std::mutex> lock(arena_mutex);
	void* p = arena.take(bytes);
	if (!p) {  // grow: one more slab (cudaMalloc synchronises the device; warm-up only)
		const size_t slab = std::max(kSlabMin, (bytes + kSlabRound - 1) & ~(kSlabRound - 1));
		void* base = nullptr;
		SlowCall sc("cudaMalloc", slab);
		cudaError_t e = cudaMalloc(&base, slab);
		if (e == cudaErrorMemoryAllocation) {  // give idle slabs back (their last users may still run: wait for them) and retry
			cudaGetLastError();
			MEMS_CUDA(cudaStreamSynchronize(stream));
			for (auto& sl : arena.drop_idle_slabs()) cudaFree(sl.first);
			e = cudaMalloc(&base, slab);
		}
		MEMS_CUDA(e);
		arena.add_slab((char*)base, slab);
		p = arena.take(bytes);
	}
	return p;
}

void Ctx::free(void* ptr) {
	if (!ptr) return;
	std::lock_guard<std::mutex> lock(arena_mutex);
	arena.give(ptr);  // (not ours: ignored)
}

// Randomised self-check of the arena's bookkeeping on made-up addresses (no device needed): blocks never overlap,
// never leave their slab, everything given back merges into whole slabs again.  0 = passed.
int arena_selftest(uint64_t seed, int rounds) {
	Arena a;
	uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1;
	auto rnd = [&]() {
		x ^= x << 13;
		x ^= x >> 7;
		x ^= x << 17;
		return x;
	};
	char* const base = reinterpret_cast<char*>(uintptr_t(1) << 40);
	size_t next_slab = 0;
	std::map<char*, size_t> live;  // what the test holds
	auto check = [&]() -> int {
		// live blocks and free blocks tile every slab exactly
		size_t total = 0;
		for (auto& kv : live) total += kv.second;
		for (auto& kv : a.free_by_addr) total += kv.second;
		if (total != a.reserved) return 1;
		if (a.free_by_addr.size() != a.free_by_size.size() || a.used.size() != live.size()) return 2;
		std::map<char*, size_t> all(live);
		for (auto& kv : a.free_by_addr)
			if (!all.insert(kv).second) return 3;
		char* end = nullptr;
		for (auto& kv : all) {
			if (end && kv.first < end) return 4;  // overlap
			end = kv.first + kv.second;
		}
		// two free blocks never touch inside one slab
		for (auto it = a.free_by_addr.begin(); it != a.free_by_addr.end(); ++it) {
			auto nx = std::next(it);
			if (nx == a.free_by_addr.end() || it->first + it->second != nx->first) continue;
			bool boundary = false;
			for (auto& sl : a.slabs) boundary |= sl.first == nx->first;
			if (!boundary) return 5;
		}
		return 0;
	};
	for (int r = 0; r < rounds; ++r) {
		const bool want_alloc = live.size() < 4 || (rnd() % 100) < 55;
		if (want_alloc) {
			const size_t bytes = Arena::round_up((size_t)(rnd() % (rnd() % 8 == 0 ? (64u << 20) : (1u << 20))));
			void* p = a.take(bytes);
			if (!p) {
				const size_t slab = std::max<size_t>(32u << 20, (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1));
				a.add_slab(base + next_slab, slab);
				next_slab += slab;  // slabs adjacent in address space: merging must still stop at their borders
				p = a.take(bytes);
				if (!p) return 10;
			}
			if ((reinterpret_cast<uintptr_t>(p) & (Arena::kAlign - 1)) != 0) return 11;
			if (!live.insert({(char*)p, bytes}).second) return 12;
		} else {
			auto it = live.begin();
			std::advance(it, (long)(rnd() % live.size()));
			if (!a.give(it->first)) return 13;
			live.erase(it);
		}
		if ((r % 64) == 0)
			if (int e = check()) return 100 + e;
	}
	if (a.give(base - 4096)) return 14;  // a foreign pointer is refused
	for (auto& kv : live)
		if (!a.give(kv.first)) return 15;
	live.clear();
	if (int e = check()) return 200 + e;
	const size_t n_slabs = a.slabs.size();
	if (a.free_by_addr.size() != n_slabs) return 16;  // every slab is one free block again
	if (a.drop_idle_slabs().size() != n_slabs || a.reserved != 0 || !a.free_by_addr.empty()) return 17;
	return 0;
}

void* Ctx::pinned_get(size_t bytes, size_t* capacity) {
	size_t best = pinned_free.size();
	for (size_t i = 0; i < pinned_free.size(); ++i)
		if (pinned_free[i].second >= bytes && (best == pinned_free.size() || pinned_free[i].second < pinned_free[best].second))
			best = i;
	if (best != pinned_free.size()) {
		void* p = pinned_free[best].first;
		*capacity = pinned_free[best].second;
		pinned_free.erase(pinned_free.begin() + best);
		return p;
	}
	size_t cap = (bytes + (1u << 20)) & ~(size_t)((1u << 20) - 1);  // round up to 1 MiB
	void* p = nullptr;
	SlowCall sc("cudaHostAlloc", cap);
	MEMS_CUDA(cudaHostAlloc(&p, cap, cudaHostAllocDefault));
	*capacity = cap;
	return p;
}

void Ctx::pinned_put(void* p, size_t capacity) {
	if (pinned_free.size() >= 8) {  // keep the pool small
		cudaFreeHost(p);
		return;
	}
	pinned_free.push_back({p, capacity});
}

uint32_t* Ctx::host_words_get() {
	if (!host_words_free.empty()) {
		uint32_t* p = host_words_free.back();
		host_words_free.pop_back();
		return p;
	}
	uint32_t* p = nullptr;
	MEMS_CUDA(cudaHostAlloc((void**)&p, 16 * sizeof(uint32_t), cudaHostAllocMapped));
	return p;
}

__global__ void fetch_words_kernel(uint32_t* __restrict__ host_dst, const uint32_t* __restrict__ src, uint32_t n) {
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) host_dst[i] = src[i];
}

void Ctx::fetch_async(uint32_t* host_words, const uint32_t* d_src, uint32_t n_words) {
	launch_count++;
	fetch_words_kernel<<<1, 32, 0, stream>>>(host_words, d_src, n_words);  // (unified addressing: the mapped pointer is the device's too)
	MEMS_CUDA(cudaGetLastError());
}

void Ctx::fetch(void* dst, const void* d_src, size_t bytes) {
	const size_t words = (bytes + 3) / 4;
	if (words > fetch_stage_words) {
		if (fetch_stage) {
			MEMS_CUDA(cudaStreamSynchronize(stream));
			cudaFreeHost(fetch_stage);
			fetch_stage = nullptr;
		}
		const size_t cap = std::max<size_t>(words, 16384);
		MEMS_CUDA(cudaHostAlloc((void**)&fetch_stage, cap * 4, cudaHostAllocMapped));
		fetch_stage_words = cap;
	}
	// (sources are 4-byte aligned device arrays; a trailing partial word reads inside the arena's 512-byte granule)
	launch_count++;
	fetch_words_kernel<<<(unsigned)std::min<size_t>((words + 255) / 256, 64), 256, 0, stream>>>(fetch_stage, (const uint32_t*)d_src, (uint32_t)words);
	MEMS_CUDA(cudaGetLastError());
	MEMS_CUDA(cudaStreamSynchronize(stream));
	memcpy(dst, fetch_stage, bytes);
}

void Ctx::host_words_put(uint32_t* p) {
	if (p) host_words_free.push_back(p);
}

cudaEvent_t Ctx::get_event() {
	if (!free_events.empty()) {
		cudaEvent_t e = free_events.back();
		free_events.pop_back();
		return e;
	}
	cudaEvent_t e;
	MEMS_CUDA(cudaEventCreate(&e));
	return e;
}

void Ctx::prof_begin(const char* name, double bytes) {
	ProfEntry& pe = prof[name];
	cudaEvent_t a = get_event(), b = get_event();
	MEMS_CUDA(cudaEventRecord(a, stream));
	pe.pending.push_back({a, b});
	pe.pending_bytes.push_back(bytes);
}

void Ctx::prof_end(const char* name) {
	ProfEntry& pe = prof[name];
	cudaEventRecord(pe.pending.back().second, stream);
}

void Ctx::prof_collect() {
	MEMS_CUDA(cudaStreamSynchronize(stream));
	for (auto& kv : prof) {
		ProfEntry& pe = kv.second;
		for (size_t i = 0; i < pe.pending.size(); ++i) {
			float ms = 0;
			if (cudaEventElapsedTime(&ms, pe.pending[i].first, pe.pending[i].second) == cudaSuccess) {
				pe.ms += ms;
				pe.bytes += pe.pending_bytes[i];
				pe.launches++;
			}
			free_events.push_back(pe.pending[i].first);
			free_events.push_back(pe.pending[i].second);
		}
		pe.pending.clear();
		pe.pending_bytes.clear();
	}
}

Ctx::~Ctx() {
	cudaSetDevice(device);
	if (stream) cudaStreamSynchronize(stream);
	if (copy_stream) {
		cudaStreamSynchronize(copy_stream);
		cudaStreamDestroy(copy_stream);
	}
	for (auto& kv : prof)
		for (auto& pr : kv.second.pending) {
			cudaEventDestroy(pr.first);
			cudaEventDestroy(pr.second);
		}
	for (auto e : free_events) cudaEventDestroy(e);
	for (auto& pb : pinned_free) cudaFreeHost(pb.first);
	for (uint32_t* p : host_words_free) cudaFreeHost(p);
	if (fetch_stage) cudaFreeHost(fetch_stage);
	for (auto& sl : arena.slabs) cudaFree(sl.first);
	if (scan_state) cudaFree(scan_state);
	if (own_stream && stream) cudaStreamDestroy(stream);
}

// Pattern -> runs of ones (MSB of the pattern is window base 0, SortedMerList.cpp:726-762).
SeedDesc make_seed_desc(uint64_t seed) {
	SeedDesc sd;
	memset(&sd, 0, sizeof sd);
	sd.seed = seed;
	sd.L = mems_get_seed_length(seed);
	sd.w = mems_get_seed_weight(seed);
	if (sd.L == 0) throw Error(MEMS_ERR_INVALID, "SMLCreateError: Can't have 0 seed length");
	if (sd.L > 32) throw Error(MEMS_ERR_INVALID, "SMLCreateError: Mer size is too large");
	if (sd.L > 31)
		throw Error(MEMS_ERR_UNSUPPORTED,
		            "seed span 32 is not supported (the reference's RevCompMer shifts by a negative amount there, "
		            "SortedMerList.cpp:611; MAX_DNA_SEED_WEIGHT is 31)");
	if (!(seed & 1ull))
		throw Error(MEMS_ERR_INVALID,
		            "seed pattern must end in a 1 (GetSeedMer, SortedMerList.cpp:738-753, reads bits L-1..0 of the "
		            "pattern, so the reference itself mis-extracts patterns with trailing zeros)");
	uint64_t pat = seed;  // bit L-1 is window base 0
	int ones_before = 0;
	int i = 0;
	while (i < sd.L) {
		if (!((pat >> (sd.L - 1 - i)) & 1)) {
			++i;
			continue;
		}
		int a = i;
		while (i < sd.L && ((pat >> (sd.L - 1 - i)) & 1)) ++i;
		int b = i - 1, len = b - a + 1;
		if (sd.n_runs >= kMaxSeedRuns) throw Error(MEMS_ERR_UNSUPPORTED, "seed pattern has too many runs");
		sd.run_rshift[sd.n_runs] = (uint8_t)(62 - 2 * b);
		sd.run_bits[sd.n_runs] = (uint8_t)(2 * len);
		sd.run_lshift[sd.n_runs] = (uint8_t)(2 * (sd.w - ones_before - len));
		sd.run_net[sd.n_runs] = (uint8_t)(sd.run_rshift[sd.n_runs] - sd.run_lshift[sd.n_runs]);
		sd.run_mask[sd.n_runs] = ((len >= 32 ? ~0ull : ((1ull << (2 * len)) - 1ull))) << sd.run_lshift[sd.n_runs];
		sd.n_runs++;
		ones_before += len;
	}
	sd.key_bits = 2 * sd.w + 1;
	int n_off = 0;
	for (int o = 0; o < sd.L; ++o)
		if ((pat >> (sd.L - 1 - o)) & 1) sd.off[n_off++] = (uint8_t)o;
	sd.palindromic = 1;
	for (int k = 0; k < sd.w; ++k) {
		sd.mirror[k] = (uint8_t)(sd.L - 1 - sd.off[sd.w - 1 - k]);
		if (sd.mirror[k] != sd.off[k]) sd.palindromic = 0;
	}
	return sd;
}

}  // namespace mems
