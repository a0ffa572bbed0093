// context.cu — per-thread execution context: stream, stream-ordered memory pool, launch profiler.
#include <cstring>

#include "common.cuh"
#include "mems_b200.h"

namespace mems {

void* Ctx::alloc(size_t bytes) {
	void* p = nullptr;
	MEMS_CUDA(cudaMallocAsync(&p, bytes ? bytes : 1, stream));
	return p;
}

void Ctx::free(void* p) {
	if (p) cudaFreeAsync(p, stream);
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
	MEMS_CUDA(cudaHostAlloc((void**)&p, 16 * sizeof(uint32_t), cudaHostAllocDefault));
	return p;
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
	for (auto& kv : prof)
		for (auto& pr : kv.second.pending) {
			cudaEventDestroy(pr.first);
			cudaEventDestroy(pr.second);
		}
	for (auto e : free_events) cudaEventDestroy(e);
	for (auto& pb : pinned_free) cudaFreeHost(pb.first);
	for (uint32_t* p : host_words_free) cudaFreeHost(p);
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
