// common.cuh — internal declarations shared by the CUDA translation units of libmems_b200.
// Nothing here is part of the public ABI (see include/mems_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <stdint.h>
#include <map>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <stdexcept>
#include <string>
#include <vector>

namespace mems {

constexpr int kMaxSeedRuns = 16;  // a pattern of span <= 31 has at most 16 runs of ones

// Spaced-seed pattern decomposed into its runs of ones: extraction is a software PEXT,
// one shift+mask+shift per run (SortedMerList::GetSeedMer, SortedMerList.cpp:726-762, does it bit-serially).
struct SeedDesc {
	uint64_t seed;
	int32_t L;       // span (getSeedLength)
	int32_t w;       // weight (getSeedWeight)
	int32_t n_runs;
	int32_t key_bits;  // 2w+1: compact sort key = (canonical w-mer << 1) | strand
	uint8_t run_rshift[kMaxSeedRuns];  // window >> rshift brings the run's last base to bits 1..0
	uint8_t run_bits[kMaxSeedRuns];    // 2 * run length
	uint8_t run_lshift[kMaxSeedRuns];  // position of the run inside the right-justified w-mer
	// the same runs as one shift + one mask each: mer |= (window >> run_net[r]) & run_mask[r]
	uint8_t run_net[kMaxSeedRuns];     // rshift - lshift (always >= 64 - 2L >= 2)
	uint64_t run_mask[kMaxSeedRuns];   // ((1 << run_bits) - 1) << run_lshift
	// the pattern as the window test of match extension reads it (kernels_match.cu):
	uint8_t off[32];     // offsets of the w cared bases inside a window, ascending
	uint8_t mirror[32];  // mirror[i] = L-1 - off[w-1-i]: where cared base i of a window sits, seen from the other strand
	int32_t palindromic; // mirror == off: a window and its reverse complement care about the same bases
	int32_t pad_;
};

// Every batch's packed buffer starts with kLeadWords zero words and ends with kTailWords: the window test loads 64-base
// chunks that may begin up to 63 bases before a sequence and end up to ~100 bases after it.
constexpr uint64_t kLeadWords = 8, kTailWords = 16;
// words one sequence of n bases occupies in a batch's packed buffer: its ceil(n/16) words + the reference's two pad
// words (SortedMerList.cpp:306-311), rounded up so that the next sequence starts 16-byte aligned
__host__ __device__ inline uint64_t seq_packed_words(uint64_t n_bases) { return ((n_bases + 15) / 16 + 2 + 3) / 4 * 4; }

// One sequence of a batch, as laid out in device memory.
struct SeqMeta {
	uint64_t word_off;  // first uint32 word of this sequence inside the batch's packed buffer
	uint64_t byte_off;  // first byte inside the batch's ASCII staging buffer
	uint64_t seed_off;  // first index inside the batch's union key/value arrays
	uint32_t n_bases;
	uint32_t n_seeds;   // SMLLength: n_bases - L + 1, or 0
	uint32_t tag;       // sequence id written into the values (index in the batch; global id when sharded)
	// A batch may hold several independent PROBLEMS (mems_find_matches_many): the sequences of one problem are a group.
	// The group number sits above the key's 2w+1 bits, so no equal-seed run spans two problems.
	uint32_t group;        // 0 for ordinary batches
	uint32_t group_first;  // tag of the group's first sequence
	uint32_t group_count;  // sequences in the group (= SeqCount of its matches)
};

struct Error : std::runtime_error {
	int code;
	Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define MEMS_CUDA(expr)                                                                          \
	do {                                                                                         \
		cudaError_t _e = (expr);                                                                 \
		if (_e != cudaSuccess)                                                                   \
			throw mems::Error(3, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +      \
			                         __FILE__ + ":" + std::to_string(__LINE__) + ")");             \
	} while (0)

// Bookkeeping of the context's device-memory arena (context.cu): blocks inside slabs, best fit, free neighbours of one
// slab merged.  Plain host code — it never touches the memory it hands out — so tests/test_abi_cpu.py can drive it without
// a GPU (mems_selftest_arena).
struct Arena {
	static constexpr size_t kAlign = 512;
	std::vector<std::pair<char*, size_t>> slabs;
	std::map<char*, size_t> free_by_addr;
	std::multimap<size_t, char*> free_by_size;
	std::unordered_map<char*, size_t> used;
	size_t reserved = 0;

	static size_t round_up(size_t bytes) { return ((bytes ? bytes : 1) + kAlign - 1) & ~(kAlign - 1); }
	void add_slab(char* base, size_t bytes);
	void* take(size_t bytes);  // bytes already rounded; nullptr: no free block is large enough
	bool give(void* p);        // false: not a block of this arena
	// slabs that are completely free leave the arena; the caller releases their memory
	std::vector<std::pair<char*, size_t>> drop_idle_slabs();

private:
	void insert_free(char* p, size_t bytes);
	void erase_free(std::map<char*, size_t>::iterator it);
};

int arena_selftest(uint64_t seed, int rounds);  // context.cu

struct ProfEntry {
	uint64_t launches = 0;
	double ms = 0, bytes = 0;
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
	std::vector<double> pending_bytes;
};

struct Ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	cudaStream_t copy_stream = nullptr;  // results leave on it behind the call's last kernel (FlatRecords::wait)
	int sm_count = 148;
	bool profiling = false;
	uint64_t launch_count = 0;
	std::map<std::string, ProfEntry> prof;
	std::vector<cudaEvent_t> free_events;
	std::string last_error;
	// test hooks (mems_test_hooks): 0 = production behaviour
	int test_hash_bits = 0;    // keep only this many bits of the diagonal hash (forces bucket collisions)
	int test_walk_budget = 0;  // probes / rounds before a walk moves on to the next larger walker
	double trace_slow_ms = 0;  // MEMS_TRACE_SLOW

	void* alloc(size_t bytes);  // stream-ordered on `stream`: arena of cudaMalloc'ed slabs, no driver call once warm
	void free(void* p);
	std::mutex arena_mutex;
	Arena arena;
	// tile states of the chained scans (exclusive_scan_u32): [0] ticket counter, [1..] one word per tile, stamped by epoch
	uint64_t* scan_state = nullptr;
	size_t scan_cap = 0;
	uint32_t scan_epoch = 0, scan_ticket_base = 0;
	// page-locked host staging buffers for results (D2H at full PCIe rate), recycled across calls
	std::vector<std::pair<void*, size_t>> pinned_free;
	void* pinned_get(size_t bytes, size_t* capacity);
	void pinned_put(void* p, size_t capacity);
	// small page-locked, device-mapped words for counters the host waits on
	std::vector<uint32_t*> host_words_free;
	uint32_t* host_words_get();  // 16 uint32
	void host_words_put(uint32_t* p);
	// Counters and small tables the host needs in the middle of a call (data-dependent sizes) are WRITTEN BY A KERNEL
	// into mapped page-locked memory instead of copied by the DMA engine: a cudaMemcpy of four bytes queues behind
	// whatever the device-to-host engine is busy with — the previous call's MatchList on the copy stream — and made
	// every round trip of the next call wait for that whole copy.
	void fetch_async(uint32_t* host_words, const uint32_t* d_src, uint32_t n_words);  // into words of host_words_get()
	void fetch(void* dst, const void* d_src, size_t bytes);  // synchronous: kernel into the staging area, wait, memcpy
	uint32_t* fetch_stage = nullptr;
	size_t fetch_stage_words = 0;
	void event_put(cudaEvent_t e) { free_events.push_back(e); }
	cudaEvent_t get_event();
	void prof_begin(const char* name, double bytes);
	void prof_end(const char* name);
	void prof_collect();
	~Ctx();
};

// RAII device buffer, stream-ordered on its context.
template <class T>
struct DevBuf {
	Ctx* ctx = nullptr;
	T* p = nullptr;
	size_t n = 0;
	DevBuf() {}
	DevBuf(Ctx* c, size_t count) : ctx(c), n(count) { p = (T*)c->alloc((count ? count : 1) * sizeof(T)); }
	DevBuf(const DevBuf&) = delete;
	DevBuf& operator=(const DevBuf&) = delete;
	DevBuf(DevBuf&& o) noexcept : ctx(o.ctx), p(o.p), n(o.n) { o.p = nullptr; }
	DevBuf& operator=(DevBuf&& o) noexcept {
		if (this != &o) {
			reset();
			ctx = o.ctx; p = o.p; n = o.n;
			o.p = nullptr;
		}
		return *this;
	}
	void reset() {
		if (p) ctx->free(p);
		p = nullptr;
		n = 0;
	}
	~DevBuf() { reset(); }
};

struct KernelScope {
	Ctx* c;
	const char* name;
	std::chrono::steady_clock::time_point t0;
	KernelScope(Ctx* ctx, const char* nm, double bytes = 0) : c(ctx), name(nm) {
		c->launch_count++;
		if (c->profiling) c->prof_begin(nm, bytes);
		if (c->trace_slow_ms > 0) t0 = std::chrono::steady_clock::now();
	}
	~KernelScope() {
		if (c->profiling) c->prof_end(name);
		if (c->trace_slow_ms > 0) {  // MEMS_TRACE_SLOW: a launch call that blocks the host (diagnosis of step-time outliers)
			const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
			if (ms >= c->trace_slow_ms) fprintf(stderr, "[mems slow] launch of %s took %.3f ms on the host\n", name, ms);
		}
	}
};

// a copy between host and device, timed like a kernel when profiling is on (names start with "copy_"; not a launch)
struct CopyScope {
	Ctx* c;
	const char* name;
	CopyScope(Ctx* ctx, const char* nm, double bytes) : c(ctx), name(nm) {
		if (c->profiling) c->prof_begin(nm, bytes);
	}
	~CopyScope() {
		if (c->profiling) c->prof_end(name);
	}
};

SeedDesc make_seed_desc(uint64_t seed);  // throws Error(MEMS_ERR_INVALID/UNSUPPORTED)

// ---- kernels_sml.cu ----
// ASCII -> 2-bit words for every sequence of a batch; sets *d_gap_flag if a '-' is seen.
void launch_pack(Ctx* c, const uint8_t* d_ascii, uint32_t* d_packed, const SeqMeta* d_meta, const SeqMeta* h_meta,
                 int n_seqs, uint32_t* d_gap_flag);
// Canonical compact keys + (seq << pos_bits | pos) values for every seed position of the batch, plus the
// digit histograms of all radix passes (n_passes x 256 counters, zeroed by the caller).
void launch_extract(Ctx* c, const uint32_t* d_packed, const SeqMeta* d_meta, const SeqMeta* h_meta, int n_seqs,
                    const SeedDesc& sd, int pos_bits, bool key64, void* d_keys, uint32_t* d_vals, uint32_t* d_hist,
                    int n_passes, const int* pass_shift, const int* pass_bits);  // keys carry SeqMeta::group above sd.key_bits
// Bit planes of a packed buffer: planes[i] = {high bits, low bits} of the 2-bit codes of bases 32i .. 32i+31
// (bit j of a word <-> base 32i + j); n_words (a multiple of 2) packed words -> n_words / 2 entries.
void launch_planes(Ctx* c, const uint32_t* d_packed, uint2* d_planes, uint64_t n_words);
// mers at arbitrary positions of one sequence (reference 64-bit layout)
void launch_seed_mers(Ctx* c, const uint32_t* d_words, uint32_t n_seeds, const SeedDesc& sd, const uint64_t* d_pos,
                      uint64_t n, uint64_t* d_fwd, uint64_t* d_dna);
// (position, reference-layout canonical mer) for sorted-list entries [offset, offset+count)
// (positions carry the sequence tag above pos_mask)
void launch_sml_read(Ctx* c, const uint32_t* d_words, const SeedDesc& sd, const uint32_t* d_positions, uint32_t pos_mask,
                     uint64_t count, uint64_t* d_mers);
// SortedMerList::bsearch (SortedMerList.cpp:380-394) over sorted-list entries [0, n): d_result[0] = index, [1] = found
void launch_find_mer(Ctx* c, const uint32_t* d_words, const SeedDesc& sd, const uint32_t* d_positions, uint32_t pos_mask,
                     uint64_t n, uint64_t query_mer, uint64_t* d_result);

// SeedOccurrenceList::construct for one sequence; d_positions = its sorted list (tagged), d_key_pos = its slice of
// the position-ordered keys; writes n_bases floats
void launch_seed_occurrence(Ctx* c, const uint32_t* d_positions, uint32_t pos_mask, const void* d_key_pos, bool key64,
                            uint32_t n_seeds, uint32_t n_bases, int L, float* d_out);

// ---- radix_sort.cu ----
struct SortPlan {
	int n_passes;
	int shift[10];
	int bits[10];
};
SortPlan make_sort_plan(int key_bits, int begin_bit = 0);
size_t radix_max_items();  // largest n one sort call accepts
// Stable LSD radix sort of (key, u32 value) pairs over plan's digits.  keys/vals are double buffers;
// d_hist holds n_passes x 256 digit counts of the input (computed by the caller, e.g. fused in extraction).
// Returns 0 or 1: the index of the buffer pair that holds the sorted result.  If first_keys_in is given, the
// first pass reads its keys from there (and leaves that array untouched) instead of d_keys[0].
int radix_sort_pairs(Ctx* c, bool key64, void* d_keys[2], uint32_t* d_vals[2], uint64_t n, const SortPlan& plan,
                     uint32_t* d_hist, const char* prof_name, const void* first_keys_in = nullptr,
                     const uint64_t* d_key_dst = nullptr, const uint64_t* d_val_dst = nullptr);
// d_key_dst / d_val_dst (optional, single-pass plans only): 256 byte addresses, one per digit; digit d's run is
// written to address[d] + g * element size (g = index in the partitioned order) instead of the second buffers —
// the addresses may lie in other GPUs' memory (exchange windows), which makes the pass the send side of an all-to-all
// standalone digit histograms of a key array (for inputs not produced by launch_extract)
void launch_histogram(Ctx* c, bool key64, const void* d_keys, uint64_t n, const SortPlan& plan, uint32_t* d_hist);
// exclusive prefix sum of n u32 values (in -> out, may alias); *d_total (optional, device) gets the sum
void exclusive_scan_u32(Ctx* c, const uint32_t* d_in, uint32_t* d_out, uint64_t n, uint32_t* d_total);

// Chained tiles with decoupled look-back (the scans, the run/hit kernel): the tile states live in Ctx::scan_state and are
// stamped with an epoch per launch, so nothing is zeroed between launches; tiles take their index from the ticket counter
// in state[0], whose start value the host tracks.
struct ChainTicket {
	uint64_t* state;  // [0] ticket counter, [1 + tile] status word: epoch << 34 | flag << 32 | value
	uint32_t ticket_base, epoch;
};
ChainTicket reserve_chain_tiles(Ctx* c, uint64_t n_tiles);
#ifdef __CUDACC__
constexpr uint64_t kScanPartial = 1ull << 32, kScanInclusive = 2ull << 32;

__device__ __forceinline__ uint64_t ld_state(const uint64_t* p) {
	uint64_t v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_state(uint64_t* p, uint64_t v) {
	asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// exclusive prefix of this tile's total over all earlier tiles, by warp 0 of the CTA (all 32 lanes call); publishes the
// tile's own state.  Returns the prefix in every lane of the warp.
__device__ __forceinline__ uint32_t chain_lookback(uint64_t* status, uint32_t tile, uint32_t epoch, uint32_t tot, int lane) {
	const uint64_t stamp = (uint64_t)epoch << 34;
	if (lane == 0) st_state(status + tile, stamp | (tile == 0 ? kScanInclusive : kScanPartial) | tot);
	uint32_t excl = 0;
	if (tile != 0) {
		int64_t look = (int64_t)tile - 1;
		for (;;) {
			const int64_t idx = look - lane;
			const uint64_t s = idx >= 0 ? ld_state(status + idx) : (stamp | kScanInclusive);
			const uint32_t flag = (s >> 34) == (uint64_t)epoch ? (uint32_t)(s >> 32) & 3u : 0u;
			const uint32_t ready = __ballot_sync(0xffffffffu, flag != 0u);
			const uint32_t inclusive = __ballot_sync(0xffffffffu, flag == 2u);
			const uint32_t need = inclusive ? (2u << (__ffs((int)inclusive) - 1)) - 1u : 0xffffffffu;  // lanes up to the first inclusive
			if ((ready & need) != need) continue;  // a predecessor in reach has not published yet: poll again
			excl += __reduce_add_sync(0xffffffffu, (need >> lane) & 1u ? (uint32_t)s : 0u);
			if (inclusive) break;
			look -= 32;
		}
		if (lane == 0) st_state(status + tile, stamp | kScanInclusive | (uint64_t)(uint32_t)(excl + tot));
	}
	return excl;
}
#endif


// ---- batch.cu ----
// All sequences handed to one create call: packed sequences, the union of their seeds sorted by key, and
// (lazily) the per-sequence sorted position lists.
struct Batch {
	std::shared_ptr<Ctx> ctx;
	SeedDesc sd;
	int n_seqs = 0;
	int pos_bits = 0;  // value = (seq << pos_bits) | position
	int seq_bits = 0;
	int group_bits = 0;  // key = (group << sd.key_bits) | compact key; 0 unless the batch holds several problems
	int n_groups = 1, max_group = 0;  // problems in the batch, sequences in the largest
	bool key64 = false;
	int sort_bits() const { return sd.key_bits + group_bits; }
	uint64_t n_total = 0;  // seeds in the union
	std::vector<SeqMeta> meta;
	DevBuf<SeqMeta> d_meta;
	uint64_t total_words = 0;  // words of `packed`, lead and tail pad included
	DevBuf<uint32_t> packed;
	DevBuf<uint2> planes;      // bit planes of `packed` (launch_planes): what the window test of match extension reads
	// '-' in a sequence (SortedMerList.cpp:433-437 throws): pack_kernel raises a flag that is copied to page-locked host
	// memory behind the pack; check_gap() waits for that copy only — the work queued after it keeps the GPU busy meanwhile
	uint32_t* h_gap = nullptr;
	cudaEvent_t gap_ready = nullptr;
	bool gap_pending() const { return gap_ready != nullptr; }
	bool take_gap_flag();  // waits for the flag, releases event and host word; true = a gap was seen
	void check_gap();      // take_gap_flag() and throw MEMS_ERR_GAP if set
	~Batch();
	DevBuf<uint8_t> keys_by_pos;  // compact key of every seed position, in (seq, position) order (extraction output)
	DevBuf<uint8_t> keys;      // union, ascending compact key (u32 or u64); ties in (seq, position) order
	DevBuf<uint32_t> vals;     // union, (seq << pos_bits) | position
	DevBuf<uint32_t> positions;  // per sequence: slice [seed_off, +n_seeds) sorted by key, still tagged
	bool have_positions = false;
	uint32_t pos_mask() const { return pos_bits >= 32 ? 0xffffffffu : ((1u << pos_bits) - 1u); }
	const uint32_t* sorted_positions();  // builds the per-sequence lists on first use
};
std::shared_ptr<Batch> build_batch_from_ascii(std::shared_ptr<Ctx> ctx, int n_seqs, const char* const* seqs,
                                              const uint64_t* lens, uint64_t seed, const std::vector<int>* group_sizes = nullptr);
// layout + H2D + pack only (no keys yet).  tag0 = id of the first sequence; pos_bits/seq_bits > 0 override the
// batch-local choice (sharded runs use the global ones).
std::shared_ptr<Batch> prepare_batch_from_ascii(std::shared_ptr<Ctx> ctx, int n_seqs, const char* const* seqs,
                                                const uint64_t* lens, uint64_t seed, uint32_t tag0, int pos_bits,
                                                int seq_bits, const std::vector<int>* group_sizes = nullptr);
int bits_for(uint64_t max_value);
struct SeqRef {
	const Batch* batch;
	int index;
};
std::shared_ptr<Batch> build_batch_from_packed(std::shared_ptr<Ctx> ctx, const std::vector<SeqRef>& seqs);

// ---- comm.cu ---- (NCCL exchange steps of the sharded path)
struct Comm;
void comm_unique_id(char* id128);
Comm* comm_create(std::shared_ptr<Ctx> ctx, const char* id128, int rank, int world);
void comm_destroy(Comm* c);
int Comm_rank(const Comm* c);
int Comm_world(const Comm* c);
void comm_all_gather_u64(Comm* c, const uint64_t* d_send, uint64_t* d_recv, size_t n);
void comm_all_to_all_v(Comm* c, const void* d_send, const uint64_t* send_counts, void* d_recv, const uint64_t* recv_counts,
                       size_t elem_bytes);
void comm_all_to_all_v_multi(Comm* c, int n_arrays, const void* const* d_send, void* const* d_recv, const size_t* elem_bytes,
                             const uint64_t* send_counts, const uint64_t* recv_counts);  // arrays sharing the counts, one group
// exchange windows (peer-mapped device buffers): reserve is collective and returns false when IPC is unavailable
bool comm_window_reserve(Comm* c, int w, size_t bytes);
void* comm_window_local(Comm* c, int w);
void* comm_window_peer(Comm* c, int w, int p);
void comm_window_barrier(Comm* c);
void comm_window_all_gather(Comm* c, int w, const void* d_send, size_t bytes, size_t offset);
void comm_window_all_to_all(Comm* c, int w, int n_arrays, const void* const* d_send, const size_t* elem_bytes,
                            const size_t* region_off, const uint64_t* counts, bool barrier);
void comm_all_gather_v(Comm* c, const void* d_send, void* d_recv, const uint64_t* byte_counts, const uint64_t* byte_offsets);
void comm_all_gather_v_wait(Comm* c);
void comm_side_synchronize(Comm* c);

// ---- sharding plan (host arithmetic, identical on every rank) ----
// contiguous block of sequences a rank extracts: [first, first + count)
void shard_sequence_range(int n_seqs, int rank, int world, int* first, int* count);
// owner rank of each of the 256 top-digit buckets, balanced by the global bucket histogram, contiguous ranges
void shard_bucket_owners(const uint64_t* hist256, int world, uint8_t* owner256);
// count matrix and slice offsets of the seed-record exchange, from the gathered histograms (see kernels_match.cu)
void shard_exchange_plan(const uint32_t* hist_all, int world, int rank, const uint8_t* owner256, uint64_t* counts,
                         uint64_t* src_elem, uint64_t* dst_elem, uint64_t* max_recv);

// ---- kernels_match.cu ----
// [SeqCount, Length, starts...] records on the host: either a page-locked buffer borrowed from the context
// (filled straight by the D2H copy) or, when the host re-ordered the records, a plain vector.
//
// Delivery is asynchronous where the host has nothing left to do with the records (ORDER_ANY): the call returns once
// its last kernel is queued, the D2H copy runs on the context's copy stream behind it, and `wait()` — called by every
// accessor of the records — blocks until they have arrived.  A caller that issues its next call first gets the copy
// overlapped with that call's kernels (the copy is 5-40 % of a step: 12 MB per config-2 step, 633 MB per rank and
// config-5 step).
struct FlatRecords {
	std::shared_ptr<Ctx> owner;
	int64_t* pinned = nullptr;
	size_t pinned_cap = 0, pinned_n = 0;
	std::vector<int64_t> vec;
	cudaEvent_t ready = nullptr;  // recorded behind the copy into `pinned`
	void* dev_keep = nullptr;     // the copy's device source, released once it has arrived
	void wait() {
		if (!ready) return;
		cudaEventSynchronize(ready);
		owner->event_put(ready);
		ready = nullptr;
		if (dev_keep) owner->free(dev_keep);
		dev_keep = nullptr;
	}
	const int64_t* data() {
		wait();
		return pinned ? pinned : vec.data();
	}
	size_t size() const { return pinned ? pinned_n : vec.size(); }
	void release() {
		wait();
		if (pinned) owner->pinned_put(pinned, pinned_cap);
		pinned = nullptr;
		pinned_n = 0;
	}
	~FlatRecords() { release(); }
};

struct MatchResult {
	FlatRecords flat;
	uint64_t n_matches = 0, n_hits = 0, mem_count = 0, collisions = 0, max_run = 0, n_segments = 0;
	uint32_t seq_count = 0, seed_length = 0;
	double host_replay_ms = 0;
};
// The reference's MemHash::mem_table replayed on the host (ORDER_REFERENCE): buckets of stored, extended matches.
// Persistent instances let several FindMatches calls (different seed patterns) accumulate into one table
// (MemHash::ClearSequences keeps the table, MemHash.cpp:72-74).
struct TableEntry {
	uint32_t seqcount;
	int64_t len, mersize, offset;
	const int64_t* start;  // seqcount values
	std::vector<int64_t> own;
};
struct HashTable {
	uint32_t size = 0;
	std::vector<std::vector<TableEntry*>> buckets;
	std::vector<std::unique_ptr<TableEntry>> stored;
	uint64_t mem_count = 0, collisions = 0;
};
void find_matches_on_batch(Batch& b, int mode, int order, uint32_t table_size, uint64_t seq_mask, MatchResult& out,
                           HashTable* persistent = nullptr, const uint64_t* start_points = nullptr);
// MemHash::AddHashEntry for an already extended match (MemHash::LoadFile); true = inserted, false = collision
bool table_add_entry(HashTable& T, uint32_t seq_count, int64_t length, const int64_t* starts, int64_t mersize);
// the table's content in output order (buckets in order, front to back)
void table_list(const HashTable& T, MatchResult& out);
// every group of the batch as its own problem (MEMS_MODE_MEMHASH / PAIRWISE, ORDER_ANY / CANONICAL): out[g] = group g's matches
void find_matches_many(Batch& b, int mode, int order, std::vector<MatchResult>& out);
void find_matches_sharded(std::shared_ptr<Ctx> ctx, Comm* comm, int n_seqs, const char* const* seqs, const uint64_t* lens,
                          uint64_t seed, int mode, int order, MatchResult& out);

}  // namespace mems
