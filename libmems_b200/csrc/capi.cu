// capi.cu — the extern "C" boundary (include/mems_b200.h).  Exceptions never cross it: every entry point
// returns an int code and leaves the message in the context (or a thread-local for creation failures).
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "mems_b200.h"

using namespace mems;

struct mems_ctx {
	std::shared_ptr<Ctx> c;
};
struct mems_sml {
	std::shared_ptr<Batch> batch;
	int index;
};
struct mems_matches {
	MatchResult r;
};
struct mems_table {
	HashTable t;
};

namespace {
thread_local std::string tl_error;

int fail(Ctx* c, int code, const std::string& msg) {
	if (c) c->last_error = msg;
	tl_error = msg;
	return code;
}

template <class F>
int guarded(Ctx* c, F&& f) {
	try {
		f();
		return MEMS_OK;
	} catch (const Error& e) {
		return fail(c, e.code, e.what());
	} catch (const std::bad_alloc&) {
		return fail(c, MEMS_ERR_CUDA, "host allocation failed");
	} catch (const std::exception& e) {
		return fail(c, MEMS_ERR_INVALID, e.what());
	}
}
}  // namespace

extern "C" {

int mems_ctx_create(int device, void* stream, mems_ctx_t* out) {
	if (!out) return fail(nullptr, MEMS_ERR_INVALID, "null output pointer");
	*out = nullptr;
	return guarded(nullptr, [&] {
		int n_dev = 0;
		cudaError_t e = cudaGetDeviceCount(&n_dev);
		if (e != cudaSuccess || n_dev == 0)
			throw Error(MEMS_ERR_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") +
			                               cudaGetErrorString(e));
		if (device < 0 || device >= n_dev) throw Error(MEMS_ERR_INVALID, "bad device ordinal");
		MEMS_CUDA(cudaSetDevice(device));
		cudaDeviceProp prop;
		MEMS_CUDA(cudaGetDeviceProperties(&prop, device));
		if (prop.major < 10)
			throw Error(MEMS_ERR_CUDA, std::string("device '") + prop.name +
			                               "' is not sm_100 class; this library ships sm_100a code only");
		auto c = std::make_shared<Ctx>();
		c->device = device;
		if (const char* e = getenv("MEMS_TRACE_SLOW")) c->trace_slow_ms = atof(e);
		c->sm_count = prop.multiProcessorCount;
		if (stream) {
			c->stream = (cudaStream_t)stream;
		} else {
			MEMS_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
			c->own_stream = true;
		}
		*out = new mems_ctx{c};
	});
}

void mems_ctx_destroy(mems_ctx_t ctx) { delete ctx; }

const char* mems_last_error(mems_ctx_t ctx) {
	if (ctx && ctx->c) return ctx->c->last_error.c_str();
	return tl_error.c_str();
}

int mems_ctx_synchronize(mems_ctx_t ctx) {
	if (!ctx) return fail(nullptr, MEMS_ERR_INVALID, "null context");
	return guarded(ctx->c.get(), [&] { MEMS_CUDA(cudaStreamSynchronize(ctx->c->stream)); });
}

int mems_ctx_trim(mems_ctx_t ctx, uint64_t* reserved_bytes) {
	if (!ctx) return fail(nullptr, MEMS_ERR_INVALID, "null context");
	return guarded(ctx->c.get(), [&] {
		Ctx* c = ctx->c.get();
		MEMS_CUDA(cudaSetDevice(c->device));
		MEMS_CUDA(cudaStreamSynchronize(c->stream));
		if (c->copy_stream) MEMS_CUDA(cudaStreamSynchronize(c->copy_stream));
		std::lock_guard<std::mutex> lock(c->arena_mutex);
		for (auto& sl : c->arena.drop_idle_slabs()) cudaFree(sl.first);
		if (reserved_bytes) *reserved_bytes = c->arena.reserved;
	});
}

int mems_host_alloc(void** ptr, uint64_t bytes) {
	return guarded(nullptr, [&] { MEMS_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault)); });
}
void mems_host_free(void* ptr) {
	if (ptr) cudaFreeHost(ptr);
}

// ------------------------------------------------------------------------------------------------ SML
int mems_sml_create_batch(mems_ctx_t ctx, int n_seqs, const char* const* seqs, const uint64_t* lens, uint64_t seed,
                          mems_sml_t* out) {
	if (!ctx) return fail(nullptr, MEMS_ERR_INVALID, "null context");
	if (!seqs || !lens || !out || n_seqs < 1) return fail(ctx->c.get(), MEMS_ERR_INVALID, "bad arguments");
	return guarded(ctx->c.get(), [&] {
		for (int g = 0; g < n_seqs; ++g)
			if (lens[g] && !seqs[g]) throw Error(MEMS_ERR_INVALID, "null sequence pointer");
		auto b = build_batch_from_ascii(ctx->c, n_seqs, seqs, lens, seed);
		for (int g = 0; g < n_seqs; ++g) out[g] = new mems_sml{b, g};
	});
}

int mems_sml_create(mems_ctx_t ctx, const char* seq, uint64_t n, uint64_t seed, mems_sml_t* out) {
	const char* seqs[1] = {seq};
	uint64_t lens[1] = {n};
	return mems_sml_create_batch(ctx, 1, seqs, lens, seed, out);
}

void mems_sml_destroy(mems_sml_t sml) {
	if (!sml) return;
	if (sml->batch) cudaSetDevice(sml->batch->ctx->device);
	delete sml;
}

int mems_sml_clone(mems_sml_t sml, mems_sml_t* out) {
	if (!sml || !out) return fail(nullptr, MEMS_ERR_INVALID, "null argument");
	return guarded(nullptr, [&] { *out = new mems_sml{sml->batch, sml->index}; });
}

int mems_sml_info(mems_sml_t sml, mems_sml_info_t* out) {
	if (!sml || !out) return fail(nullptr, MEMS_ERR_INVALID, "null argument");
	const Batch& b = *sml->batch;
	const SeqMeta& m = b.meta[sml->index];
	out->length = m.n_bases;
	out->sml_length = m.n_seeds;
	out->seed = b.sd.seed;
	out->seed_length = (uint32_t)b.sd.L;
	out->seed_weight = (uint32_t)b.sd.w;
	out->seed_mask = ~0ull << (64 - 2 * b.sd.w);  // SortedMerList.cpp:819-821
	out->mer_mask = ~0ull << (64 - 2 * b.sd.L);
	return MEMS_OK;
}

int mems_sml_read(mems_sml_t sml, uint64_t offset, uint64_t count, uint32_t* positions_out, uint64_t* mers_out,
                  uint64_t* n_read) {
	if (!sml) return fail(nullptr, MEMS_ERR_INVALID, "null sml");
	Batch& b = *sml->batch;
	Ctx* c = b.ctx.get();
	return guarded(c, [&] {
		MEMS_CUDA(cudaSetDevice(c->device));
		const SeqMeta& m = b.meta[sml->index];
		uint64_t cnt = offset >= m.n_seeds ? 0 : std::min<uint64_t>(count, m.n_seeds - offset);
		if (n_read) *n_read = cnt;
		if (cnt == 0) return;
		const uint32_t* pos = b.sorted_positions() + m.seed_off + offset;
		if (mers_out) {
			DevBuf<uint64_t> mers(c, cnt);
			launch_sml_read(c, b.packed.p + m.word_off, b.sd, pos, b.pos_mask(), cnt, mers.p);
			MEMS_CUDA(cudaMemcpyAsync(mers_out, mers.p, cnt * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
		}
		if (positions_out)
			MEMS_CUDA(cudaMemcpyAsync(positions_out, pos, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
		MEMS_CUDA(cudaStreamSynchronize(c->stream));
		if (positions_out) {
			const uint32_t mask = b.pos_mask();
			for (uint64_t i = 0; i < cnt; ++i) positions_out[i] &= mask;
		}
	});
}

int mems_sml_seed_mers(mems_sml_t sml, const uint64_t* positions, uint64_t n, uint64_t* fwd_out, uint64_t* dna_out) {
	if (!sml) return fail(nullptr, MEMS_ERR_INVALID, "null sml");
	Batch& b = *sml->batch;
	Ctx* c = b.ctx.get();
	return guarded(c, [&] {
		if (n == 0) return;
		MEMS_CUDA(cudaSetDevice(c->device));
		const SeqMeta& m = b.meta[sml->index];
		DevBuf<uint64_t> d_pos(c, n), d_fwd(c, n), d_dna(c, n);
		MEMS_CUDA(cudaMemcpyAsync(d_pos.p, positions, n * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
		launch_seed_mers(c, b.packed.p + m.word_off, m.n_seeds, b.sd, d_pos.p, n, d_fwd.p, d_dna.p);
		if (fwd_out) MEMS_CUDA(cudaMemcpyAsync(fwd_out, d_fwd.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
		if (dna_out) MEMS_CUDA(cudaMemcpyAsync(dna_out, d_dna.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
		MEMS_CUDA(cudaStreamSynchronize(c->stream));
	});
}

int mems_sml_find_mer(mems_sml_t sml, uint64_t query_mer, int* found, uint64_t* index) {
	if (!sml || !found || !index) return fail(nullptr, MEMS_ERR_INVALID, "null argument");
	Batch& b = *sml->batch;
	Ctx* c = b.ctx.get();
	return guarded(c, [&] {
		MEMS_CUDA(cudaSetDevice(c->device));
		const SeqMeta& m = b.meta[sml->index];
		*found = 0;
		*index = 0;
		if (m.n_seeds == 0) return;  // FindMer returns false on sequences shorter than the seed
		const uint32_t* pos = b.sorted_positions() + m.seed_off;
		DevBuf<uint64_t> res(c, 2);
		launch_find_mer(c, b.packed.p + m.word_off, b.sd, pos, b.pos_mask(), m.n_seeds, query_mer, res.p);
		uint64_t h[2];
		MEMS_CUDA(cudaMemcpyAsync(h, res.p, sizeof h, cudaMemcpyDeviceToHost, c->stream));
		MEMS_CUDA(cudaStreamSynchronize(c->stream));
		*index = h[0];
		*found = (int)h[1];
	});
}

int mems_sml_packed(mems_sml_t sml, uint32_t* words_out, uint64_t* n_words) {
	if (!sml) return fail(nullptr, MEMS_ERR_INVALID, "null sml");
	Batch& b = *sml->batch;
	Ctx* c = b.ctx.get();
	return guarded(c, [&] {
		MEMS_CUDA(cudaSetDevice(c->device));
		const SeqMeta& m = b.meta[sml->index];
		uint64_t nw = ((uint64_t)m.n_bases * 2 + 31) / 32 + 2;  // SortedMerList.cpp:306-311
		if (n_words) *n_words = nw;
		if (!words_out) return;
		MEMS_CUDA(cudaMemcpyAsync(words_out, b.packed.p + m.word_off, nw * sizeof(uint32_t), cudaMemcpyDeviceToHost,
		                          c->stream));
		MEMS_CUDA(cudaStreamSynchronize(c->stream));
	});
}

int mems_sml_seed_occurrence(mems_sml_t sml, float* out) {
	if (!sml || !out) return fail(nullptr, MEMS_ERR_INVALID, "null argument");
	Batch& b = *sml->batch;
	Ctx* c = b.ctx.get();
	return guarded(c, [&] {
		MEMS_CUDA(cudaSetDevice(c->device));
		const SeqMeta& m = b.meta[sml->index];
		if (m.n_bases == 0) return;
		const uint32_t* pos = b.sorted_positions() + m.seed_off;
		const size_t K = b.key64 ? 8 : 4;
		DevBuf<float> d_out(c, m.n_bases);
		launch_seed_occurrence(c, pos, b.pos_mask(), b.keys_by_pos.p + m.seed_off * K, b.key64, m.n_seeds, m.n_bases, b.sd.L,
		                       d_out.p);
		MEMS_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)m.n_bases * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
		MEMS_CUDA(cudaStreamSynchronize(c->stream));
	});
}

// ------------------------------------------------------------------------------------------------ matches
int mems_find_matches(mems_ctx_t ctx, int n_smls, const mems_sml_t* smls, const mems_match_params_t* params,
                      mems_matches_t* out) {
	if (!ctx) return fail(nullptr, MEMS_ERR_INVALID, "null context");
	Ctx* c = ctx->c.get();
	if (!smls || !out || n_smls < 1) return fail(c, MEMS_ERR_INVALID, "bad arguments");
	return guarded(c, [&] {
		MEMS_CUDA(cudaSetDevice(c->device));
		mems_match_params_t p;
		memset(&p, 0, sizeof p);
		if (params) p = *params;
		if (p.mode < 0 || p.mode > MEMS_MODE_PAIRWISE) throw Error(MEMS_ERR_INVALID, "bad mode");
		if (p.mode == MEMS_MODE_REPEAT && n_smls != 1)
			throw Error(MEMS_ERR_INVALID, "RepeatHash works on exactly one sequence (RepeatHash.cpp:26-32)");
		if (n_smls > MEMS_MAX_SEQS) throw Error(MEMS_ERR_UNSUPPORTED, "more than MEMS_MAX_SEQS sequences");
		for (int g = 0; g < n_smls; ++g) {
			if (!smls[g]) throw Error(MEMS_ERR_INVALID, "Null SortedMerList pointer");  // MatchFinder.cpp:59-62
			if (smls[g]->batch->sd.seed != smls[0]->batch->sd.seed)
				throw Error(MEMS_ERR_SEED_MISMATCH, "Different seed patterns.");  // MatchFinder.cpp:190-199
		}
		// fast path: the handles are exactly one batch in order, on this context -> its sorted union is reused
		std::shared_ptr<Batch> b = smls[0]->batch;
		bool same = b->n_seqs == n_smls && b->ctx.get() == c;
		for (int g = 0; same && g < n_smls; ++g) same = smls[g]->batch == b && smls[g]->index == g;
		if (!same) {
			std::vector<SeqRef> refs;
			for (int g = 0; g < n_smls; ++g) refs.push_back({smls[g]->batch.get(), smls[g]->index});
			b = build_batch_from_packed(ctx->c, refs);
		}
		auto* m = new mems_matches();
		try {
			if (p.table && p.table->t.size == 0) p.table->t.size = p.table_size ? p.table_size : 40000u;
			find_matches_on_batch(*b, p.mode, (p.mode == MEMS_MODE_REPEAT || p.table) ? MEMS_ORDER_REFERENCE : p.order,
			                      p.table_size ? p.table_size : 40000u, p.seq_mask, m->r, p.table ? &p.table->t : nullptr,
			                      p.start_points);
		} catch (...) {
			delete m;
			throw;
		}
		*out = m;
	});
}

int mems_find_matches_many(mems_ctx_t ctx, int n_problems, const int* n_seqs, const char* const* seqs, const uint64_t* lens,
                           uint64_t seed, const mems_match_params_t* params, mems_matches_t* out) {
	if (!ctx) return fail(nullptr, MEMS_ERR_INVALID, "null context");
	Ctx* c = ctx->c.get();
	if (!n_seqs || !seqs || !lens || !out || n_problems < 1) return fail(c, MEMS_ERR_INVALID, "bad arguments");
	return guarded(c, [&] {
		MEMS_CUDA(cudaSetDevice(c->device));
		mems_match_params_t p;
		memset(&p, 0, sizeof p);
		if (params) p = *params;
		if (p.table || p.seq_mask || p.start_points) throw Error(MEMS_ERR_INVALID, "table, seq_mask and start_points do not apply to mems_find_matches_many");
		std::vector<int> groups(n_seqs, n_seqs + n_problems);
		int total = 0;
		for (int g : groups) {
			if (g < 1 || g > MEMS_MAX_SEQS) throw Error(MEMS_ERR_UNSUPPORTED, "1..MEMS_MAX_SEQS sequences per problem");
			total += g;
		}
		for (int i = 0; i < total; ++i)
			if (lens[i] && !seqs[i]) throw Error(MEMS_ERR_INVALID, "null sequence pointer");
		auto b = build_batch_from_ascii(ctx->c, total, seqs, lens, seed, &groups);
		std::vector<MatchResult> results;
		find_matches_many(*b, p.mode, p.order, results);
		for (int g = 0; g < n_problems; ++g) {
			auto* m = new mems_matches();
			m->r = std::move(results[g]);
			out[g] = m;
		}
	});
}

int mems_table_create(uint32_t table_size, mems_table_t* out) {
	if (!out) return fail(nullptr, MEMS_ERR_INVALID, "null argument");
	return guarded(nullptr, [&] {
		auto* t = new mems_table();
		t->t.size = table_size ? table_size : 40000u;
		*out = t;
	});
}

void mems_table_clear(mems_table_t t) {
	if (!t) return;
	const uint32_t size = t->t.size;
	t->t = HashTable();
	t->t.size = size;
}

void mems_table_destroy(mems_table_t t) { delete t; }

int mems_table_add(mems_table_t t, uint32_t seq_count, uint64_t length, const int64_t* starts, uint32_t mersize, int* inserted) {
	if (!t || !starts || seq_count == 0) return fail(nullptr, MEMS_ERR_INVALID, "bad arguments");
	return guarded(nullptr, [&] {
		const bool ins = table_add_entry(t->t, seq_count, (int64_t)length, starts, (int64_t)mersize);
		if (inserted) *inserted = ins ? 1 : 0;
	});
}

int mems_table_matches(mems_table_t t, mems_matches_t* out) {
	if (!t || !out) return fail(nullptr, MEMS_ERR_INVALID, "null argument");
	return guarded(nullptr, [&] {
		auto* m = new mems_matches();
		table_list(t->t, m->r);
		*out = m;
	});
}

int mems_matches_info(mems_matches_t m, mems_matches_info_t* out) {
	if (!m || !out) return fail(nullptr, MEMS_ERR_INVALID, "null argument");
	out->n_matches = m->r.n_matches;
	out->n_flat = m->r.flat.size();
	out->n_hits = m->r.n_hits;
	out->mem_count = m->r.mem_count;
	out->collisions = m->r.collisions;
	out->max_run = m->r.max_run;
	out->n_segments = m->r.n_segments;
	out->seq_count = m->r.seq_count;
	out->seed_length = m->r.seed_length;
	out->host_replay_ms = m->r.host_replay_ms;
	return MEMS_OK;
}

int mems_matches_copy(mems_matches_t m, int64_t* flat_out) {
	if (!m || (!flat_out && m->r.flat.size())) return fail(nullptr, MEMS_ERR_INVALID, "null argument");
	if (m->r.flat.size()) memcpy(flat_out, m->r.flat.data(), m->r.flat.size() * sizeof(int64_t));
	return MEMS_OK;
}

const int64_t* mems_matches_data(mems_matches_t m) { return m ? m->r.flat.data() : nullptr; }

int mems_matches_wait(mems_matches_t m) {
	if (!m) return fail(nullptr, MEMS_ERR_INVALID, "null argument");
	m->r.flat.wait();
	return MEMS_OK;
}

void mems_matches_destroy(mems_matches_t m) { delete m; }

// ------------------------------------------------------------------------------------------------ sharded
struct mems_comm {
	Comm* c;
};

int mems_comm_unique_id(char* id_out) {
	if (!id_out) return fail(nullptr, MEMS_ERR_INVALID, "null argument");
	return guarded(nullptr, [&] { comm_unique_id(id_out); });
}

int mems_comm_create(mems_ctx_t ctx, const char* id, int rank, int world, mems_comm_t* out) {
	if (!ctx || !id || !out) return fail(nullptr, MEMS_ERR_INVALID, "null argument");
	return guarded(ctx->c.get(), [&] { *out = new mems_comm{comm_create(ctx->c, id, rank, world)}; });
}

void mems_comm_destroy(mems_comm_t comm) {
	if (!comm) return;
	comm_destroy(comm->c);
	delete comm;
}

int mems_shard_sequence_range(int n_seqs, int rank, int world, int* first, int* count) {
	if (!first || !count || world < 1 || rank < 0 || rank >= world || n_seqs < 0) return fail(nullptr, MEMS_ERR_INVALID, "bad arguments");
	shard_sequence_range(n_seqs, rank, world, first, count);
	return MEMS_OK;
}

int mems_shard_bucket_owners(const uint64_t* hist256, int world, uint8_t* owner256) {
	if (!hist256 || !owner256 || world < 1 || world > 256) return fail(nullptr, MEMS_ERR_INVALID, "bad arguments");
	shard_bucket_owners(hist256, world, owner256);
	return MEMS_OK;
}

int mems_shard_exchange_plan(const uint32_t* hist_all, int world, int rank, const uint8_t* owner256, uint64_t* counts,
                             uint64_t* src_elem, uint64_t* dst_elem, uint64_t* max_recv) {
	if (!hist_all || !owner256 || !counts || !src_elem || !dst_elem || !max_recv || world < 1 || world > 256 || rank < 0 ||
	    rank >= world)
		return fail(nullptr, MEMS_ERR_INVALID, "bad arguments");
	for (int b = 0; b < 256; ++b)
		if (owner256[b] >= world) return fail(nullptr, MEMS_ERR_INVALID, "owner out of range");
	shard_exchange_plan(hist_all, world, rank, owner256, counts, src_elem, dst_elem, max_recv);
	return MEMS_OK;
}

int mems_find_matches_sharded(mems_ctx_t ctx, mems_comm_t comm, int n_seqs, const char* const* seqs, const uint64_t* lens,
                              uint64_t seed, const mems_match_params_t* params, mems_matches_t* out) {
	if (!ctx) return fail(nullptr, MEMS_ERR_INVALID, "null context");
	Ctx* c = ctx->c.get();
	if (!comm || !seqs || !lens || !out) return fail(c, MEMS_ERR_INVALID, "bad arguments");
	return guarded(c, [&] {
		mems_match_params_t p;
		memset(&p, 0, sizeof p);
		if (params) p = *params;
		auto* m = new mems_matches();
		try {
			find_matches_sharded(ctx->c, comm->c, n_seqs, seqs, lens, seed, p.mode, p.order, m->r);
		} catch (...) {
			delete m;
			throw;
		}
		*out = m;
	});
}

int mems_selftest_arena(uint64_t seed, int rounds) { return arena_selftest(seed, rounds); }

// ------------------------------------------------------------------------------------------------ measurement
int mems_profile_enable(mems_ctx_t ctx, int on) {
	if (!ctx) return fail(nullptr, MEMS_ERR_INVALID, "null context");
	return guarded(ctx->c.get(), [&] {
		MEMS_CUDA(cudaSetDevice(ctx->c->device));
		if (ctx->c->profiling && !on) ctx->c->prof_collect();
		ctx->c->profiling = on != 0;
	});
}

int mems_profile_reset(mems_ctx_t ctx) {
	if (!ctx) return fail(nullptr, MEMS_ERR_INVALID, "null context");
	return guarded(ctx->c.get(), [&] {
		MEMS_CUDA(cudaSetDevice(ctx->c->device));
		ctx->c->prof_collect();
		ctx->c->prof.clear();
		ctx->c->launch_count = 0;
	});
}

int mems_profile_get(mems_ctx_t ctx, mems_profile_entry_t* entries, int cap, int* n) {
	if (!ctx || !n) return fail(nullptr, MEMS_ERR_INVALID, "null argument");
	return guarded(ctx->c.get(), [&] {
		MEMS_CUDA(cudaSetDevice(ctx->c->device));
		ctx->c->prof_collect();
		int i = 0;
		for (auto& kv : ctx->c->prof) {
			if (entries && i < cap) {
				memset(&entries[i], 0, sizeof entries[i]);
				strncpy(entries[i].name, kv.first.c_str(), sizeof(entries[i].name) - 1);
				entries[i].launches = kv.second.launches;
				entries[i].ms = kv.second.ms;
				entries[i].bytes = kv.second.bytes;
			}
			++i;
		}
		*n = i;
	});
}

uint64_t mems_launch_count(mems_ctx_t ctx) { return ctx ? ctx->c->launch_count : 0; }

int mems_test_hooks(mems_ctx_t ctx, int hash_bits, int walk_budget) {
	if (!ctx || hash_bits < 0 || hash_bits > 63 || walk_budget < 0) return fail(nullptr, MEMS_ERR_INVALID, "bad arguments");
	ctx->c->test_hash_bits = hash_bits;
	ctx->c->test_walk_budget = walk_budget;
	return MEMS_OK;
}

}  // extern "C"
