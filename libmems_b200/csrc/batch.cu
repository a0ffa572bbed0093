// batch.cu — SML construction for a batch of sequences:
//   H2D of the ASCII bases -> pack -> extract (+ digit histograms) -> LSD radix sort of the union.
// MemorySML::Create (MemorySML.cpp:45-60) builds one list per sequence; here all sequences of a create
// call are sorted together as ONE array whose ties are in (sequence, position) order.  That single sort
// serves both consumers:
//   * match finding reads the union directly — equal-seed runs across sequences are contiguous, which is
//     what MatchFinder::SearchRange's k-way merge (MatchFinder.cpp:172-340) produces on the CPU;
//   * the per-sequence lists the SortedMerList interface exposes are one more stable counting pass on the
//     sequence tag (done on first use).
#include <algorithm>

#include "common.cuh"
#include "mems_b200.h"

namespace mems {

int bits_for(uint64_t max_value) {  // bits needed to store values 0..max_value
	int b = 1;
	while (b < 64 && (max_value >> b)) ++b;
	return b;
}

static void layout_batch(Batch& b, const std::vector<uint64_t>& lens, uint32_t tag0 = 0, int force_pos_bits = 0,
                         int force_seq_bits = 0, const std::vector<int>* group_sizes = nullptr) {
	// the per-sequence lists are one counting pass on the sequence tag (Batch::sorted_positions): 8 digit bits
	if (lens.size() > 256)
		throw Error(MEMS_ERR_UNSUPPORTED, "more than 256 sequences in one create call; build them in several batches");
	b.n_seqs = (int)lens.size();
	b.meta.resize(b.n_seqs);
	uint64_t byte_off = 0, word_off = kLeadWords, seed_off = 0;
	uint32_t max_seeds = 0;
	for (int g = 0; g < b.n_seqs; ++g) {
		if (lens[g] > 0xffffffffull)
			throw Error(MEMS_ERR_UNSUPPORTED, "sequence longer than 2^32-1 bases (positions are uint32, SortedMerList.h:40)");
		SeqMeta& m = b.meta[g];
		m.n_bases = (uint32_t)lens[g];
		m.n_seeds = lens[g] >= (uint64_t)b.sd.L ? (uint32_t)(lens[g] - b.sd.L + 1) : 0u;  // SMLLength, linear
		m.byte_off = byte_off;
		m.word_off = word_off;
		m.seed_off = seed_off;
		m.tag = tag0 + (uint32_t)g;
		m.group = 0;
		m.group_first = tag0;
		m.group_count = (uint32_t)lens.size();
		byte_off += (lens[g] + 15) / 16 * 16;
		word_off += seq_packed_words(lens[g]);  // keeps every sequence 16-byte aligned
		seed_off += m.n_seeds;
		max_seeds = std::max(max_seeds, m.n_seeds);
	}
	b.n_total = seed_off;
	b.total_words = word_off + kTailWords;
	b.pos_bits = bits_for(max_seeds ? max_seeds - 1 : 0);
	b.seq_bits = b.n_seqs > 1 ? bits_for((uint64_t)b.n_seqs - 1) : 0;
	if (force_pos_bits) b.pos_bits = force_pos_bits;
	if (force_seq_bits) b.seq_bits = force_seq_bits;
	if (b.pos_bits + b.seq_bits > 32)
		throw Error(MEMS_ERR_UNSUPPORTED,
		            "sequence count x longest sequence does not fit the 32-bit (sequence, position) tag; "
		            "shard the sequences across devices");
	if (b.n_total > radix_max_items())
		throw Error(MEMS_ERR_UNSUPPORTED, "more than 2^30-1 seed positions in one device batch; shard across devices");
	b.n_groups = 1;
	b.max_group = b.n_seqs;
	b.group_bits = 0;
	if (group_sizes) {  // several independent problems in one batch
		size_t at = 0;
		b.max_group = 0;
		for (size_t k = 0; k < group_sizes->size(); ++k) {
			const int cnt = (*group_sizes)[k];
			if (cnt < 1 || at + (size_t)cnt > lens.size()) throw Error(MEMS_ERR_INVALID, "bad problem sizes");
			for (int j = 0; j < cnt; ++j) {
				SeqMeta& m = b.meta[at + j];
				m.group = (uint32_t)k;
				m.group_first = (uint32_t)at;
				m.group_count = (uint32_t)cnt;
			}
			at += (size_t)cnt;
			b.max_group = std::max(b.max_group, cnt);
		}
		if (at != lens.size()) throw Error(MEMS_ERR_INVALID, "problem sizes do not add up to the sequence count");
		b.n_groups = (int)group_sizes->size();
		b.group_bits = b.n_groups > 1 ? bits_for((uint64_t)b.n_groups - 1) : 0;
		if (b.sd.key_bits + b.group_bits > 64) throw Error(MEMS_ERR_UNSUPPORTED, "seed weight x problem count exceeds 64 key bits");
	}
	b.key64 = b.sd.key_bits + b.group_bits > 32;
}

// packed sequences are in place: extract keys, sort the union
static void extract_and_sort(Batch& b) {
	Ctx* c = b.ctx.get();
	const size_t key_bytes = b.key64 ? 8 : 4;
	const uint64_t n = b.n_total;
	SortPlan plan = make_sort_plan(b.sort_bits());
	// extraction output (position order) is the input of the first pass and stays resident for SeedOccurrenceList / start points
	DevBuf<uint8_t> keys_pos(c, n * key_bytes), keys_a(c, n * key_bytes), keys_b(c, n * key_bytes);
	DevBuf<uint32_t> vals_a(c, n), vals_b(c, n);
	DevBuf<uint32_t> hist(c, (size_t)plan.n_passes * 256);
	MEMS_CUDA(cudaMemsetAsync(hist.p, 0, (size_t)plan.n_passes * 256 * sizeof(uint32_t), c->stream));
	launch_extract(c, b.packed.p, b.d_meta.p, b.meta.data(), b.n_seqs, b.sd, b.pos_bits, b.key64, keys_pos.p, vals_a.p,
	               hist.p, plan.n_passes, plan.shift, plan.bits);
	void* kp[2] = {keys_a.p, keys_b.p};
	uint32_t* vp[2] = {vals_a.p, vals_b.p};
	int r = radix_sort_pairs(c, b.key64, kp, vp, n, plan, hist.p, "radix_pass", keys_pos.p);
	b.keys_by_pos = std::move(keys_pos);
	if (r == 0) {
		b.keys = std::move(keys_a);
		b.vals = std::move(vals_a);
	} else {
		b.keys = std::move(keys_b);
		b.vals = std::move(vals_b);
	}
}

std::shared_ptr<Batch> prepare_batch_from_ascii(std::shared_ptr<Ctx> ctx, int n_seqs, const char* const* seqs,
                                                const uint64_t* lens, uint64_t seed, uint32_t tag0, int pos_bits,
                                                int seq_bits, const std::vector<int>* group_sizes) {
	auto b = std::make_shared<Batch>();
	b->ctx = ctx;
	b->sd = make_seed_desc(seed);
	Ctx* c = ctx.get();
	MEMS_CUDA(cudaSetDevice(c->device));
	layout_batch(*b, std::vector<uint64_t>(lens, lens + n_seqs), tag0, pos_bits, seq_bits, group_sizes);
	if (n_seqs == 0) return b;
	const SeqMeta& last = b->meta.back();
	const uint64_t total_bytes = last.byte_off + ((uint64_t)last.n_bases + 15) / 16 * 16;

	b->d_meta = DevBuf<SeqMeta>(c, b->n_seqs);
	b->packed = DevBuf<uint32_t>(c, b->total_words);
	{
		CopyScope zs(c, "copy_zero_packed", (double)b->total_words * 4);
		MEMS_CUDA(cudaMemsetAsync(b->packed.p, 0, b->total_words * sizeof(uint32_t), c->stream));  // pads and alignment gaps
	}
	// Sequences in host memory (pageable or page-locked) are staged in one device buffer; a sequence that already lives
	// in this device's memory, 16-byte aligned, is packed where it lies.
	std::vector<char> in_place(n_seqs, 0);
	uint64_t staged_bytes = 0;
	for (int g = 0; g < n_seqs; ++g) {
		if (!lens[g]) continue;
		cudaPointerAttributes attr;
		if (cudaPointerGetAttributes(&attr, seqs[g]) == cudaSuccess && attr.type == cudaMemoryTypeDevice && attr.device == c->device &&
		    (reinterpret_cast<uintptr_t>(seqs[g]) & 15u) == 0)
			in_place[g] = 1;
		else
			staged_bytes += (lens[g] + 15) / 16 * 16;
		cudaGetLastError();  // (an unregistered host pointer makes older runtimes report an error here)
	}
	DevBuf<uint8_t> ascii(c, (staged_bytes ? total_bytes : 0) + 16);
	DevBuf<uint32_t> gap_flag(c, 1);
	MEMS_CUDA(cudaMemsetAsync(gap_flag.p, 0, sizeof(uint32_t), c->stream));
	{
		CopyScope cs(c, "copy_in_sequences", (double)staged_bytes);
		for (int g = 0; g < n_seqs; ++g) {
			if (!lens[g]) continue;
			if (in_place[g])
				b->meta[g].byte_off = reinterpret_cast<uint64_t>(seqs[g]) - reinterpret_cast<uint64_t>(ascii.p);  // see pack_kernel
			else
				MEMS_CUDA(cudaMemcpyAsync(ascii.p + b->meta[g].byte_off, seqs[g], lens[g], cudaMemcpyDefault, c->stream));
		}
	}
	MEMS_CUDA(cudaMemcpyAsync(b->d_meta.p, b->meta.data(), sizeof(SeqMeta) * b->n_seqs, cudaMemcpyHostToDevice, c->stream));
	launch_pack(c, ascii.p, b->packed.p, b->d_meta.p, b->meta.data(), n_seqs, gap_flag.p);
	b->planes = DevBuf<uint2>(c, b->total_words / 2);
	launch_planes(c, b->packed.p, b->planes.p, b->total_words);
	// the flag travels to the host behind the pack; whoever queued the rest of the build calls check_gap()
	b->h_gap = c->host_words_get();
	b->gap_ready = c->get_event();
	c->fetch_async(b->h_gap, gap_flag.p, 1);
	MEMS_CUDA(cudaEventRecord(b->gap_ready, c->stream));
	return b;
}

bool Batch::take_gap_flag() {
	if (!gap_ready) return false;
	Ctx* c = ctx.get();
	const cudaError_t e = cudaEventSynchronize(gap_ready);
	const bool gap = e == cudaSuccess && *h_gap != 0u;
	c->event_put(gap_ready);
	c->host_words_put(h_gap);
	gap_ready = nullptr;
	h_gap = nullptr;
	MEMS_CUDA(e);
	return gap;
}

void Batch::check_gap() {
	if (take_gap_flag())
		throw Error(MEMS_ERR_GAP, "Gap in genome sequence: input sequences must be unaligned and ungapped "
		                          "(SortedMerList.cpp:433-437)");
}

Batch::~Batch() {
	if (gap_ready) {
		cudaEventSynchronize(gap_ready);  // the copy into h_gap must have landed before the word is reused
		ctx->event_put(gap_ready);
		ctx->host_words_put(h_gap);
	}
}

std::shared_ptr<Batch> build_batch_from_ascii(std::shared_ptr<Ctx> ctx, int n_seqs, const char* const* seqs,
                                              const uint64_t* lens, uint64_t seed, const std::vector<int>* group_sizes) {
	if (n_seqs < 1) throw Error(MEMS_ERR_INVALID, "need at least one sequence");
	auto b = prepare_batch_from_ascii(ctx, n_seqs, seqs, lens, seed, 0, 0, 0, group_sizes);
	extract_and_sort(*b);  // queued behind the pack; the GPU works on it while the host looks at the gap flag
	b->check_gap();
	return b;
}

std::shared_ptr<Batch> build_batch_from_packed(std::shared_ptr<Ctx> ctx, const std::vector<SeqRef>& seqs) {
	if (seqs.empty()) throw Error(MEMS_ERR_INVALID, "need at least one sequence");
	auto b = std::make_shared<Batch>();
	b->ctx = ctx;
	b->sd = seqs[0].batch->sd;
	Ctx* c = ctx.get();
	MEMS_CUDA(cudaSetDevice(c->device));
	std::vector<uint64_t> lens;
	for (const SeqRef& s : seqs) lens.push_back(s.batch->meta[s.index].n_bases);
	layout_batch(*b, lens);
	b->d_meta = DevBuf<SeqMeta>(c, b->n_seqs);
	MEMS_CUDA(cudaMemcpyAsync(b->d_meta.p, b->meta.data(), sizeof(SeqMeta) * b->n_seqs, cudaMemcpyHostToDevice, c->stream));
	b->packed = DevBuf<uint32_t>(c, b->total_words);
	MEMS_CUDA(cudaMemsetAsync(b->packed.p, 0, b->total_words * sizeof(uint32_t), c->stream));
	for (size_t g = 0; g < seqs.size(); ++g) {
		const SeqMeta& src = seqs[g].batch->meta[seqs[g].index];
		// the source lives on another context's stream: make sure its producer has finished
		if (seqs[g].batch->ctx.get() != c) MEMS_CUDA(cudaStreamSynchronize(seqs[g].batch->ctx->stream));
		uint64_t words = ((uint64_t)src.n_bases + 15) / 16 + 2;
		MEMS_CUDA(cudaMemcpyAsync(b->packed.p + b->meta[g].word_off, seqs[g].batch->packed.p + src.word_off,
		                          words * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
	}
	b->planes = DevBuf<uint2>(c, b->total_words / 2);
	launch_planes(c, b->packed.p, b->planes.p, b->total_words);
	extract_and_sort(*b);
	return b;
}

// Per-sequence sorted position lists: one stable counting pass of the union on the sequence tag.
const uint32_t* Batch::sorted_positions() {
	if (have_positions) return n_seqs == 1 ? vals.p : positions.p;
	Ctx* c = ctx.get();
	MEMS_CUDA(cudaSetDevice(c->device));
	if (n_seqs > 1 && n_total > 0) {
		SortPlan plan;
		plan.n_passes = 1;
		plan.shift[0] = pos_bits;
		plan.bits[0] = seq_bits;
		std::vector<uint32_t> h(256, 0);
		for (int g = 0; g < n_seqs; ++g) h[g] = meta[g].n_seeds;
		DevBuf<uint32_t> hist(c, 256);
		MEMS_CUDA(cudaMemcpyAsync(hist.p, h.data(), 256 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
		DevBuf<uint32_t> out_keys(c, n_total), out_vals(c, n_total);
		void* kp[2] = {vals.p, out_keys.p};
		uint32_t* vp[2] = {vals.p, out_vals.p};
		radix_sort_pairs(c, false, kp, vp, n_total, plan, hist.p, "split_by_seq");
		MEMS_CUDA(cudaStreamSynchronize(c->stream));  // h (host) is read by the async copy above
		positions = std::move(out_keys);
	}
	have_positions = true;
	return n_seqs == 1 ? vals.p : positions.p;
}

}  // namespace mems
