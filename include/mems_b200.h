/* mems_b200.h — C-ABI of the B200-native seed-and-extend anchoring path of libMems.
 *
 * This is the drop-in boundary: plain C, opaque handles, int error codes, no C++ or torch types.
 * Every entry point names the libMems interface it replaces (file:line under libMems/ of
 * koadman/libMems 1.6.1).  The C++ façade in libmems_b200/host/ (same class and method names as
 * the reference) and the ctypes binding in libmems_b200/__init__.py are the only callers.
 *
 * Threading: a context owns one CUDA stream and its scratch memory.  Contexts are independent, so
 * one context per host thread reproduces the reference's "one MemHash per thread" rule
 * (TLS<MemHash> gap_mh, Aligner.h:198).  Calls on ONE context must not overlap.
 *
 * There is no CPU fallback: every compute entry point fails with MEMS_ERR_CUDA when no sm_100
 * device is usable.
 */
#ifndef MEMS_B200_H
#define MEMS_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MEMS_OK 0
#define MEMS_ERR_INVALID 1      /* bad argument (InvalidData / SMLCreateError in the reference) */
#define MEMS_ERR_GAP 2          /* '-' in a sequence (SortedMerList.cpp:433-437 throws) */
#define MEMS_ERR_CUDA 3         /* CUDA runtime error; see mems_last_error */
#define MEMS_ERR_UNSUPPORTED 4  /* outside this implementation's limits; see mems_last_error */
#define MEMS_ERR_SEED_MISMATCH 5 /* SMLs with different seed patterns (MatchFinder.cpp:190-199) */
#define MEMS_ERR_NCCL 6

#define MEMS_MODE_MEMHASH 0   /* MemHash::FindMatches, multi-MUM (MemHash.cpp:109-162) */
#define MEMS_MODE_REPEAT 1    /* RepeatHash::FindMatches (RepeatHash.cpp:34-62) */
#define MEMS_MODE_PAIRWISE 2  /* PairwiseMatchFinder::FindMatches (PairwiseMatchFinder.cpp:37-71) */

/* output ordering of a match list */
#define MEMS_ORDER_ANY 0       /* distinct matches in device order (deterministic, unspecified) — fastest */
#define MEMS_ORDER_REFERENCE 1 /* the reference's hash-table order incl. its collision drops
                                  (MemHash.cpp:209-251, MemHash.h:183-203); always used for RepeatHash */
#define MEMS_ORDER_CANONICAL 2 /* distinct matches sorted by (SeqCount, Length, Start(0..)) on the host */

#define MEMS_MAX_SEQS 64           /* sequences per FindMatches call */
#define MEMS_MER_REPEAT_LIMIT 1000 /* MatchFinder.cpp:166 — larger seed runs are outside the parity contract */

typedef struct mems_ctx* mems_ctx_t;
typedef struct mems_sml* mems_sml_t;
typedef struct mems_matches* mems_matches_t;
typedef struct mems_table* mems_table_t;

/* ---- seed patterns: SeedMasks.h:276-401 (host arithmetic, no device needed) ---- */
uint64_t mems_get_seed(int weight, int seed_rank);           /* getSeed */
uint64_t mems_get_solid_seed(int weight);                    /* getSolidSeed */
int mems_get_seed_length(uint64_t seed);                     /* getSeedLength */
int mems_get_seed_weight(uint64_t seed);                     /* getSeedWeight */
unsigned mems_get_default_seed_weight(uint64_t avg_seq_len); /* getDefaultSeedWeight */

/* ---- context ---- */
/* device: CUDA ordinal.  stream: a cudaStream_t to run on, or NULL for a private stream. */
int mems_ctx_create(int device, void* stream, mems_ctx_t* out);
void mems_ctx_destroy(mems_ctx_t ctx);
/* last error text of this context (ctx may be NULL for creation failures) */
const char* mems_last_error(mems_ctx_t ctx);
/* block until all work queued on the context's stream is done */
int mems_ctx_synchronize(mems_ctx_t ctx);
/* A context keeps the device memory of earlier calls (an arena of cudaMalloc'ed slabs) so that repeated work makes no
 * driver call.  mems_ctx_trim waits for the context's work and gives the slabs no live object (SML, MatchList in flight)
 * occupies back to the driver; *reserved_bytes (may be NULL) = what the context still holds. */
int mems_ctx_trim(mems_ctx_t ctx, uint64_t* reserved_bytes);
/* pinned host memory helpers (optional; any host pointer is accepted by the calls below) */
int mems_host_alloc(void** ptr, uint64_t bytes);
void mems_host_free(void* ptr);

/* ---- sorted mer lists ---- */
typedef struct {
	uint64_t length;      /* SortedMerList::Length()      — bases */
	uint64_t sml_length;  /* SortedMerList::SMLLength()   — seed positions (linear sequences) */
	uint64_t seed;        /* SortedMerList::Seed() */
	uint32_t seed_length; /* SeedLength() */
	uint32_t seed_weight; /* SeedWeight() */
	uint64_t seed_mask;   /* GetSeedMask(): top 2*weight bits */
	uint64_t mer_mask;    /* GetMerMask():  top 2*length bits */
} mems_sml_info_t;

/* DNAMemorySML::Create(seq, seed) (MemorySML.cpp:45-60, SortedMerList.cpp:786-824) for one linear
 * sequence of n ASCII bases.  The sorted list stays resident in device memory. */
int mems_sml_create(mems_ctx_t ctx, const char* seq, uint64_t n, uint64_t seed, mems_sml_t* out);
/* MatchList::CreateMemorySMLs (MatchList.h:408-435): one SML per sequence, same seed.  The batch
 * is extracted and sorted as one union so a following mems_find_matches over exactly these
 * handles (same order) reuses the sorted union instead of merging. */
int mems_sml_create_batch(mems_ctx_t ctx, int n_seqs, const char* const* seqs, const uint64_t* lens,
                          uint64_t seed, mems_sml_t* out);
void mems_sml_destroy(mems_sml_t sml);
/* MemorySML::Clone / DNAMemorySML::Clone (MemorySML.cpp:36-38): a second handle to the same sorted list (the device
 * data is shared and lives until the last handle is destroyed). */
int mems_sml_clone(mems_sml_t sml, mems_sml_t* out);
int mems_sml_info(mems_sml_t sml, mems_sml_info_t* out);
/* MemorySML::Read / operator[] (MemorySML.cpp:62-94): entries [offset, offset+count) of the sorted
 * list as (position, canonical mer).  Either output may be NULL.  *n_read receives the count. */
int mems_sml_read(mems_sml_t sml, uint64_t offset, uint64_t count, uint32_t* positions_out,
                  uint64_t* mers_out, uint64_t* n_read);
/* SortedMerList::GetSeedMer (forward, SortedMerList.cpp:726-762) and DNAMemorySML::GetSeedMer ==
 * GetDnaSeedMer (canonical, :764-769) at arbitrary positions.  Either output may be NULL. */
int mems_sml_seed_mers(mems_sml_t sml, const uint64_t* positions, uint64_t n, uint64_t* fwd_out,
                       uint64_t* dna_out);
/* SortedMerList::FindMer (SortedMerList.cpp:170-179): *found=1 and *index = a sorted-list index
 * whose masked mer equals query_mer's, else *found=0 and *index = insertion point. */
int mems_sml_find_mer(mems_sml_t sml, uint64_t query_mer, int* found, uint64_t* index);
/* The 2-bit packed sequence as SortedMerList::SetSequence lays it out (MSB-first uint32 words,
 * ceil(2n/32)+2 words, pad words zero). words_out may be NULL to query *n_words only. */
int mems_sml_packed(mems_sml_t sml, uint32_t* words_out, uint64_t* n_words);

/* SeedOccurrenceList::construct (SeedOccurrenceList.h:21-92): out[p] for every base position p (Length()
 * floats) = multiplicity of the seed at p, averaged over the seeds that contain p. */
int mems_sml_seed_occurrence(mems_sml_t sml, float* out);

/* ---- match finding ---- */
typedef struct {
	int mode;          /* MEMS_MODE_* */
	int order;         /* MEMS_ORDER_* */
	uint32_t table_size; /* MemHash::SetTableSize; 0 = DEFAULT_MEM_TABLE_SIZE 40000 (MemHash.h:30) */
	uint32_t reserved;
	mems_table_t table;  /* NULL, or a persistent MemHash table (mems_table_create): the call then runs in
	                        MEMS_ORDER_REFERENCE against that table and returns the table's WHOLE content, so several
	                        FindMatches calls with different seed patterns accumulate exactly like the reference's
	                        ClearSequences() + FindMatches() loop (ProgressiveAligner.cpp:619-653) */
	uint64_t seq_mask;   /* MaskedMemHash::SetMask (MaskedMemHash.h:22-32): with MEMS_MODE_MEMHASH keep only hits whose
	                        sequence set equals this mask, sequence 0 = most significant of n_smls bits; 0 = no filter */
	const uint64_t* start_points; /* NULL, or n_smls indices: MemHash::FindMatchesFromPosition (MemHash.cpp:117-127,
	                        MatchFinder.cpp:137-164) — sorted mer list g is searched from its entry start_points[g] on */
} mems_match_params_t;

typedef struct {
	uint64_t n_matches;
	uint64_t n_flat;      /* int64 entries mems_matches_copy writes */
	uint64_t n_hits;      /* seed hits handed to HashMatch (MemHash.cpp:167) */
	uint64_t mem_count;   /* MemHash::MemCount()          (ORDER_REFERENCE only, else == n_matches) */
	uint64_t collisions;  /* MemHash::MemCollisionCount() (ORDER_REFERENCE only, else n_hits - n_matches) */
	uint64_t max_run;     /* largest equal-seed run seen; > MEMS_MER_REPEAT_LIMIT voids parity */
	uint64_t n_segments;  /* groups of hits connected without a window test (extension work items) */
	uint32_t seq_count;   /* sequences searched */
	uint32_t seed_length;
	double host_replay_ms; /* host time of the reference-order table replay (ORDER_REFERENCE / RepeatHash), else 0 */
} mems_matches_info_t;

/* MemHash/RepeatHash/PairwiseMatchFinder::FindMatches (MemHash.cpp:109-127) over n_smls sorted mer
 * lists that share one seed: equal-seed runs -> hits -> ungapped extension (MatchFinder.h:219-374)
 * -> distinct matches. */
int mems_find_matches(mems_ctx_t ctx, int n_smls, const mems_sml_t* smls,
                      const mems_match_params_t* params, mems_matches_t* out);
/* Many independent small problems in ONE launch set: the gap re-anchoring callers run CreateMemorySMLs + FindMatches on
 * thousands of short sequence sets (ProgressiveAligner::pairwiseAnchorSearch, ProgressiveAligner.cpp:589-678, under
 * `omp parallel for` at :695; Aligner::SearchLCBGaps, Aligner.cpp:784-930), where a call costs launch latency, not
 * work.  Problem p owns the next n_seqs[p] entries of seqs / lens (at most MEMS_MAX_SEQS each, 256 sequences in all);
 * all problems share one seed pattern.  out receives n_problems match lists; out[p] is exactly what
 * mems_sml_create_batch + mems_find_matches return for problem p alone.  MEMS_MODE_MEMHASH or MEMS_MODE_PAIRWISE,
 * MEMS_ORDER_ANY or MEMS_ORDER_CANONICAL. */
int mems_find_matches_many(mems_ctx_t ctx, int n_problems, const int* n_seqs, const char* const* seqs, const uint64_t* lens,
                           uint64_t seed, const mems_match_params_t* params, mems_matches_t* out);
/* MemHash's mem_table as a persistent object (MemHash::Clear empties it, MemHash.cpp:76-93; ClearSequences keeps it) */
int mems_table_create(uint32_t table_size /* 0 = 40000 */, mems_table_t* out);
void mems_table_clear(mems_table_t t);
void mems_table_destroy(mems_table_t t);
/* MemHash::AddHashEntry (MemHash.cpp:209-251) for a match that is already extended — what MemHash::LoadFile
 * (MemHash.cpp:266-305) does per line of a .mems file: *inserted = 0 if the table already holds a match that
 * contains it on its diagonal (a collision), else 1.  mersize = the seed weight MatchFinder had when the line was
 * read (MatchFinder::mer_size; DNA_MER_SIZE before any search). */
int mems_table_add(mems_table_t t, uint32_t seq_count, uint64_t length, const int64_t* starts, uint32_t mersize, int* inserted);
/* The table's content in the reference's output order (MemHash::GetMatchList, MemHash.h:183-203) as a match list. */
int mems_table_matches(mems_table_t t, mems_matches_t* out);
int mems_matches_info(mems_matches_t m, mems_matches_info_t* out);
/* Flat records [SeqCount, Length, Start(0) .. Start(SeqCount-1)] per match; starts are 1-based,
 * negative = reverse strand, 0 = NO_MATCH (AbstractMatch.h:27, UngappedLocalAlignment.h:201-206). */
int mems_matches_copy(mems_matches_t m, int64_t* flat_out);
/* The same records in place (n_flat int64 values), valid until mems_matches_destroy: no copy. */
const int64_t* mems_matches_data(mems_matches_t m);
/* Records in MEMS_ORDER_ANY are delivered asynchronously: mems_find_matches[_sharded] returns when its last kernel is
 * queued and the device-to-host copy of the records runs behind it.  mems_matches_data / mems_matches_copy /
 * mems_matches_destroy wait for the copy themselves; mems_matches_wait only waits.  A caller that issues its next call
 * before touching the records of the previous one gets the copy overlapped with that call's kernels; the counts of
 * mems_matches_info are valid at once. */
int mems_matches_wait(mems_matches_t m);
void mems_matches_destroy(mems_matches_t m);

/* ---- sharded match finding: one process per GPU, NCCL over NVLink ----
 * Every rank extracts a contiguous block of the sequences; seed space is range-partitioned (owners balanced by
 * the global key histogram) and one all-to-all delivers each seed range to its owner, which finds that range's
 * hits; hits are re-partitioned by diagonal so that extension is disjoint across ranks.  Design precedent in the
 * reference: ParallelMemHash's seed-range chunks (ParallelMemHash.cpp:42-121, never built). */
typedef struct mems_comm* mems_comm_t;
/* rank 0 creates the 128-byte NCCL id and hands it to the other ranks by any host channel */
int mems_comm_unique_id(char* id_out /* 128 bytes */);
int mems_comm_create(mems_ctx_t ctx, const char* id /* 128 bytes */, int rank, int world, mems_comm_t* out);
/* Collective: every rank must call it (the ranks agree that nobody maps anybody's exchange window any more before
 * the windows are freed).  After a failed collective call the communicator must not be used for further calls. */
void mems_comm_destroy(mems_comm_t comm);
/* the block of sequences [first, first+count) rank `rank` must supply (host arithmetic) */
int mems_shard_sequence_range(int n_seqs, int rank, int world, int* first, int* count);
/* owner rank of each of the 256 top-key-digit buckets given their global counts (host arithmetic) */
int mems_shard_bucket_owners(const uint64_t* hist256, int world, uint8_t* owner256);
/* the seed-record exchange as derived from the gathered histograms (hist_all[q * 256 + b], host arithmetic, the
 * same on every rank): counts[q * world + p] = records rank q sends to rank p; for `rank`: src_elem[p] = start of its
 * slice for p in its partitioned order, dst_elem[p] = start of that slice inside p's receive region; *max_recv = the
 * largest receive region (ParallelMemHash's chunk bookkeeping, ParallelMemHash.cpp:75-82, made explicit) */
int mems_shard_exchange_plan(const uint32_t* hist_all, int world, int rank, const uint8_t* owner256, uint64_t* counts,
                             uint64_t* src_elem, uint64_t* dst_elem, uint64_t* max_recv);
/* Collective: all ranks call it with the same n_seqs / lens / seed / params; seqs[g] must be valid for the
 * sequences of this rank's block (others may be NULL).  MEMS_MODE_MEMHASH, ORDER_ANY or ORDER_CANONICAL.
 * *out holds THIS rank's share of the distinct matches; the union over ranks is the MatchList. */
int mems_find_matches_sharded(mems_ctx_t ctx, mems_comm_t comm, int n_seqs, const char* const* seqs,
                              const uint64_t* lens, uint64_t seed, const mems_match_params_t* params,
                              mems_matches_t* out);

/* Environment switches read by the library (defaults are the measured best; the others exist for tests and
 * experiments): MEMS_NO_PEER_WINDOWS=1 sharded exchanges through NCCL send/recv instead of CUDA-IPC exchange windows;
 * MEMS_PEER_SCATTER=1 the partition kernel scatters straight into the peers' windows instead of partition + DMA copies;
 * MEMS_NO_SIDE_COMM=1 no second NCCL communicator; MEMS_HOST_THREADS=n host threads of the reference-order table
 * replay (default min(16, cores)); MEMS_TRACE=1 stage timings on stderr. */

/* ---- measurement ---- */
/* With profiling on, every kernel launch is bracketed by CUDA events on the context's stream. */
int mems_profile_enable(mems_ctx_t ctx, int on);
int mems_profile_reset(mems_ctx_t ctx);
typedef struct {
	char name[48];
	uint64_t launches;
	double ms;       /* summed device time of those launches */
	double bytes;    /* summed algorithmic bytes the launches were given (0 if not tracked) */
} mems_profile_entry_t;
/* Synchronises the stream, then fills up to cap entries; *n receives the number available. */
int mems_profile_get(mems_ctx_t ctx, mems_profile_entry_t* entries, int cap, int* n);
/* kernels launched on this context since creation / last reset (counted even with profiling off) */
uint64_t mems_launch_count(mems_ctx_t ctx);

/* ---- testing ----
 * Shrinks internal budgets of ONE context so that small inputs reach the rare paths: hash_bits > 0 keeps only that many
 * bits of the diagonal hash (bucket collisions, de-dup of marked components); walk_budget > 0 = probes / rounds a walk
 * gets before it moves on to the CTA-wide and grid-wide walkers.  0 restores production behaviour.  Results are the
 * same either way. */
int mems_test_hooks(mems_ctx_t ctx, int hash_bits, int walk_budget);
/* Randomised self-check of the bookkeeping of the context's device-memory arena on made-up addresses (host code only,
 * runs without a GPU): 0 = passed, anything else names the violated invariant (csrc/context.cu). */
int mems_selftest_arena(uint64_t seed, int rounds);

#ifdef __cplusplus
}
#endif
#endif /* MEMS_B200_H */
