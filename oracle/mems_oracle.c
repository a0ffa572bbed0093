/* oracle/mems_oracle.c — CPU restatement (plain C) of libMems' anchoring hot path.
 *
 * TEST INFRASTRUCTURE ONLY — see mems_oracle.h.  Parity pinned against oracle/_ref (the unmodified
 * reference) by tests/test_oracle_vs_ref.py and against tests/golden/ fixtures.
 * Citations are file:line under /root/reference/libMems/.
 */
#define _GNU_SOURCE
#include "mems_oracle.h"
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static char g_err[256];
const char* orc_last_error(void) { return g_err; }
void orc_free(void* p) { free(p); }

/* ------------------------------------------------------------------ seeds (SeedMasks.h) */
static const uint64_t seed_table[32][6] = {
#include "seed_table.inc"
};

uint64_t orc_get_solid_seed(int weight) { /* SeedMasks.h:276-281 */
	uint64_t s = 1;
	s <<= weight;
	return s - 1;
}

uint64_t orc_get_seed(int weight, int rank) { /* SeedMasks.h:298-321 */
	if (rank == INT_MAX) return orc_get_solid_seed(weight);
	if (weight > 31) return orc_get_solid_seed(32);
	if (rank > 5) return orc_get_solid_seed(weight);
	if (weight < 0 || rank < 0) return 0; /* undefined in the reference */
	return seed_table[weight][rank];
}

int orc_seed_length(uint64_t seed) { /* SeedMasks.h:335-350: span between lowest and highest 1-bit */
	int lo = -1, hi = -1;
	for (int b = 0; b < 64; ++b)
		if ((seed >> b) & 1) {
			hi = b;
			if (lo < 0) lo = b;
		}
	return hi < 0 ? 0 : hi - lo + 1;
}

int orc_seed_weight(uint64_t seed) { /* SeedMasks.h:362-373 */
	int w = 0;
	for (int b = 0; b < 64; ++b) w += (int)((seed >> b) & 1);
	return w;
}

unsigned orc_default_seed_weight(uint64_t avg_len) { /* SeedMasks.h:389-401 */
	if (avg_len == 0) return 0;
	unsigned m = (unsigned)ceil((log((double)avg_len) / log(2.0)) / 1.5);
	if (!(m & 1)) ++m;
	if (m < 5) m = 0;
	if (m > 31) m = 31;
	return m;
}

/* ------------------------------------------------------------------ 2-bit packing */
static uint8_t dna_code(unsigned char c) { /* SortedMerList.cpp:29-47 */
	switch (c) {
	case 'c': case 'C': case 'b': case 'B': case 'y': case 'Y': return 1;
	case 'g': case 'G': case 's': case 'S': case 'k': case 'K': return 2;
	case 't': case 'T': return 3;
	default: return 0;
	}
}

uint64_t orc_packed_words(uint64_t n) { /* SortedMerList.cpp:306-311 */
	return (2 * n + 31) / 32 + 2;
}

int orc_pack(const char* seq, uint64_t n, uint32_t* words) { /* translate32, :425-460 */
	uint64_t nw = orc_packed_words(n);
	memset(words, 0, nw * sizeof(uint32_t));
	for (uint64_t i = 0; i < n; ++i) {
		if (seq[i] == '-') {
			snprintf(g_err, sizeof g_err, "Gap in genome sequence at %llu", (unsigned long long)i);
			return 1;
		}
		/* base i occupies bits 31-2(i%16) .. 30-2(i%16) of word i/16 (MSB first) */
		words[i >> 4] |= (uint32_t)dna_code((unsigned char)seq[i]) << (30 - 2 * (i & 15));
	}
	return 0;
}

/* ------------------------------------------------------------------ mers */
typedef struct {
	uint64_t seed;
	int L, w;
	uint64_t seed_mask; /* top 2w bits (SortedMerList.cpp:819-821) */
	uint64_t mer_mask;  /* top 2L bits */
} seedinfo;

static int seedinfo_init(seedinfo* si, uint64_t seed) { /* SortedMerList::Create checks, :786-798 */
	si->seed = seed;
	si->L = orc_seed_length(seed);
	si->w = orc_seed_weight(seed);
	if (si->L == 0) {
		snprintf(g_err, sizeof g_err, "Can't have 0 seed length");
		return 1;
	}
	if (si->L > 32) {
		snprintf(g_err, sizeof g_err, "Mer size is too large");
		return 1;
	}
	si->seed_mask = si->w >= 32 ? ~0ULL : (~0ULL << (64 - 2 * si->w));
	si->mer_mask = si->L >= 32 ? ~0ULL : (~0ULL << (64 - 2 * si->L));
	return 0;
}

/* SortedMerList::GetMer, :321-342 — 64-bit window starting at base `pos`, left-justified */
static uint64_t get_mer(const uint32_t* words, uint64_t pos, uint64_t mer_mask) {
	uint64_t word = (pos * 2) / 32;
	unsigned bit = (unsigned)((pos * 2) % 32);
	uint64_t m = ((uint64_t)words[word] << 32) | words[word + 1];
	if (bit > 0) m = (m << bit) | (words[word + 2] >> (32 - bit));
	return m & mer_mask;
}

/* SortedMerList::GetSeedMer, :726-762 — gather the bases under the pattern's 1-bits (pattern MSB
 * <-> window base 0), left-justify to the top 2w bits */
static uint64_t seed_mer_fwd(const uint32_t* words, uint64_t pos, const seedinfo* si) {
	uint64_t win = get_mer(words, pos, si->mer_mask);
	uint64_t out = 0;
	for (int i = 0; i < si->L; ++i) {
		if ((si->seed >> (si->L - 1 - i)) & 1) {
			uint64_t base = (win >> (62 - 2 * i)) & 3;
			out = (out << 2) | base;
		}
	}
	return out << (64 - 2 * si->w);
}

/* SortedMerList::RevCompMer, :597-614 — reverse complement of the top mer_length bases, strand flag in bit 0 */
static uint64_t revcomp_mer(uint64_t mer, int mer_length) {
	uint64_t b = ~mer, c = 0;
	for (int i = 0; i < 64; i += 2) {
		c |= b & 3;
		b >>= 2;
		c <<= 2;
	}
	/* the loop's last shift leaves the reversed bases in bits 63..2; drop the complemented padding */
	c <<= 64 - 2 * (mer_length + 1);
	return c | 1;
}

/* SortedMerList::GetDnaSeedMer, :764-769 */
static uint64_t dna_seed_mer(const uint32_t* words, uint64_t pos, const seedinfo* si) {
	uint64_t f = seed_mer_fwd(words, pos, si);
	uint64_t r = revcomp_mer(f, si->w);
	return f < r ? f : r;
}

static uint64_t sml_length(uint64_t n, int L) { /* SortedMerList::SMLLength, :288-295 (linear) */
	return n < (uint64_t)L ? 0 : n - L + 1;
}

int orc_seed_mers(const char* seq, uint64_t n, uint64_t seed, const uint64_t* pos, uint64_t npos,
                  uint64_t* fwd_out, uint64_t* dna_out) {
	seedinfo si;
	if (seedinfo_init(&si, seed)) return 1;
	uint32_t* words = (uint32_t*)malloc(orc_packed_words(n) * sizeof(uint32_t));
	if (orc_pack(seq, n, words)) {
		free(words);
		return 1;
	}
	for (uint64_t i = 0; i < npos; ++i) {
		if (fwd_out) fwd_out[i] = seed_mer_fwd(words, pos[i], &si);
		if (dna_out) dna_out[i] = dna_seed_mer(words, pos[i], &si);
	}
	free(words);
	return 0;
}

/* ------------------------------------------------------------------ SML */
typedef struct {
	uint64_t mer;
	uint32_t pos;
} bmer_t;

static int bmer_cmp(const void* a, const void* b) {
	const bmer_t* x = (const bmer_t*)a;
	const bmer_t* y = (const bmer_t*)b;
	if (x->mer != y->mer) return x->mer < y->mer ? -1 : 1; /* bmer_lessthan, SortedMerList.h:312-314 */
	return x->pos < y->pos ? -1 : (x->pos > y->pos);      /* tie-break: not specified by the reference */
}

/* canonical keys of every seed position of one sequence (FillDnaSeedSML, :771-783) */
static uint64_t* all_keys(const char* seq, uint64_t n, const seedinfo* si, uint64_t* count_out) {
	uint64_t cnt = sml_length(n, si->L);
	*count_out = cnt;
	uint32_t* words = (uint32_t*)malloc(orc_packed_words(n) * sizeof(uint32_t));
	if (orc_pack(seq, n, words)) {
		free(words);
		return NULL;
	}
	uint64_t* keys = (uint64_t*)malloc((cnt ? cnt : 1) * sizeof(uint64_t));
	for (uint64_t p = 0; p < cnt; ++p) keys[p] = dna_seed_mer(words, p, si);
	free(words);
	return keys;
}

int orc_sml_build(const char* seq, uint64_t n, uint64_t seed, uint32_t* positions_out,
                  uint64_t* mers_out, uint64_t* sml_len_out) {
	seedinfo si;
	if (seedinfo_init(&si, seed)) return 1;
	uint64_t cnt;
	uint64_t* keys = all_keys(seq, n, &si, &cnt);
	if (!keys) return 1;
	bmer_t* arr = (bmer_t*)malloc((cnt ? cnt : 1) * sizeof(bmer_t));
	for (uint64_t p = 0; p < cnt; ++p) {
		arr[p].mer = keys[p];
		arr[p].pos = (uint32_t)p;
	}
	qsort(arr, cnt, sizeof(bmer_t), bmer_cmp); /* MemorySML.cpp:54 */
	for (uint64_t i = 0; i < cnt; ++i) {
		if (positions_out) positions_out[i] = arr[i].pos;
		if (mers_out) mers_out[i] = arr[i].mer;
	}
	if (sml_len_out) *sml_len_out = cnt;
	free(arr);
	free(keys);
	return 0;
}

/* ------------------------------------------------------------------ match finding */
typedef struct {
	int seqcount;
	int64_t len;
	int64_t mersize; /* L for a probing seed hit, 0 for stored copies (MatchHashEntry.cpp:118-126) */
	int64_t offset;
	int64_t* start; /* seqcount entries, 0 = NO_MATCH (AbstractMatch.h:27) */
} mhe_t;

typedef struct {
	int mode;
	int n_seqs;
	seedinfo si;
	const uint64_t* lens;
	uint64_t** keys; /* per sequence: canonical key at every seed position */
	uint64_t* nkeys;
} finder_t;

/* RepeatHash::GetSar always returns SML 0 (RepeatHash.h:37-40) */
static int sar_of(const finder_t* f, int seqI) { return f->mode == ORC_MODE_REPEAT ? 0 : seqI; }

static mhe_t* mhe_new(int seqcount) {
	mhe_t* e = (mhe_t*)malloc(sizeof(mhe_t));
	e->seqcount = seqcount;
	e->start = (int64_t*)calloc((size_t)seqcount, sizeof(int64_t));
	e->len = 0;
	e->mersize = 0;
	e->offset = 0;
	return e;
}
static void mhe_free(mhe_t* e) {
	free(e->start);
	free(e);
}
static int64_t mhe_start(const mhe_t* e, int i) { return i < e->seqcount ? e->start[i] : 0; }
static int mhe_first(const mhe_t* e) { /* HybridAbstractMatch::FirstStart */
	for (int i = 0; i < e->seqcount; ++i)
		if (e->start[i] != 0) return i;
	return INT_MAX;
}

/* MatchHashEntry::CalculateOffset, MatchHashEntry.cpp:141-160 */
static void mhe_calc_offset(mhe_t* e) {
	e->offset = 0;
	int i = mhe_first(e);
	if (i == INT_MAX) return;
	int64_t ref = e->start[i];
	for (++i; i < e->seqcount; ++i)
		if (e->start[i] != 0) {
			int64_t t = e->start[i] - ref;
			if (e->start[i] < 0) t -= e->len;
			e->offset += t;
		}
}

/* MatchHashEntry::Contains, MatchHashEntry.cpp:164-200 — does A contain m? */
static int mhe_contains(const mhe_t* A, const mhe_t* m) {
	if (A->seqcount != m->seqcount || A->offset != m->offset) return 0;
	int i = mhe_first(m);
	int64_t diff = mhe_start(m, i) - mhe_start(A, i);
	if (mhe_start(A, i) == 0) return 0;
	if (diff < 0 || A->len < m->len + diff) return 0;
	int64_t diff_rc = m->len - A->len + diff;
	for (++i; i < m->seqcount; ++i) {
		int64_t d = m->start[i] - A->start[i];
		if (m->start[i] == 0 && A->start[i] == 0) continue;
		else if (m->start[i] < 0 && diff_rc == d) continue;
		else if (diff != d) return 0;
	}
	return 1;
}

/* MatchHashEntry::strict_start_lessthan_ptr, MatchHashEntry.cpp:48-67 */
static int mhe_strict_less(const mhe_t* a, const mhe_t* b) {
	int fa = mhe_first(a), fb = mhe_first(b);
	int start_diff = fa - fb;
	if (start_diff == 0) {
		int cnt = a->seqcount <= b->seqcount ? a->seqcount : b->seqcount;
		for (int s = 0; s < cnt; ++s) {
			int64_t x = a->start[s], y = b->start[s];
			if (x < 0) x = -x + a->len - a->mersize;
			if (y < 0) y = -y + b->len - b->mersize;
			if (x != y) return x < y;
		}
	}
	return start_diff < 0;
}

/* MheCompare, MatchHashEntry.h:121-143 */
static int mhe_compare(const mhe_t* a, const mhe_t* b) {
	int fa = mhe_first(a), fb = mhe_first(b);
	if (fa > fb) return 1;
	if (fa == fb) {
		for (int i = fa; i < a->seqcount; ++i) {
			int64_t as = mhe_start(a, i), bs = mhe_start(b, i);
			if (as == 0 && bs != 0) return 1;
			else if (as != 0 && bs == 0) return 0;
		}
		if (mhe_contains(a, b) || mhe_contains(b, a)) return 0;
		return mhe_strict_less(a, b);
	}
	return 0;
}

typedef struct {
	mhe_t** v;
	size_t n, cap;
} bucket_t;

/* std::lower_bound (libstdc++ bisection) with MheCompare */
static size_t bucket_lower_bound(const bucket_t* b, const mhe_t* x) {
	size_t first = 0, len = b->n;
	while (len > 0) {
		size_t half = len >> 1;
		size_t mid = first + half;
		if (mhe_compare(b->v[mid], x)) {
			first = mid + 1;
			len = len - half - 1;
		} else
			len = half;
	}
	return first;
}
static void bucket_insert(bucket_t* b, size_t at, mhe_t* e) {
	if (b->n == b->cap) {
		b->cap = b->cap ? b->cap * 2 : 4;
		b->v = (mhe_t**)realloc(b->v, b->cap * sizeof(mhe_t*));
	}
	memmove(b->v + at + 1, b->v + at, (b->n - at) * sizeof(mhe_t*));
	b->v[at] = e;
	b->n++;
}

/* The window test of MatchFinder::ExtendMatch (MatchFinder.h:264-308): with the match's current
 * starts/length, does window index k (0 = leftmost window of the match) hold a matching seed in
 * every member?  Member g's 0-based seed position: start>0: start-1+k ; start<0: |start|-1+(len-L-k). */
static int window_matches(const finder_t* f, const mhe_t* e, int64_t k) {
	int have = 0;
	uint64_t key0 = 0;
	int par0 = 0;
	const int L = f->si.L;
	for (int s = 0; s < e->seqcount; ++s) {
		int64_t st = e->start[s];
		if (st == 0) continue;
		int g = sar_of(f, s);
		int64_t p = st > 0 ? st - 1 + k : -st - 1 + (e->len - L - k);
		if (p < 0 || (uint64_t)p >= f->nkeys[g]) return 0;
		uint64_t key = f->keys[g][p];
		int parity = st < 0 ? (int)(key & 1) : !(key & 1);
		key &= f->si.seed_mask;
		if (!have) {
			key0 = key;
			par0 = parity;
			have = 1;
		} else if (key != key0 || parity != par0)
			return 0;
	}
	return 1;
}

/* Fixed point of MatchFinder::ExtendMatch (MatchFinder.h:219-374): grow the match left and right
 * to the farthest matching seed window within L of its current ends until neither end moves
 * (SURVEY.md Appendix A.3). */
static void extend_match(const finder_t* f, mhe_t* e) {
	const int L = f->si.L;
	for (;;) {
		int moved = 0;
		/* room on the "backward" side: forward members grow at their left end, reverse members at
		 * their right end (MatchFinder.h:241-257) */
		int64_t room = INT64_MAX;
		for (int s = 0; s < e->seqcount; ++s) {
			int64_t st = e->start[s];
			if (st == 0) continue;
			int64_t n = (int64_t)f->lens[sar_of(f, s)];
			int64_t r = st > 0 ? st - 1 : n - (-st + e->len - 1);
			if (r < room) room = r;
		}
		int64_t dmax = room < L ? room : L, best = 0;
		for (int64_t d = dmax; d >= 1; --d)
			if (window_matches(f, e, -d)) {
				best = d;
				break;
			}
		if (best) {
			for (int s = 0; s < e->seqcount; ++s)
				if (e->start[s] > 0) e->start[s] -= best;
			e->len += best;
			moved = 1;
		}
		room = INT64_MAX;
		for (int s = 0; s < e->seqcount; ++s) {
			int64_t st = e->start[s];
			if (st == 0) continue;
			int64_t n = (int64_t)f->lens[sar_of(f, s)];
			int64_t r = st > 0 ? n - (st + e->len - 1) : -st - 1;
			if (r < room) room = r;
		}
		dmax = room < L ? room : L;
		best = 0;
		for (int64_t d = dmax; d >= 1; --d)
			if (window_matches(f, e, e->len - L + d)) {
				best = d;
				break;
			}
		if (best) {
			for (int s = 0; s < e->seqcount; ++s)
				if (e->start[s] < 0) e->start[s] += best; /* magnitude shrinks: the left end moves left */
			e->len += best;
			moved = 1;
		}
		if (!moved) break;
	}
}

typedef struct {
	bucket_t* buckets;
	uint32_t table_size;
	uint64_t mem_count, collision_count, hit_count;
} table_t;

/* MemHash::AddHashEntry, MemHash.cpp:209-251 */
static void add_hash_entry(const finder_t* f, table_t* t, mhe_t* probe) {
	int64_t ts = (int64_t)t->table_size;
	uint32_t bi = (uint32_t)(((probe->offset % ts) + ts) % ts);
	bucket_t* b = &t->buckets[bi];
	t->hit_count++;
	size_t at = bucket_lower_bound(b, probe);
	if (at != b->n && !mhe_compare(b->v[at], probe) && !mhe_compare(probe, b->v[at])) {
		t->collision_count++;
		mhe_free(probe);
		return;
	}
	extend_match(f, probe);
	probe->mersize = 0; /* stored copy: MatchHashEntry::operator= zeroes m_mersize */
	at = bucket_lower_bound(b, probe);
	bucket_insert(b, at, probe);
	t->mem_count++;
}

typedef struct {
	uint64_t masked;
	uint32_t seq;
	uint32_t pos;
	uint8_t strand;
} occ_t;

static int occ_cmp(const void* a, const void* b) {
	const occ_t* x = (const occ_t*)a;
	const occ_t* y = (const occ_t*)b;
	if (x->masked != y->masked) return x->masked < y->masked ? -1 : 1;
	if (x->seq != y->seq) return x->seq < y->seq ? -1 : 1;
	return x->pos < y->pos ? -1 : (x->pos > y->pos);
}

/* MemHash::HashMatch + SetDirection (MemHash.cpp:167-203): members are (seq slot, 0-based pos, strand) */
static void hash_match(const finder_t* f, table_t* t, int seqcount, const int* slot, const occ_t* const* occ,
                       int n_members) {
	mhe_t* e = mhe_new(seqcount);
	e->len = f->si.L;
	e->mersize = f->si.L;
	for (int i = 0; i < n_members; ++i) e->start[slot[i]] = (int64_t)occ[i]->pos + 1;
	/* SetDirection: the first present member stays positive; others are negated iff their strand differs */
	int first = -1;
	int first_strand = 0;
	for (int i = 0; i < n_members; ++i)
		if (first < 0 || slot[i] < first) {
			first = slot[i];
			first_strand = occ[i]->strand;
		}
	for (int i = 0; i < n_members; ++i)
		if (slot[i] != first && occ[i]->strand != first_strand) e->start[slot[i]] = -e->start[slot[i]];
	mhe_calc_offset(e);
	add_hash_entry(f, t, e);
}

static table_t* g_accum = NULL; /* open accumulation, see orc_accumulate_begin */
static uint64_t g_seq_mask = 0; /* MaskedMemHash::SetMask; 0 = no filter (MaskedMemHash.h:22-32) */

int orc_find_matches_masked(int n_seqs, const char* const* seqs, const uint64_t* lens, uint64_t seed,
                            uint64_t seq_mask, int64_t** flat_out, uint64_t* n_flat_out,
                            uint64_t* n_matches_out, uint64_t* counts_out) {
	g_seq_mask = seq_mask;
	int rc = orc_find_matches(ORC_MODE_MEMHASH, n_seqs, seqs, lens, seed, flat_out, n_flat_out, n_matches_out, counts_out);
	g_seq_mask = 0;
	return rc;
}

int orc_find_matches(int mode, int n_seqs, const char* const* seqs, const uint64_t* lens,
                     uint64_t seed, int64_t** flat_out, uint64_t* n_flat_out,
                     uint64_t* n_matches_out, uint64_t* counts_out) {
	finder_t f;
	memset(&f, 0, sizeof f);
	f.mode = mode;
	f.n_seqs = n_seqs;
	f.lens = lens;
	if (seedinfo_init(&f.si, seed)) return 1;
	if (mode == ORC_MODE_REPEAT && n_seqs != 1) {
		snprintf(g_err, sizeof g_err, "RepeatHash needs exactly one sequence (RepeatHash.cpp:26-32)");
		return 1;
	}
	f.keys = (uint64_t**)calloc((size_t)n_seqs, sizeof(uint64_t*));
	f.nkeys = (uint64_t*)calloc((size_t)n_seqs, sizeof(uint64_t));
	uint64_t total = 0;
	for (int g = 0; g < n_seqs; ++g) {
		f.keys[g] = all_keys(seqs[g], lens[g], &f.si, &f.nkeys[g]);
		if (!f.keys[g]) return 1;
		total += f.nkeys[g];
	}
	/* MatchFinder::SearchRange (MatchFinder.cpp:172-340) walks the distinct masked keys of all SMLs in
	 * ascending order; sorting the union by (masked key, sequence, position) visits the same runs. */
	occ_t* occ = (occ_t*)malloc((total ? total : 1) * sizeof(occ_t));
	uint64_t o = 0;
	for (int g = 0; g < n_seqs; ++g)
		for (uint64_t p = 0; p < f.nkeys[g]; ++p) {
			occ[o].masked = f.keys[g][p] & f.si.seed_mask;
			occ[o].strand = (uint8_t)(f.keys[g][p] & 1);
			occ[o].seq = (uint32_t)g;
			occ[o].pos = (uint32_t)p;
			++o;
		}
	qsort(occ, total, sizeof(occ_t), occ_cmp);

	/* the table lives across calls while an accumulation is open (MemHash::ClearSequences keeps mem_table,
	 * MemHash.cpp:72-74; several seed patterns then share one table, ProgressiveAligner.cpp:619-653) */
	table_t t_local;
	table_t* tp = g_accum ? g_accum : &t_local;
	if (!g_accum) {
		memset(tp, 0, sizeof *tp);
		tp->table_size = 40000; /* DEFAULT_MEM_TABLE_SIZE, MemHash.h:30 */
		tp->buckets = (bucket_t*)calloc(tp->table_size, sizeof(bucket_t));
	}
#define t (*tp)
	uint64_t max_run = 0;

	int* slot = (int*)malloc(sizeof(int) * 1024);
	const occ_t** mem = (const occ_t**)malloc(sizeof(occ_t*) * 1024);
	size_t memcap = 1024;
	for (uint64_t s = 0; s < total;) {
		uint64_t e = s + 1;
		while (e < total && occ[e].masked == occ[s].masked) ++e;
		uint64_t run = e - s;
		if (run > max_run) max_run = run;
		if (run >= 2) {
			if (run > memcap) {
				memcap = run;
				slot = (int*)realloc(slot, sizeof(int) * memcap);
				mem = (const occ_t**)realloc(mem, sizeof(occ_t*) * memcap);
			}
			if (mode == ORC_MODE_MEMHASH) {
				/* MemHash::EnumerateMatches with repeat_tolerance 0, enumeration_tolerance 1 (MemHash.cpp:139-162):
				 * a second occurrence in any sequence discards the whole run */
				int ok = 1;
				for (uint64_t i = s + 1; i < e; ++i)
					if (occ[i].seq == occ[i - 1].seq) ok = 0;
				if (ok && g_seq_mask) {
					/* MaskedMemHash::HashMatch (MaskedMemHash.cpp:38-63): "match number" has sequence 0 in the
					 * most significant of n_seqs bits; only hits whose number equals the mask are hashed */
					uint64_t number = 0;
					for (int g = 0; g < n_seqs; ++g) {
						number <<= 1;
						for (uint64_t i = s; i < e; ++i)
							if ((int)occ[i].seq == g) number |= 1;
					}
					if (number != g_seq_mask) ok = 0;
				}
				if (ok) {
					for (uint64_t i = s; i < e; ++i) {
						slot[i - s] = (int)occ[i].seq;
						mem[i - s] = &occ[i];
					}
					hash_match(&f, &t, n_seqs, slot, mem, (int)run);
				}
			} else if (mode == ORC_MODE_REPEAT) {
				/* RepeatHash::HashMatch (RepeatHash.cpp:40-62): occurrences sorted by position become sequences 0..k-1 */
				for (uint64_t i = s; i < e; ++i) {
					slot[i - s] = (int)(i - s);
					mem[i - s] = &occ[i];
				}
				hash_match(&f, &t, (int)run, slot, mem, (int)run);
			} else {
				/* PairwiseMatchFinder::EnumerateMatches (PairwiseMatchFinder.cpp:37-71): sequences holding
				 * the key exactly once, every pair i<j */
				uint64_t nu = 0;
				for (uint64_t i = s; i < e;) {
					uint64_t j = i + 1;
					while (j < e && occ[j].seq == occ[i].seq) ++j;
					if (j - i == 1) mem[nu++] = &occ[i];
					i = j;
				}
				for (uint64_t a = 0; a < nu; ++a)
					for (uint64_t b = a + 1; b < nu; ++b) {
						const occ_t* pr[2] = {mem[a], mem[b]};
						int sl[2] = {(int)mem[a]->seq, (int)mem[b]->seq};
						hash_match(&f, &t, n_seqs, sl, pr, 2);
					}
			}
		}
		s = e;
	}
	free(slot);
	free(mem);

	/* MemHash::GetMatchList (MemHash.h:183-203): buckets in index order, each front to back */
	uint64_t nflat = 0, nm = 0;
	for (uint32_t b = 0; b < t.table_size; ++b)
		for (size_t i = 0; i < t.buckets[b].n; ++i) {
			nflat += 2 + (uint64_t)t.buckets[b].v[i]->seqcount;
			++nm;
		}
	int64_t* flat = (int64_t*)malloc((nflat ? nflat : 1) * sizeof(int64_t));
	uint64_t w = 0;
	for (uint32_t b = 0; b < t.table_size; ++b) {
		for (size_t i = 0; i < t.buckets[b].n; ++i) {
			mhe_t* e = t.buckets[b].v[i];
			flat[w++] = e->seqcount;
			flat[w++] = e->len;
			for (int s = 0; s < e->seqcount; ++s) flat[w++] = e->start[s];
			if (!g_accum) mhe_free(e);
		}
		if (!g_accum) free(t.buckets[b].v);
	}
	if (!g_accum) free(t.buckets);
	free(occ);
	for (int g = 0; g < n_seqs; ++g) free(f.keys[g]);
	free(f.keys);
	free(f.nkeys);
	*flat_out = flat;
	*n_flat_out = nflat;
	*n_matches_out = nm;
	if (counts_out) {
		counts_out[0] = t.mem_count;
		counts_out[1] = t.collision_count;
		counts_out[2] = t.hit_count;
		counts_out[3] = max_run;
	}
	return 0;
#undef t
}

int orc_accumulate_begin(void) {
	if (g_accum) return 1;
	g_accum = (table_t*)calloc(1, sizeof(table_t));
	g_accum->table_size = 40000;
	g_accum->buckets = (bucket_t*)calloc(g_accum->table_size, sizeof(bucket_t));
	return 0;
}

void orc_accumulate_end(void) {
	if (!g_accum) return;
	for (uint32_t b = 0; b < g_accum->table_size; ++b) {
		for (size_t i = 0; i < g_accum->buckets[b].n; ++i) mhe_free(g_accum->buckets[b].v[i]);
		free(g_accum->buckets[b].v);
	}
	free(g_accum->buckets);
	free(g_accum);
	g_accum = NULL;
}

/* ------------------------------------------------------------------ SeedOccurrenceList */
int orc_seed_occurrence(const char* seq, uint64_t n, uint64_t seed, float* count) {
	/* SeedOccurrenceList::construct, SeedOccurrenceList.h:21-63 */
	seedinfo si;
	if (seedinfo_init(&si, seed)) return 1;
	uint64_t cnt = sml_length(n, si.L);
	uint32_t* pos = (uint32_t*)malloc((cnt ? cnt : 1) * sizeof(uint32_t));
	uint64_t* mers = (uint64_t*)malloc((cnt ? cnt : 1) * sizeof(uint64_t));
	if (orc_sml_build(seq, n, seed, pos, mers, NULL)) return 1;
	for (uint64_t i = 0; i < n; ++i) count[i] = 0.f;
	for (uint64_t s = 0; s < cnt;) {
		uint64_t e = s + 1;
		while (e < cnt && (mers[e] & si.seed_mask) == (mers[s] & si.seed_mask)) ++e;
		for (uint64_t i = s; i < e; ++i) count[pos[i]] = (float)(e - s);
		s = e;
	}
	for (uint64_t i = cnt ? cnt : 1; i < n; ++i) count[i] = 1.f; /* :53-54 (seedI starts at 1) */
	/* smoothFrequencies, :73-90 */
	if (n > 0) {
		uint64_t L = (uint64_t)si.L;
		double sum = (double)(L - 1) + count[0];
		float* buf = (float*)malloc(L * sizeof(float));
		for (uint64_t i = 0; i < L; ++i) buf[i] = 1.f;
		buf[0] = count[0];
		for (uint64_t i = 1; i < n; ++i) {
			float ci = count[i];
			count[i - 1] = (float)(sum / (double)L);
			sum += ci;
			uint64_t bi = i % L;
			sum -= buf[bi];
			buf[bi] = ci;
		}
		free(buf);
	}
	for (uint64_t i = 0; i < n; ++i)
		if (count[i] == 0.f) count[i] = 1.f;
	free(pos);
	free(mers);
	return 0;
}
