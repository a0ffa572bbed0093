// oracle/ref_driver.cpp — C entry points over the UNMODIFIED libMems reference sources.
//
// TEST INFRASTRUCTURE ONLY (checker + CPU baseline).  This file contains no algorithm: it
// instantiates the reference's own classes (DNAMemorySML, MemHash, RepeatHash, MatchList —
// /root/reference/libMems/*.cpp compiled where they lie, see oracle/Makefile) and copies their
// results into flat C arrays so that Python tests and bench.py can read them through ctypes.
// The built library lives in oracle/_ref/ (git-ignored) and is never linked into the product.
#include "libMems/DNAMemorySML.h"
#include "libMems/MemHash.h"
#include "libMems/RepeatHash.h"
#include "libMems/PairwiseMatchFinder.h"
#include "libMems/MaskedMemHash.h"
#include "libMems/MatchList.h"
#include "libMems/SeedMasks.h"
#include "libMems/SeedOccurrenceList.h"

#include <chrono>
#include <cstddef>
#include <vector>
#include <string>
#include <cstring>
#include <sstream>

using namespace mems;
using namespace genome;
using namespace std;

// the reference's EliminateOverlaps, unmodified (cut out of Aligner.cpp by oracle/Makefile)
namespace mems {
#include "eliminate_overlaps.inc"
}

namespace {
double now_s() {
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
std::string g_err;
uint64 g_seq_mask = 0;
MemHash* g_accum_mh = NULL;
}

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// SeedMasks.h:298-401
uint64_t ref_get_seed(int weight, int rank) { return (uint64_t)getSeed(weight, rank); }
int ref_seed_length(uint64_t seed) { return getSeedLength((int64)seed); }
int ref_seed_weight(uint64_t seed) { return getSeedWeight((int64)seed); }
unsigned ref_default_seed_weight(uint64_t avg_len) { return getDefaultSeedWeight(avg_len); }

// DNAMemorySML::Create (MemorySML.cpp:45-60) on one sequence.
// positions_out / mers_out (both optional) receive SMLLength() entries: sml[i].position / sml[i].mer.
// packed_out (optional) receives the 2-bit words of SortedMerList::SetSequence without the two pad words.
int ref_sml_build(const char* seq, uint64_t n, uint64_t seed, uint32_t* positions_out, uint64_t* mers_out,
                  uint64_t* sml_len_out, uint64_t* seed_mask_out, uint64_t* mer_mask_out, double* secs_out) {
	try {
		gnSequence gs(seq, n);
		DNAMemorySML sml;
		double t0 = now_s();
		sml.Create(gs, seed);
		double t1 = now_s();
		if (secs_out) *secs_out = t1 - t0;
		uint64_t len = sml.SMLLength();
		if (sml_len_out) *sml_len_out = len;
		if (seed_mask_out) *seed_mask_out = sml.GetSeedMask();
		if (mer_mask_out) *mer_mask_out = sml.GetMerMask();
		if (positions_out || mers_out) {
			for (uint64_t i = 0; i < len; ++i) {
				bmer b = sml[i];
				if (positions_out) positions_out[i] = b.position;
				if (mers_out) mers_out[i] = b.mer;
			}
		}
		return 0;
	} catch (gnException& e) {
		g_err = e.code.name + ": " + e.msg;
		return 1;
	} catch (const char* s) {
		g_err = s;
		return 2;
	}
}

// The bytes FileSML::Create would write for this sequence (FileSML.cpp:344-366: header struct, SetSequence's
// words, sorted positions), assembled from the reference's own in-memory DNAMemorySML — FileSML.cpp itself
// needs the whole aligner to compile.  header.version is set like DNAFileSML's constructors do
// (DNAFileSML.cpp:22-31).  layout_out (optional, 16 entries) receives sizeof(SMLHeader) and the offset of
// every field, so the façade's copy of the struct can be pinned to the reference's header.
namespace {
struct ExposedSML : public DNAMemorySML {
	const uint32* words() const { return sequence; }
	uint64 n_words() const { return binary_seq_len; }
	SMLHeader& hdr() { return header; }
};
}
int64_t ref_sml_file_image(const char* seq, uint64_t n, uint64_t seed, uint8_t* out, uint64_t cap, uint32_t* layout_out) {
	try {
		if (layout_out) {
			uint32_t* l = layout_out;
			*l++ = (uint32_t)sizeof(SMLHeader);
			*l++ = (uint32_t)offsetof(SMLHeader, version);
			*l++ = (uint32_t)offsetof(SMLHeader, alphabet_bits);
			*l++ = (uint32_t)offsetof(SMLHeader, seed);
			*l++ = (uint32_t)offsetof(SMLHeader, seed_length);
			*l++ = (uint32_t)offsetof(SMLHeader, seed_weight);
			*l++ = (uint32_t)offsetof(SMLHeader, length);
			*l++ = (uint32_t)offsetof(SMLHeader, unique_mers);
			*l++ = (uint32_t)offsetof(SMLHeader, word_size);
			*l++ = (uint32_t)offsetof(SMLHeader, little_endian);
			*l++ = (uint32_t)offsetof(SMLHeader, id);
			*l++ = (uint32_t)offsetof(SMLHeader, circular);
			*l++ = (uint32_t)offsetof(SMLHeader, translation_table);
			*l++ = (uint32_t)offsetof(SMLHeader, description);
			*l++ = (uint32_t)sizeof(smlSeqI_t);
			*l++ = 0;
		}
		gnSequence gs(seq, n);
		ExposedSML sml;
		sml.Create(gs, seed);
		sml.hdr().version = 5;  // DNAFileSML::FormatVersion()
		const uint64_t len = sml.SMLLength();
		const uint64_t total = sizeof(SMLHeader) + sml.n_words() * sizeof(uint32) + len * sizeof(smlSeqI_t);
		if (!out) return (int64_t)total;
		if (cap < total) {
			g_err = "buffer too small";
			return -1;
		}
		uint8_t* w = out;
		SMLHeader h = sml.GetHeader();
		memcpy(w, &h, sizeof h);
		w += sizeof h;
		memcpy(w, sml.words(), sml.n_words() * sizeof(uint32));
		w += sml.n_words() * sizeof(uint32);
		std::vector<bmer> all;
		sml.Read(all, len, 0);
		for (uint64_t i = 0; i < len; ++i) {
			memcpy(w, &all[i].position, sizeof(smlSeqI_t));
			w += sizeof(smlSeqI_t);
		}
		return (int64_t)total;
	} catch (gnException& e) {
		g_err = e.code.name + ": " + e.msg;
		return -1;
	} catch (const char* m) {
		g_err = m;
		return -1;
	}
}

// SortedMerList::GetSeedMer (forward, SortedMerList.cpp:726-762) and GetDnaSeedMer (canonical, :764-769)
// at a list of positions.
int ref_seed_mers(const char* seq, uint64_t n, uint64_t seed, const uint64_t* pos, uint64_t npos,
                  uint64_t* fwd_out, uint64_t* dna_out) {
	try {
		gnSequence gs(seq, n);
		DNAMemorySML sml;
		// Create() would also sort; SortedMerList::Create only packs and sets masks.
		sml.SortedMerList::Create(gs, seed);
		for (uint64_t i = 0; i < npos; ++i) {
			if (fwd_out) fwd_out[i] = sml.SortedMerList::GetSeedMer(pos[i]);
			if (dna_out) dna_out[i] = sml.GetDnaSeedMer(pos[i]);
		}
		return 0;
	} catch (gnException& e) {
		g_err = e.code.name + ": " + e.msg;
		return 1;
	}
}

// MatchList::CreateMemorySMLs + {MemHash,RepeatHash,PairwiseMatchFinder}::FindMatches.
// mode: 0 = MemHash (multi-MUM), 1 = RepeatHash (single genome self-match), 2 = PairwiseMatchFinder.
// Output (malloc'd, release with ref_free): flat int64 records in the reference's own output order
//   [SeqCount, Length, Start(0) … Start(SeqCount-1)] per match.
// times_out[0] = SML build seconds (all genomes), times_out[1] = FindMatches seconds.
// counts_out[0] = MemCount, counts_out[1] = MemCollisionCount.
int ref_find_matches(int mode, int n_seqs, const char* const* seqs, const uint64_t* lens, uint64_t seed,
                     int64_t** flat_out, uint64_t* n_flat_out, uint64_t* n_matches_out,
                     double* times_out, uint64_t* counts_out) {
	try {
		MatchList ml;
		for (int g = 0; g < n_seqs; ++g) {
			ml.seq_table.push_back(new gnSequence(seqs[g], lens[g]));
			ml.seq_filename.push_back("mem");
		}
		double t0 = now_s();
		for (int g = 0; g < n_seqs; ++g) {
			DNAMemorySML* sml = new DNAMemorySML();
			sml->Create(*ml.seq_table[g], seed);
			ml.sml_table.push_back(sml);
		}
		double t1 = now_s();
		MemHash* mh = NULL;
		if (g_accum_mh) mh = g_accum_mh;
		else if (mode == 0) mh = new MemHash();
		else if (mode == 1) mh = new RepeatHash();
		else if (mode == 2) mh = new PairwiseMatchFinder();
		else if (mode == 3) {
			MaskedMemHash* mm = new MaskedMemHash();
			mm->SetMask(g_seq_mask);
			mh = mm;
		}
		else { g_err = "bad mode"; return 3; }
		if (g_accum_mh) mh->ClearSequences();  // keeps the table (MemHash.cpp:72-74)
		mh->FindMatches(ml);
		double t2 = now_s();
		if (times_out) { times_out[0] = t1 - t0; times_out[1] = t2 - t1; }
		if (counts_out) { counts_out[0] = mh->MemCount(); counts_out[1] = mh->MemCollisionCount(); }
		std::vector<int64_t> flat;
		for (size_t i = 0; i < ml.size(); ++i) {
			Match* m = ml[i];
			flat.push_back((int64_t)m->SeqCount());
			flat.push_back((int64_t)m->Length());
			for (uint s = 0; s < m->SeqCount(); ++s) flat.push_back((int64_t)m->Start(s));
		}
		*n_matches_out = ml.size();
		*n_flat_out = flat.size();
		*flat_out = (int64_t*)malloc(sizeof(int64_t) * (flat.size() ? flat.size() : 1));
		memcpy(*flat_out, flat.data(), sizeof(int64_t) * flat.size());
		if (!g_accum_mh) {
			mh->Clear();
			delete mh;
		}
		ml.Clear();
		return 0;
	} catch (gnException& e) {
		g_err = e.code.name + ": " + e.msg;
		return 1;
	} catch (const char* s) {
		g_err = s;
		return 2;
	}
}

// One MemHash kept across ref_find_matches calls (mode 0), the way ProgressiveAligner.cpp:619-653 accumulates the
// matches of several seed patterns: Clear() once, then ClearSequences() + FindMatches() per pattern.
void ref_accumulate_begin() {
	if (!g_accum_mh) g_accum_mh = new MemHash();
	g_accum_mh->Clear();
}
void ref_accumulate_end() {
	if (g_accum_mh) {
		g_accum_mh->Clear();
		delete g_accum_mh;
		g_accum_mh = NULL;
	}
}

// SeedOccurrenceList::construct (SeedOccurrenceList.h:21-63) — per-position seed multiplicity, smoothed.
int ref_seed_occurrence(const char* seq, uint64_t n, uint64_t seed, float* out, uint64_t* n_out) {
	try {
		gnSequence gs(seq, n);
		DNAMemorySML sml;
		sml.Create(gs, seed);
		SeedOccurrenceList sol;
		sol.construct(sml);
		*n_out = sml.Length();  // construct() sizes its table to sml.Length()
		if (out)
			for (uint64_t i = 0; i < sml.Length(); ++i) out[i] = sol.getFrequency(i);
		return 0;
	} catch (gnException& e) {
		g_err = e.code.name + ": " + e.msg;
		return 1;
	}
}

// ReadList (MatchList.h:498-587): parse .mums text with the reference's own reader.
int ref_read_list(const char* text, int64_t** flat_out, uint64_t* n_flat_out, uint64_t* n_matches_out) {
	try {
		std::istringstream is(text);
		MatchList ml;
		ReadList(ml, is);
		std::vector<int64_t> flat;
		for (size_t i = 0; i < ml.size(); ++i) {
			flat.push_back((int64_t)ml[i]->SeqCount());
			flat.push_back((int64_t)ml[i]->Length());
			for (uint s = 0; s < ml[i]->SeqCount(); ++s) flat.push_back((int64_t)ml[i]->Start(s));
		}
		*n_matches_out = ml.size();
		*n_flat_out = flat.size();
		*flat_out = (int64_t*)malloc(sizeof(int64_t) * (flat.size() ? flat.size() : 1));
		memcpy(*flat_out, flat.data(), sizeof(int64_t) * flat.size());
		for (size_t i = 0; i < ml.size(); ++i) ml[i]->Free();
		return 0;
	} catch (gnException& e) {
		g_err = e.code.name + ": " + e.msg;
		return 1;
	} catch (const char* s) {
		g_err = s;
		return 2;
	}
}

// WriteList (MatchList.h:589-634) of the MemHash result of two or more sequences.
int ref_write_list(int n_seqs, const char* const* seqs, const uint64_t* lens, uint64_t seed, char** text_out) {
	try {
		MatchList ml;
		for (int g = 0; g < n_seqs; ++g) {
			ml.seq_table.push_back(new gnSequence(seqs[g], lens[g]));
			ml.seq_filename.push_back("mem");
			DNAMemorySML* sml = new DNAMemorySML();
			sml->Create(*ml.seq_table[g], seed);
			ml.sml_table.push_back(sml);
		}
		MemHash mh;
		mh.FindMatches(ml);
		std::ostringstream os;
		WriteList(ml, os);
		*text_out = strdup(os.str().c_str());
		mh.Clear();
		ml.Clear();
		return 0;
	} catch (gnException& e) {
		g_err = e.code.name + ": " + e.msg;
		return 1;
	}
}

namespace {
void flatten(MatchList& ml, int64_t** flat_out, uint64_t* n_flat_out, uint64_t* n_matches_out) {
	std::vector<int64_t> flat;
	for (size_t i = 0; i < ml.size(); ++i) {
		flat.push_back((int64_t)ml[i]->SeqCount());
		flat.push_back((int64_t)ml[i]->Length());
		for (uint s = 0; s < ml[i]->SeqCount(); ++s) flat.push_back((int64_t)ml[i]->Start(s));
	}
	*n_matches_out = ml.size();
	*n_flat_out = flat.size();
	*flat_out = (int64_t*)malloc(sizeof(int64_t) * (flat.size() ? flat.size() : 1));
	memcpy(*flat_out, flat.data(), sizeof(int64_t) * flat.size());
}
void build_list(MatchList& ml, int n_seqs, const char* const* seqs, const uint64_t* lens, uint64_t seed) {
	for (int g = 0; g < n_seqs; ++g) {
		ml.seq_table.push_back(new gnSequence(seqs[g], lens[g]));
		ml.seq_filename.push_back("mem");
		DNAMemorySML* sml = new DNAMemorySML();
		sml->Create(*ml.seq_table[g], seed);
		ml.sml_table.push_back(sml);
	}
}
}

// MemHash::FindMatchesFromPosition (MemHash.cpp:117-127) with LogProgress and SetMatchLog streams attached
// (MatchFinder.cpp:298-309, MemHash.cpp:237-241); the two log texts are strdup'ed (ref_free).
int ref_find_matches_from(int n_seqs, const char* const* seqs, const uint64_t* lens, uint64_t seed, const uint64_t* start_points,
                          int64_t** flat_out, uint64_t* n_flat_out, uint64_t* n_matches_out, uint64_t* counts_out,
                          char** progress_out, char** match_log_out) {
	try {
		MatchList ml;
		build_list(ml, n_seqs, seqs, lens, seed);
		MemHash mh;
		std::ostringstream progress, matches;
		mh.LogProgress(&progress);
		mh.SetMatchLog(&matches);
		std::vector<gnSeqI> sp(start_points, start_points + n_seqs);
		mh.FindMatchesFromPosition(ml, sp);
		if (counts_out) { counts_out[0] = mh.MemCount(); counts_out[1] = mh.MemCollisionCount(); }
		flatten(ml, flat_out, n_flat_out, n_matches_out);
		if (progress_out) *progress_out = strdup(progress.str().c_str());
		if (match_log_out) *match_log_out = strdup(matches.str().c_str());
		mh.Clear();
		ml.Clear();
		return 0;
	} catch (gnException& e) {
		g_err = e.code.name + ": " + e.msg;
		return 1;
	} catch (const char* s) {
		g_err = s;
		return 2;
	}
}

// MemHash::WriteFile (MemHash.cpp:307-328) of the MemHash result: the .mems text
int ref_mems_write_file(int n_seqs, const char* const* seqs, const uint64_t* lens, uint64_t seed, char** text_out) {
	try {
		MatchList ml;
		build_list(ml, n_seqs, seqs, lens, seed);
		MemHash mh;
		mh.FindMatches(ml);
		std::ostringstream os;
		mh.WriteFile(os);
		*text_out = strdup(os.str().c_str());
		mh.Clear();
		ml.Clear();
		return 0;
	} catch (gnException& e) {
		g_err = e.code.name + ": " + e.msg;
		return 1;
	}
}

// MemHash::LoadFile (MemHash.cpp:266-305) into a MemHash that holds the sequences (AddHashEntry extends what it
// inserts, so the sequences must be there) -> its match list and counters
int ref_mems_load_file(int n_seqs, const char* const* seqs, const uint64_t* lens, uint64_t seed, const char* text,
                       int64_t** flat_out, uint64_t* n_flat_out, uint64_t* n_matches_out, uint64_t* counts_out) {
	try {
		MatchList ml;
		build_list(ml, n_seqs, seqs, lens, seed);
		MemHash mh;
		for (int g = 0; g < n_seqs; ++g) mh.AddSequence(ml.sml_table[g], ml.seq_table[g]);
		std::istringstream is(text);
		mh.LoadFile(is);
		if (counts_out) { counts_out[0] = mh.MemCount(); counts_out[1] = mh.MemCollisionCount(); }
		mh.GetMatchList(ml);
		flatten(ml, flat_out, n_flat_out, n_matches_out);
		mh.Clear();
		ml.Clear();
		return 0;
	} catch (gnException& e) {
		g_err = e.code.name + ": " + e.msg;
		return 1;
	} catch (const char* s) {
		g_err = s;
		return 2;
	}
}

// EliminateOverlaps (Aligner.cpp:62-180) on a match list given as flat records -> the list it leaves behind, in its order
int ref_eliminate_overlaps(const int64_t* flat_in, uint64_t n_flat_in, int64_t** flat_out, uint64_t* n_flat_out, uint64_t* n_matches_out) {
	try {
		MatchList ml;
		for (uint64_t i = 0; i < n_flat_in;) {
			const uint k = (uint)flat_in[i];
			Match m(k);
			Match* mm = m.Copy();
			mm->SetLength((gnSeqI)flat_in[i + 1]);
			for (uint s = 0; s < k; ++s) mm->SetStart(s, flat_in[i + 2 + s]);
			ml.push_back(mm);
			i += 2 + k;
		}
		EliminateOverlaps(ml);
		flatten(ml, flat_out, n_flat_out, n_matches_out);
		for (size_t i = 0; i < ml.size(); ++i) ml[i]->Free();
		return 0;
	} catch (gnException& e) {
		g_err = e.code.name + ": " + e.msg;
		return 1;
	} catch (const char* s) {
		g_err = s;
		return 2;
	}
}

// mask for mode 3 (MaskedMemHash::SetMask, MaskedMemHash.h:32)
void ref_set_seq_mask(uint64_t mask) { g_seq_mask = mask; }

void ref_free(void* p) { free(p); }

}  // extern "C"
