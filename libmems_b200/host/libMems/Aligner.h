// libMems/Aligner.h façade — the one step of Aligner that sits directly behind the anchoring path: EliminateOverlaps
// (Aligner.cpp:58-180), which the callers run on the MatchList right after MemHash::FindMatches
// (Aligner.cpp:1127-1130, ProgressiveAligner.cpp:659-665).  Host code, on purpose.
//
// Why not a GPU sort-and-sweep: the function is a sequential greedy pass whose result depends on its own mutation
// order.  Per sequence it sorts the list by left end with std::sort — unstable, and matches of different sequence
// sets routinely share a left end, so the processing order of ties is whatever libstdc++'s introsort leaves — then
// walks it once, cropping or deleting the "smaller" of two overlapping matches IN PLACE: a cropped match is compared
// again with its next neighbours using its new length, the pieces cut off are appended as new matches that the
// passes for the following sequences see, and a deleted match ends the inner scan.  ("This code isn't perfect, it can
// delete too many base pairs in some cases", Aligner.cpp:60.)  A data-parallel formulation would have to fix one
// tie order and one cropping order and would then differ from the reference on exactly the inputs where those
// matter; being a drop-in means the same list, so the pass runs on the host with the same std::sort on the same
// container in the same order.  It is linear in the list after the sort (10^5 matches: milliseconds), far below the
// cost of the search that produced the list.
#pragma once
#include <algorithm>
#include <vector>

#include "libMems/MatchList.h"

namespace mems {

// SingleStartComparator (AbstractMatch.h:324-350): by left end in one sequence, undefined before defined
struct SingleStartComparator {
	unsigned seq;
	explicit SingleStartComparator(unsigned s = 0) : seq(s) {}
	bool operator()(const Match* a, const Match* b) const {
		const int64_t a_start = (int64_t)a->LeftEnd(seq), b_start = (int64_t)b->LeftEnd(seq);
		if (a_start == NO_MATCH || b_start == NO_MATCH) return b_start != NO_MATCH;
		return a_start < b_start;
	}
};

// EliminateOverlaps (Aligner.cpp:62-180): per sequence, left to right, of two matches that overlap there the one with
// fewer sequences (or, at equal multiplicity, the shorter) loses the overlapping columns; what it loses lives on as a
// new match without that sequence if at least two sequences remain.
template <class MatchListType>
void EliminateOverlaps(MatchListType& ml) {
	if (ml.size() < 2) return;
	const unsigned seq_count = ml[0]->SeqCount();
	for (unsigned seqI = 0; seqI < seq_count; ++seqI) {
		std::sort(ml.begin(), ml.end(), SingleStartComparator(seqI));
		std::vector<Match*> cut_off;  // the columns the losers gave up, as matches of the remaining sequences
		size_t deleted = 0;
		int64_t cur = 0;
		const int64_t n = (int64_t)ml.size();
		while (cur != n && ml[cur]->Start(seqI) == NO_MATCH) ++cur;  // undefined matches sort first
		for (; cur < n; ++cur) {
			if (ml[cur] == nullptr) continue;
			for (int64_t nxt = cur + 1; nxt < n; ++nxt) {
				if (ml[nxt] == nullptr) continue;
				const int64_t start_cur = ml[cur]->Start(seqI), len_cur = (int64_t)ml[cur]->Length();
				const int64_t start_nxt = ml[nxt]->Start(seqI);
				int64_t overlap = (start_cur < 0 ? -start_cur : start_cur) + len_cur - (start_nxt < 0 ? -start_nxt : start_nxt);
				if (overlap <= 0) break;  // sorted by left end: nothing further overlaps either
				const bool cur_loses = ml[nxt]->Multiplicity() > ml[cur]->Multiplicity() ||
				                       (ml[nxt]->Multiplicity() == ml[cur]->Multiplicity() && ml[nxt]->Length() > ml[cur]->Length());
				Match* loser = cur_loses ? ml[cur] : ml[nxt];
				Match* piece = loser->Copy();
				bool cur_gone = false;
				if (overlap >= (int64_t)loser->Length()) {  // swallowed whole
					loser->Free();
					++deleted;
					if (cur_loses) {
						ml[cur] = nullptr;
						--cur;  // (as the reference does; the outer loop moves on from here)
						cur_gone = true;
					} else {
						ml[nxt] = nullptr;
					}
				} else if (cur_loses) {  // the current match ends earlier in this sequence: it loses its right part
					if (start_cur > 0) {
						loser->CropEnd((uint64_t)overlap);
						piece->CropStart(piece->Length() - (uint64_t)overlap);
					} else {
						loser->CropStart((uint64_t)overlap);
						piece->CropEnd(piece->Length() - (uint64_t)overlap);
					}
				} else {  // the next match starts later: it loses its left part
					if (start_nxt > 0) {
						loser->CropStart((uint64_t)overlap);
						piece->CropEnd(piece->Length() - (uint64_t)overlap);
					} else {
						loser->CropEnd((uint64_t)overlap);
						piece->CropStart(piece->Length() - (uint64_t)overlap);
					}
				}
				piece->SetStart(seqI, 0);
				if (piece->Multiplicity() > 1 && piece->Length() > 0) cut_off.push_back(piece);
				else piece->Free();
				if (cur_gone) break;
			}
		}
		if (deleted > 0) {
			std::vector<Match*> kept;
			kept.reserve(ml.size() - deleted);
			for (Match* m : ml)
				if (m != nullptr) kept.push_back(m);
			ml.clear();
			ml.insert(ml.end(), kept.begin(), kept.end());
		}
		ml.insert(ml.end(), cut_off.begin(), cut_off.end());
	}
}

}  // namespace mems
