// libMems/MemHash.h façade — MatchFinder / MemHash / RepeatHash over the C-ABI (MatchFinder.h:46-118,
// MemHash.h:38-175, RepeatHash.h:24-46).  FindMatches(MatchList&) runs equal-seed run detection, the MUM /
// repeat policy and ungapped extension on the GPU and appends Match objects to the list.
#pragma once
#include <ostream>
#include <vector>

#include "libMems/MatchList.h"

namespace mems {

static const uint32_t DEFAULT_MEM_TABLE_SIZE = 40000;  // MemHash.h:30-32
static const uint32_t DEFAULT_REPEAT_TOLERANCE = 0;
static const uint32_t DEFAULT_ENUMERATION_TOLERANCE = 1;

class MemHash {
public:
	MemHash() : table(nullptr) { reset(); }
	virtual ~MemHash() {
		free_stored();
		if (table) mems_table_destroy(table);
	}
	MemHash(const MemHash&) = delete;
	MemHash& operator=(const MemHash&) = delete;
	virtual void Clear() {  // MemHash.cpp:76-93: also drops the stored matches
		ClearSequences();
		free_stored();
		if (table) mems_table_clear(table);
		reset();
	}
	virtual void ClearSequences() { sar_table.clear(); }  // keeps the table: further FindMatches calls accumulate

	bool AddSequence(SortedMerList* sar) {  // MatchFinder.cpp:59-87
		if (sar == nullptr) throw MemsException(MEMS_ERR_INVALID, "Null SortedMerList pointer");
		sar_table.push_back(sar);
		return true;
	}
	template <class MatchListType>
	void FindMatches(MatchListType& ml) {  // MemHash.cpp:109-127
		for (size_t i = 0; i < ml.seq_table.size(); ++i) AddSequence(ml.sml_table[i]);
		CreateMatches();
		GetMatchList(ml);
	}
	virtual bool CreateMatches() {
		if (m_repeat_tolerance != DEFAULT_REPEAT_TOLERANCE || m_enumeration_tolerance != DEFAULT_ENUMERATION_TOLERANCE)
			throw MemsException(MEMS_ERR_UNSUPPORTED, "only repeat_tolerance 0 / enumeration_tolerance 1 (MUMs) are supported");
		std::vector<mems_sml_t> h;
		for (auto* s : sar_table) h.push_back(s->handle());
		mems_match_params_t p{};
		p.mode = mode();
		p.order = order;
		p.table_size = table_size;
		p.seq_mask = seq_mask();
		if (order == MEMS_ORDER_REFERENCE && mode() != MEMS_MODE_REPEAT) {
			// the reference's table lives across FindMatches calls (several seed patterns accumulate in it)
			if (!table) Context::check(mems_table_create(table_size, &table));
			p.table = table;
		}
		mems_matches_t m = nullptr;
		Context::check(mems_find_matches(Context::get(), (int)h.size(), h.data(), &p, &m));
		mems_matches_info_t info;
		Context::check(mems_matches_info(m, &info));
		const int64_t* flat = mems_matches_data(m);
		if (p.table) {  // the result is the whole table: it replaces what was stored, counters are cumulative
			free_stored();
			m_mem_count = m_collision_count = 0;
		}
		for (uint64_t i = 0; i < info.n_flat;) {
			const unsigned k = (unsigned)flat[i];
			Match* mm = new Match(k);
			mm->SetLength((uint64_t)flat[i + 1]);
			for (unsigned s = 0; s < k; ++s) mm->SetStart(s, flat[i + 2 + s]);
			stored.push_back(mm);
			i += 2 + k;
		}
		m_mem_count += info.mem_count;
		m_collision_count += info.collisions;
		mems_matches_destroy(m);
		return true;
	}
	template <class MatchListType>
	void GetMatchList(MatchListType& mem_list) const {  // MemHash.h:183-203: copies, caller Free()s them
		mem_list.clear();
		for (const Match* m : stored) mem_list.push_back(m->Copy());
	}
	virtual uint32_t TableSize() const { return table_size; }
	virtual void SetTableSize(uint32_t n) { table_size = n; }
	virtual uint32_t MemCount() { return (uint32_t)m_mem_count; }
	virtual uint32_t MemCollisionCount() { return (uint32_t)m_collision_count; }
	virtual void SetRepeatTolerance(uint32_t t) { m_repeat_tolerance = t; }
	virtual uint32_t GetRepeatTolerance() const { return m_repeat_tolerance; }
	virtual void SetEnumerationTolerance(uint32_t t) { m_enumeration_tolerance = t; }
	virtual uint32_t GetEnumerationTolerance() const { return m_enumeration_tolerance; }
	// MEMS_ORDER_REFERENCE reproduces the reference's bucket order and collision counters exactly
	void SetOutputOrder(int o) { order = o; }

protected:
	virtual int mode() const { return MEMS_MODE_MEMHASH; }
	virtual uint64_t seq_mask() const { return 0; }
	void reset() {
		table_size = DEFAULT_MEM_TABLE_SIZE;
		m_repeat_tolerance = DEFAULT_REPEAT_TOLERANCE;
		m_enumeration_tolerance = DEFAULT_ENUMERATION_TOLERANCE;
		m_mem_count = m_collision_count = 0;
		order = MEMS_ORDER_ANY;
	}
	void free_stored() {
		for (Match* m : stored) m->Free();
		stored.clear();
	}
	std::vector<SortedMerList*> sar_table;
	std::vector<Match*> stored;
	uint32_t table_size, m_repeat_tolerance, m_enumeration_tolerance;
	uint64_t m_mem_count, m_collision_count;
	int order;
	mems_table_t table;
};

class MaskedMemHash : public MemHash {  // MaskedMemHash.h:21-40
public:
	virtual void SetMask(uint64_t m) { mask = m; }
protected:
	uint64_t seq_mask() const override { return mask; }
	uint64_t mask = 0;
};

class PairwiseMatchFinder : public MemHash {  // PairwiseMatchFinder.h:24-40
protected:
	int mode() const override { return MEMS_MODE_PAIRWISE; }
};

class RepeatHash : public MemHash {  // RepeatHash.h:24-46
protected:
	int mode() const override { return MEMS_MODE_REPEAT; }
};

}  // namespace mems
