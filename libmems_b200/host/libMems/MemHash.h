// libMems/MemHash.h façade — MatchFinder / MemHash / RepeatHash over the C-ABI (MatchFinder.h:46-118,
// MemHash.h:38-175, RepeatHash.h:24-46).  FindMatches(MatchList&) runs equal-seed run detection, the MUM /
// repeat policy and ungapped extension on the GPU and appends Match objects to the list.
#pragma once
#include <algorithm>
#include <istream>
#include <ostream>
#include <sstream>
#include <string>
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
	void FindMatches(MatchListType& ml) {  // MemHash.cpp:109-115
		FindMatchesFromPosition(ml, std::vector<uint64_t>(ml.seq_table.size(), 0));
	}
	// MemHash::FindMatchesFromPosition (MemHash.cpp:117-127): sorted mer list g is searched from entry start_points[g] on
	template <class MatchListType>
	void FindMatchesFromPosition(MatchListType& ml, const std::vector<uint64_t>& start_points) {
		for (size_t i = 0; i < ml.seq_table.size(); ++i) AddSequence(ml.sml_table[i]);
		if (start_points.size() != sar_table.size())
			throw MemsException(MEMS_ERR_INVALID, "Inconsistent search range specification.");  // MatchFinder.cpp:201-204
		m_start_points = start_points;
		try {
			CreateMatches();
		} catch (...) {
			m_start_points.clear();
			throw;
		}
		m_start_points.clear();
		GetMatchList(ml);
	}
	// MatchFinder::LogProgress (MatchFinder.cpp:298-309): the reference prints "<p>%.." whenever the share of processed
	// mers passes a whole percent, and a line break every ten.  The device call has no intermediate states, so the same
	// text is written in one piece when the search returns.
	virtual void LogProgress(std::ostream* os) { log_stream = os; }
	// MatchFinder::SetOffsetLog (MatchFinder.h:81, MatchFinder.cpp:149-163): the reference writes the per-list offsets each
	// time SearchRange gives up on a seed that occurs more than MER_REPEAT_LIMIT times and restarts; the device search
	// never restarts (such seeds are skipped in place), so nothing is ever written to it.
	void SetOffsetLog(std::ostream* os) { offset_stream = os; }
	// MemHash::SetMatchLog (MemHash.cpp:237-241): every match that enters the table is written as "len\tstart0\t..." —
	// here after the search, in table order rather than in order of discovery.
	virtual void SetMatchLog(std::ostream* os) { match_log = os; }
	// MemHash::WriteFile (MemHash.cpp:307-328): the .mems header and every stored match in table order
	virtual void WriteFile(std::ostream& mem_file) const {
		mem_file << "FormatVersion" << '\t' << 1 << "\n";
		mem_file << "SequenceCount" << '\t' << sar_table.size() << "\n";
		for (size_t seqI = 0; seqI < sar_table.size(); ++seqI) {
			mem_file << "Sequence" << seqI << "File" << '\t' << "null" << "\n";  // (sequences carry no source name here)
			mem_file << "Sequence" << seqI << "Length" << '\t' << sar_table[seqI]->Length() << "\n";
		}
		mem_file << "MatchCount" << '\t' << m_mem_count << std::endl;
		for (const Match* m : stored) mem_file << *m << "\n";
	}
	// MemHash::LoadFile (MemHash.cpp:266-305): bare "len start0 start1 ..." lines (NOT what WriteFile writes: the
	// reference's reader has no header parsing either); every line goes through AddHashEntry, so matches contained in
	// an earlier one on its diagonal are dropped and counted as collisions
	virtual void LoadFile(std::istream& mem_file) {
		if (mode() == MEMS_MODE_REPEAT) throw MemsException(MEMS_ERR_UNSUPPORTED, "LoadFile needs the persistent table (not RepeatHash)");
		if (!table) Context::check(mems_table_create(table_size, &table));
		std::string tag;
		std::getline(mem_file, tag);
		std::stringstream first_mum(tag);
		uint64_t len = 0;
		int64_t start = 0;
		std::vector<int64_t> starts;
		first_mum >> len;
		while (first_mum >> start) starts.push_back(start);
		if (!starts.empty()) {
			const uint32_t seq_count = (uint32_t)starts.size();
			// The reference rewinds the first line and reads its seq_count starts from the BEGINNING of the line
			// (MemHash.cpp:280-285), so the first match gets the length as start 0 and loses its last start: kept,
			// files written for the reference load the same way here.
			starts.insert(starts.begin(), (int64_t)len);
			starts.pop_back();
			Context::check(mems_table_add(table, seq_count, len, starts.data(), m_mersize, nullptr));
			while (mem_file.good()) {
				mem_file >> len;
				if (!mem_file.good()) break;
				for (uint32_t seqI = 0; seqI < seq_count; ++seqI) mem_file >> starts[seqI];
				if (!mem_file.good()) break;  // (like the reference: a last line without a line end is not taken)
				Context::check(mems_table_add(table, seq_count, len, starts.data(), m_mersize, nullptr));
			}
		}
		mems_matches_t m = nullptr;
		Context::check(mems_table_matches(table, &m));
		take_result(m, true);
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
		p.start_points = m_start_points.empty() ? nullptr : m_start_points.data();
		if (order == MEMS_ORDER_REFERENCE && mode() != MEMS_MODE_REPEAT) {
			// the reference's table lives across FindMatches calls (several seed patterns accumulate in it)
			if (!table) Context::check(mems_table_create(table_size, &table));
			p.table = table;
		}
		mems_matches_t m = nullptr;
		Context::check(mems_find_matches(Context::get(), (int)h.size(), h.data(), &p, &m));
		if (!sar_table.empty()) m_mersize = sar_table[0]->SeedWeight();  // MatchFinder.cpp:188
		const size_t before = p.table ? 0 : stored.size();
		take_result(m, p.table != nullptr);
		if (log_stream) write_progress();
		if (match_log)
			for (size_t i = before; i < stored.size(); ++i) (*match_log) << *stored[i] << std::endl;
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
		m_mersize = 31;  // DNA_MER_SIZE until a search sets the seed weight (MatchFinder.cpp:48, :188)
		order = MEMS_ORDER_ANY;
	}
	// copies a match list out of the library; whole_table: it replaces what was stored and the counters are the table's
	void take_result(mems_matches_t m, bool whole_table) {
		mems_matches_info_t info;
		Context::check(mems_matches_info(m, &info));
		const int64_t* flat = mems_matches_data(m);
		if (whole_table) {
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
	}
	// The text MatchFinder.cpp:298-309 produces over a whole search.  The reference reads every sorted mer list in
	// buffers of MER_BUFFER_SIZE = 10000 entries and updates the progress each time a buffer is used up, i.e. when the
	// merge passes the mer of the buffer's last entry: the same events are replayed here in order of that mer.
	void write_progress() {
		struct Event {
			uint64_t mer, size;
			size_t sml;
		};
		std::vector<Event> events;
		double total_mers = 0, mers_processed = 0;
		for (size_t i = 0; i < sar_table.size(); ++i) {
			SortedMerList* s = sar_table[i];
			const uint64_t start = m_start_points.empty() ? 0 : m_start_points[i];
			total_mers += (double)s->Length();  // MatchFinder.cpp:146: bases, while the processed mers are list entries
			mers_processed += (double)start;
			for (uint64_t at = start; at < s->SMLLength(); at += 10000) {
				const uint64_t size = std::min<uint64_t>(10000, s->SMLLength() - at);
				events.push_back({(*s)[at + size - 1].mer & s->GetSeedMask(), size, i});
			}
		}
		if (total_mers == 0) return;
		// (lists that reach the same mer: the reference's merge list takes the one that arrived last first; with the common
		// case — all lists end on the same largest mer — that is the higher list index)
		std::stable_sort(events.begin(), events.end(), [](const Event& a, const Event& b) { return a.mer < b.mer || (a.mer == b.mer && a.sml > b.sml); });
		float progress = -1;
		for (const Event& e : events) {
			mers_processed += (double)e.size;
			const double old = progress;
			progress = (float)((mers_processed / total_mers) * 100.0);
			if ((int)old != (int)progress) (*log_stream) << (int)((progress / 100.0f) * 100) << "%..";
			if (((int)old / 10) != ((int)progress / 10)) (*log_stream) << std::endl;
		}
		log_stream->flush();
	}
	void free_stored() {
		for (Match* m : stored) m->Free();
		stored.clear();
	}
	std::vector<SortedMerList*> sar_table;
	std::vector<Match*> stored;
	uint32_t table_size, m_repeat_tolerance, m_enumeration_tolerance;
	uint64_t m_mem_count, m_collision_count;
	uint32_t m_mersize;
	int order;
	mems_table_t table;
	std::vector<uint64_t> m_start_points;
	std::ostream* log_stream = nullptr;
	std::ostream* offset_stream = nullptr;
	std::ostream* match_log = nullptr;
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
