// libMems/Match.h façade — ungapped multi-sequence match with the reference's accessors
// (AbstractMatch.h:87-206, UngappedLocalAlignment.h:28-85, HybridAbstractMatch.h:29-98): 1-based signed
// starts, negative = reverse strand, NO_MATCH = 0.
#pragma once
#include <cstdint>
#include <ostream>
#include <stdexcept>
#include <vector>

namespace mems {

static const int64_t NO_MATCH = 0;  // AbstractMatch.h:27

class Match {
public:
	enum orientation { forward, reverse, undefined };
	Match() : m_length(0) {}
	explicit Match(unsigned seq_count) : m_start(seq_count, NO_MATCH), m_length(0) {}
	Match* Copy() const { return new Match(*this); }
	void Free() { delete this; }
	unsigned SeqCount() const { return (unsigned)m_start.size(); }
	int64_t Start(unsigned seqI) const { return seqI < m_start.size() ? m_start[seqI] : NO_MATCH; }
	int64_t operator[](unsigned seqI) const { return Start(seqI); }
	void SetStart(unsigned seqI, int64_t s) { m_start[seqI] = s; }
	uint64_t Length(unsigned = 0) const { return m_length; }
	void SetLength(uint64_t len, unsigned = 0) { m_length = len; }
	uint64_t AlignmentLength() const { return m_length; }
	uint64_t LeftEnd(unsigned seqI) const { return (uint64_t)(Start(seqI) < 0 ? -Start(seqI) : Start(seqI)); }
	uint64_t RightEnd(unsigned seqI) const { return LeftEnd(seqI) + m_length - 1; }
	int64_t End(unsigned seqI) const { return Start(seqI) > 0 ? Start(seqI) + (int64_t)m_length - 1 : Start(seqI); }
	orientation Orientation(unsigned seqI) const { return Start(seqI) > 0 ? forward : (Start(seqI) < 0 ? reverse : undefined); }
	unsigned Multiplicity() const {
		unsigned m = 0;
		for (int64_t s : m_start) m += s != NO_MATCH;
		return m;
	}
	unsigned FirstStart() const {
		for (unsigned i = 0; i < m_start.size(); ++i)
			if (m_start[i] != NO_MATCH) return i;
		return (unsigned)-1;
	}
	void Invert() {
		for (int64_t& s : m_start) s = -s;
	}
	// HybridAbstractMatch::MoveStart / MoveEnd (HybridAbstractMatch.h:271-290): the alignment's first column sits at the
	// left end of forward members and at the right end of reverse ones, so moving the start touches forward starts only
	// and moving the end grows the magnitude of reverse starts
	void MoveStart(int64_t amount) {
		for (int64_t& s : m_start)
			if (s > 0) s += amount;
	}
	void MoveEnd(int64_t amount) {
		for (int64_t& s : m_start)
			if (s < 0) s -= amount;
	}
	// UngappedLocalAlignment::CropStart / CropEnd (UngappedLocalAlignment.h:138-152)
	void CropStart(uint64_t crop_amount) {
		if (crop_amount > m_length) throw std::out_of_range("SeqIndexOutOfBounds");
		m_length -= crop_amount;
		MoveStart((int64_t)crop_amount);
	}
	void CropEnd(uint64_t crop_amount) {
		if (crop_amount > m_length) throw std::out_of_range("SeqIndexOutOfBounds");
		m_length -= crop_amount;
		MoveEnd((int64_t)crop_amount);
	}
	bool operator==(const Match& o) const { return m_length == o.m_length && m_start == o.m_start; }

private:
	std::vector<int64_t> m_start;
	uint64_t m_length;
};

// UngappedLocalAlignment.h:201-206: "len \t start0 \t start1 ..."
inline std::ostream& operator<<(std::ostream& os, const Match& m) {
	os << m.Length();
	for (unsigned i = 0; i < m.SeqCount(); ++i) os << '\t' << m.Start(i);
	return os;
}

}  // namespace mems
