// oracle/shim: stand-in for libGenome/gnSequence.h — an in-memory linear DNA string with the
// handful of members the libMems hot path calls (SortedMerList.cpp:801-811, MatchList.h).
// Test infrastructure.
#pragma once
#include "libGenome/gnDefs.h"
#include "libGenome/gnClone.h"
#include "libGenome/gnException.h"
#include "libGenome/gnDebug.h"
#include <string>
#include <vector>
namespace genome {
template <class T> class Array {
public:
	explicit Array(uint64 n) { data = new T[n]; }
	~Array() { delete[] data; }
	T* data;
private:
	Array(const Array&);
	Array& operator=(const Array&);
};
class gnBaseFeature;
class gnBaseHeader;
class gnLocation;
class gnGenomeSpec {
public:
	std::string GetName() { return ""; }
};
class gnSequence : public gnClone {
public:
	gnSequence() : circ(false) {}
	gnSequence(const std::string& s) : seq(s), circ(false) {}
	gnSequence(const char* s, size_t n) : seq(s, n), circ(false) {}
	virtual ~gnSequence() {}
	virtual gnSequence* Clone() const { return new gnSequence(*this); }
	virtual gnSeqI length() const { return seq.size(); }
	virtual boolean isCircular() const { return circ; }
	virtual void setCircular(const boolean v) { circ = v; }
	// 1-based offset, like libGenome
	virtual boolean ToArray(gnSeqC* p, gnSeqI len, const gnSeqI offset = 1) const {
		memcpy(p, seq.data() + offset - 1, len);
		return true;
	}
	virtual std::string ToString(const gnSeqI len = GNSEQI_END, const gnSeqI offset = 1) const {
		return seq.substr(offset - 1, len == GNSEQI_END ? std::string::npos : len);
	}
	virtual bool LoadSource(const std::string) { return false; }
	gnSeqI contigListSize() const { return 1; }
	gnSequence contig(uint32) const { return *this; }
	gnGenomeSpec* GetSpec() { static gnGenomeSpec s; return &s; }
	std::string seq;
	boolean circ;
};
}
