// libGenome/gnSequence.h — the sliver of genome::gnSequence that libMems' anchoring path touches
// (length / isCircular / ToArray, SortedMerList.cpp:801-811), for using this façade without libGenome.
// With the real libGenome on the include path this file is simply not picked up: the façade classes are
// templates over the sequence type and only call those three members.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>

namespace genome {
typedef uint64_t gnSeqI;
typedef char gnSeqC;
class gnSequence {
public:
	gnSequence() {}
	gnSequence(const std::string& s) : seq(s) {}
	gnSequence(const char* s, size_t n) : seq(s, n) {}
	virtual ~gnSequence() {}
	virtual gnSeqI length() const { return seq.size(); }
	virtual bool isCircular() const { return false; }
	virtual bool ToArray(gnSeqC* dest, gnSeqI len, gnSeqI offset = 1) const {  // 1-based offset like libGenome
		memcpy(dest, seq.data() + offset - 1, len);
		return true;
	}
	const char* data() const { return seq.data(); }
private:
	std::string seq;
};
}  // namespace genome
