// libMems/SortedMerList.h façade — SortedMerList / MemorySML / DNAMemorySML over the C-ABI.
// Same member names and meaning as the reference (SortedMerList.h:69-282, MemorySML.h:27-55,
// DNAMemorySML.h:24-48); the sorted list itself lives in GPU memory behind a mems_sml_t handle.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "libGenome/gnSequence.h"
#include "libMems/SeedMasks.h"
#include "mems_b200.h"

namespace mems {

typedef uint32_t smlSeqI_t;  // SortedMerList.h:40
struct bmer {                // SortedMerList.h:43-46
	smlSeqI_t position;
	uint64_t mer;
};

// the reference throws genome::gnException(SMLCreateError / InvalidData ...); here one exception type carries
// the C-ABI code and message
class MemsException : public std::runtime_error {
public:
	MemsException(int c, const std::string& m) : std::runtime_error(m), code(c) {}
	int code;
};

// One execution context (stream + scratch) per thread, like the reference's TLS<MemHash> (Aligner.h:198).
class Context {
public:
	static mems_ctx_t get() {
		static thread_local Context c;
		return c.h;
	}
	static void check(int rc) {
		if (rc != MEMS_OK) throw MemsException(rc, mems_last_error(holder()));
	}
private:
	static mems_ctx_t& holder() {
		static thread_local mems_ctx_t p = nullptr;
		return p;
	}
	Context() : h(nullptr) {
		int dev = 0;
		if (const char* e = getenv("MEMS_DEVICE")) dev = atoi(e);
		int rc = mems_ctx_create(dev, nullptr, &h);
		if (rc != MEMS_OK) throw MemsException(rc, mems_last_error(nullptr));
		holder() = h;
	}
	~Context() {
		mems_ctx_destroy(h);
		holder() = nullptr;
	}
	mems_ctx_t h;
};

class SortedMerList {
public:
	SortedMerList() : sml(nullptr) {}
	virtual ~SortedMerList() { Clear(); }
	SortedMerList(const SortedMerList&) = delete;
	SortedMerList& operator=(const SortedMerList&) = delete;

	virtual void Clear() {
		if (sml) mems_sml_destroy(sml);
		sml = nullptr;
	}
	// SortedMerList::Create / MemorySML::Create (SortedMerList.cpp:786-824, MemorySML.cpp:45-60)
	template <class Sequence>
	void Create(const Sequence& seq, const uint64_t seed) {
		if (seq.isCircular()) throw MemsException(MEMS_ERR_UNSUPPORTED, "circular sequences are not supported");
		std::vector<char> buf(seq.length() ? seq.length() : 1);
		if (seq.length()) seq.ToArray(buf.data(), seq.length());
		Clear();
		Context::check(mems_sml_create(Context::get(), buf.data(), seq.length(), seed, &sml));
		Context::check(mems_sml_info(sml, &info));
	}
	// MemorySML::Read (MemorySML.cpp:62-82)
	virtual bool Read(std::vector<bmer>& readVector, uint64_t size, uint64_t offset) {
		readVector.clear();
		if (offset > info.sml_length) return false;
		std::vector<uint32_t> pos(size ? size : 1);
		std::vector<uint64_t> mer(size ? size : 1);
		uint64_t n = 0;
		Context::check(mems_sml_read(sml, offset, size, pos.data(), mer.data(), &n));
		readVector.resize(n);
		for (uint64_t i = 0; i < n; ++i) readVector[i] = bmer{pos[i], mer[i]};
		return offset + size <= info.sml_length;
	}
	virtual bmer operator[](uint64_t index) {  // MemorySML.cpp:88-94
		bmer b;
		uint64_t n = 0;
		Context::check(mems_sml_read(sml, index, 1, &b.position, &b.mer, &n));
		return b;
	}
	virtual uint64_t GetSeedMer(uint64_t offset) const { return mer_at(offset, false); }  // forward (SortedMerList.cpp:726)
	uint64_t GetDnaSeedMer(uint64_t offset) const { return mer_at(offset, true); }        // canonical (:764)
	virtual bool FindMer(const uint64_t query_mer, uint64_t& result) {                     // SortedMerList.cpp:170-179
		int found = 0;
		Context::check(mems_sml_find_mer(sml, query_mer, &found, &result));
		return found != 0;
	}
	uint64_t Seed() const { return info.seed; }
	uint32_t SeedLength() const { return info.seed_length; }
	uint32_t SeedWeight() const { return info.seed_weight; }
	uint64_t GetSeedMask() const { return info.seed_mask; }
	uint64_t GetMerMask() const { return info.mer_mask; }
	uint64_t Length() const { return info.length; }
	uint64_t SMLLength() const { return info.sml_length; }
	bool IsCircular() const { return false; }
	mems_sml_t handle() const { return sml; }
	void adopt(mems_sml_t h) {  // used by MatchList::CreateMemorySMLs (batch build)
		Clear();
		sml = h;
		Context::check(mems_sml_info(sml, &info));
	}

protected:
	uint64_t mer_at(uint64_t offset, bool dna) const {
		uint64_t f = 0, d = 0;
		Context::check(mems_sml_seed_mers(sml, &offset, 1, &f, &d));
		return dna ? d : f;
	}
	mems_sml_t sml;
	mems_sml_info_t info{};
};

typedef SortedMerList MemorySML;

// DNAMemorySML: GetSeedMer is the canonical (strand-minimal) mer (DNAMemorySML.cpp:35-41)
class DNAMemorySML : public SortedMerList {
public:
	uint64_t GetSeedMer(uint64_t offset) const override { return mer_at(offset, true); }
};

}  // namespace mems
