// libMems/SortedMerList.h façade — SortedMerList / MemorySML / DNAMemorySML over the C-ABI.
// Same member names and meaning as the reference (SortedMerList.h:69-282, MemorySML.h:27-55,
// DNAMemorySML.h:24-48); the sorted list itself lives in GPU memory behind a mems_sml_t handle.
#pragma once
#include <cstdint>
#include <cstring>
#include <fstream>
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
	// MemorySML::Clone / DNAMemorySML::Clone (MemorySML.cpp:36-38): a new object over the same device-resident list
	virtual SortedMerList* Clone() const {
		SortedMerList* c = make_empty();
		if (sml) {
			mems_sml_t h = nullptr;
			Context::check(mems_sml_clone(sml, &h));
			c->adopt(h);
		}
		return c;
	}
	void adopt(mems_sml_t h) {  // used by MatchList::CreateMemorySMLs (batch build)
		Clear();
		sml = h;
		Context::check(mems_sml_info(sml, &info));
	}

protected:
	virtual SortedMerList* make_empty() const { return new SortedMerList(); }
	uint64_t mer_at(uint64_t offset, bool dna) const {
		uint64_t f = 0, d = 0;
		Context::check(mems_sml_seed_mers(sml, &offset, 1, &f, &d));
		return dna ? d : f;
	}
	mems_sml_t sml;
	mems_sml_info_t info{};
};

typedef SortedMerList MemorySML;

// ---- the .sml file of FileSML / DNAFileSML (format version 5, DNAFileSML.h:58-62) -------------------------
// Layout written by FileSML::Create (FileSML.cpp:344-366): SMLHeader as the raw struct, the 2-bit packed
// sequence (SetSequence's words incl. the two zero pad words, SortedMerList.cpp:306-317), then the sorted
// positions as smlSeqI_t.  The struct below repeats SortedMerList.h:48-63 field by field (libGenome's
// `boolean` is one byte, `sarID_t` is int16), so its size and offsets are the reference's on this ABI.
constexpr int kSmlDescriptionSize = 2048;  // DESCRIPTION_SIZE, SortedMerList.h:34
struct SMLHeader {
	uint32_t version;
	uint32_t alphabet_bits;
	uint64_t seed;
	uint32_t seed_length;
	uint32_t seed_weight;
	uint64_t length;
	uint32_t unique_mers;
	uint32_t word_size;
	uint8_t little_endian;
	int16_t id;
	uint8_t circular;
	uint8_t translation_table[255];  // UINT8_MAX entries, as in the reference
	char description[kSmlDescriptionSize];
};

// BasicDNATable (SortedMerList.cpp:29-47): c,b,y -> 1; g,s,k -> 2; t -> 3; everything else 0; 255 entries
inline void FillBasicDNATable(uint8_t* t) {
	std::memset(t, 0, 255);
	t['c'] = 1; t['C'] = 1; t['b'] = 1; t['B'] = 1; t['y'] = 1; t['Y'] = 1;
	t['g'] = 2; t['G'] = 2; t['s'] = 2; t['S'] = 2; t['k'] = 2; t['K'] = 2;
	t['t'] = 3; t['T'] = 3;
}

// DNAFileSML: Create() builds the list on the GPU and writes the reference's file; LoadFile() reads such a
// file (the reference's or ours), rebuilds the list on the GPU from the packed sequence — a rebuild is
// faster than reading the positions — and checks it against the stored positions.
class DNAFileSML : public SortedMerList {
public:
	DNAFileSML() {}
	explicit DNAFileSML(const std::string& fname) : filename(fname) {}
	static uint32_t FormatVersion() { return 5; }
	uint64_t GetSeedMer(uint64_t offset) const override { return mer_at(offset, true); }  // DNAFileSML.cpp: GetDnaSeedMer

	template <class Sequence>
	void Create(const Sequence& seq, const uint64_t seed) {
		SortedMerList::Create(seq, seed);
		WriteFile(filename);
	}
	void WriteFile(const std::string& fname) const {
		SMLHeader h;
		std::memset(&h, 0, sizeof h);  // the reference leaves word_size, little_endian and the padding uninitialised
		h.version = FormatVersion();
		h.alphabet_bits = 2;
		h.seed = info.seed;
		h.seed_length = info.seed_length;
		h.seed_weight = info.seed_weight;
		h.length = info.length;
		h.unique_mers = 0xffffffffu;  // NO_UNIQUE_COUNT
		h.word_size = 32;
		h.little_endian = 1;
		h.id = 0;
		h.circular = 0;
		FillBasicDNATable(h.translation_table);
		uint64_t n_words = 0;
		Context::check(mems_sml_packed(sml, nullptr, &n_words));
		std::vector<uint32_t> words(n_words ? n_words : 1);
		Context::check(mems_sml_packed(sml, words.data(), &n_words));
		std::vector<uint32_t> pos(info.sml_length ? info.sml_length : 1);
		std::vector<uint64_t> mer(info.sml_length ? info.sml_length : 1);
		uint64_t n = 0;
		Context::check(mems_sml_read(sml, 0, info.sml_length, pos.data(), mer.data(), &n));
		std::ofstream f(fname.c_str(), std::ios::binary | std::ios::trunc);
		if (!f.is_open()) throw MemsException(MEMS_ERR_INVALID, "FileSML: unable to open " + fname);
		f.write(reinterpret_cast<const char*>(&h), sizeof h);
		f.write(reinterpret_cast<const char*>(words.data()), (std::streamsize)(n_words * sizeof(uint32_t)));
		f.write(reinterpret_cast<const char*>(pos.data()), (std::streamsize)(n * sizeof(smlSeqI_t)));
		if (!f.good()) throw MemsException(MEMS_ERR_INVALID, "FileSML: error writing " + fname);
	}
	// FileSML::LoadFile (FileSML.cpp:46-110): header checks as there; throws on a missing/short file or a
	// foreign format version
	void LoadFile(const std::string& fname) {
		std::ifstream f(fname.c_str(), std::ios::binary);
		if (!f.is_open()) throw MemsException(MEMS_ERR_INVALID, "FileSML: unable to open file");
		SMLHeader h;
		f.read(reinterpret_cast<char*>(&h), sizeof h);
		if (f.gcount() < (std::streamsize)sizeof h) throw MemsException(MEMS_ERR_INVALID, "FileSML: unable to read file");
		if (h.version != FormatVersion()) throw MemsException(MEMS_ERR_UNSUPPORTED, "FileSML: unsupported file format");
		if (h.circular) throw MemsException(MEMS_ERR_UNSUPPORTED, "circular sequences are not supported");
		if (h.alphabet_bits != 2) throw MemsException(MEMS_ERR_UNSUPPORTED, "FileSML: not a DNA list");
		// nothing is sized from the header before the header is plausible and the file is as long as it says
		if (h.length > 0xffffffffull) throw MemsException(MEMS_ERR_INVALID, "FileSML: corrupt header (sequence length)");
		if (h.seed_length < 1 || h.seed_length > 31 || h.seed == 0 ||
		    (uint32_t)mems_get_seed_length(h.seed) != h.seed_length || (uint32_t)mems_get_seed_weight(h.seed) != h.seed_weight)
			throw MemsException(MEMS_ERR_INVALID, "FileSML: corrupt header (seed pattern)");
		{
			const uint64_t want_words = (h.length * 2 + 31) / 32 + 2;
			const uint64_t want_pos = h.length >= h.seed_length ? h.length - h.seed_length + 1 : 0;
			f.seekg(0, std::ios::end);
			const uint64_t file_size = (uint64_t)f.tellg();
			f.seekg((std::streamoff)sizeof h, std::ios::beg);
			if (file_size < sizeof h + want_words * sizeof(uint32_t))
				throw MemsException(MEMS_ERR_INVALID, "FileSML: error reading sequence data");
			if (file_size < sizeof h + want_words * sizeof(uint32_t) + want_pos * sizeof(smlSeqI_t))
				throw MemsException(MEMS_ERR_INVALID, "FileSML: premature end of file");
		}
		const uint64_t n_words = (h.length * 2 + 31) / 32 + 2;
		std::vector<uint32_t> words(n_words);
		f.read(reinterpret_cast<char*>(words.data()), (std::streamsize)(n_words * sizeof(uint32_t)));
		if (f.gcount() < (std::streamsize)(n_words * sizeof(uint32_t)))
			throw MemsException(MEMS_ERR_INVALID, "FileSML: error reading sequence data");
		const uint64_t n_pos = h.length >= h.seed_length ? h.length - h.seed_length + 1 : 0;
		std::vector<uint32_t> stored(n_pos ? n_pos : 1);
		f.read(reinterpret_cast<char*>(stored.data()), (std::streamsize)(n_pos * sizeof(smlSeqI_t)));
		if (f.gcount() < (std::streamsize)(n_pos * sizeof(smlSeqI_t)))
			throw MemsException(MEMS_ERR_INVALID, "FileSML: premature end of file");
		// 2-bit words back to letters (A C G T), then the usual build
		std::string seq(h.length, 'A');
		static const char letters[4] = {'A', 'C', 'G', 'T'};
		for (uint64_t p = 0; p < h.length; ++p) seq[p] = letters[(words[p >> 4] >> (30 - 2 * (p & 15))) & 3u];
		Clear();
		Context::check(mems_sml_create(Context::get(), seq.data(), h.length, h.seed, &sml));
		Context::check(mems_sml_info(sml, &info));
		filename = fname;
		// the stored list must order the same mers the same way (ties inside an equal-mer run may differ: the
		// reference's std::sort leaves them unspecified)
		std::vector<uint32_t> pos(n_pos ? n_pos : 1);
		std::vector<uint64_t> mer(n_pos ? n_pos : 1), stored_mer(n_pos ? n_pos : 1);
		uint64_t n = 0;
		Context::check(mems_sml_read(sml, 0, n_pos, pos.data(), mer.data(), &n));
		std::vector<uint64_t> where(stored.begin(), stored.begin() + n_pos);
		if (n_pos) Context::check(mems_sml_seed_mers(sml, where.data(), n_pos, nullptr, stored_mer.data()));
		for (uint64_t i = 0; i < n_pos; ++i)
			if (stored_mer[i] != mer[i]) throw MemsException(MEMS_ERR_INVALID, "FileSML: the stored list is not sorted by this seed");
	}
	const std::string& FileName() const { return filename; }

protected:
	SortedMerList* make_empty() const override { return new DNAFileSML(filename); }
	std::string filename;
};

// DNAMemorySML: GetSeedMer is the canonical (strand-minimal) mer (DNAMemorySML.cpp:35-41)
class DNAMemorySML : public SortedMerList {
public:
	uint64_t GetSeedMer(uint64_t offset) const override { return mer_at(offset, true); }
	DNAMemorySML* Clone() const override { return static_cast<DNAMemorySML*>(SortedMerList::Clone()); }
protected:
	SortedMerList* make_empty() const override { return new DNAMemorySML(); }
};

}  // namespace mems
