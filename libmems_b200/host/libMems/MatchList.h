// libMems/MatchList.h façade — GenericMatchList<Match*> with seq_table / sml_table (MatchList.h:33-116)
// and CreateMemorySMLs (MatchList.h:408-435), which here builds all lists as ONE device batch.
#pragma once
#include <cmath>
#include <istream>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>

#include "libMems/Match.h"
#include "libMems/SortedMerList.h"

namespace mems {

template <class Sequence = genome::gnSequence>
class GenericMatchList : public std::vector<Match*> {
public:
	std::vector<std::string> sml_filename, seq_filename;
	std::vector<SortedMerList*> sml_table;
	std::vector<Sequence*> seq_table;

	static unsigned GetDefaultMerSize(const std::vector<Sequence*>& seq_table) {  // MatchList.h:352-357
		uint64_t total = 0;
		for (auto* s : seq_table) total += s->length();
		return getDefaultSeedWeight(seq_table.empty() ? 0 : total / seq_table.size());
	}
	void CreateMemorySMLs(unsigned mer_size, std::ostream* log_stream, int seed_rank = 0) {
		if (mer_size == 0) {
			mer_size = GetDefaultMerSize(seq_table);
			if (log_stream) (*log_stream) << "Using " << mer_size << "-mers for initial seeds\n";
		}
		const uint64_t seed = (uint64_t)getSeed((int)mer_size, seed_rank);
		const size_t n = seq_table.size();
		std::vector<std::vector<char>> bufs(n);
		std::vector<const char*> ptrs(n);
		std::vector<uint64_t> lens(n);
		for (size_t i = 0; i < n; ++i) {
			lens[i] = seq_table[i]->length();
			bufs[i].resize(lens[i] ? lens[i] : 1);
			if (lens[i]) seq_table[i]->ToArray(bufs[i].data(), lens[i]);
			ptrs[i] = bufs[i].data();
		}
		if (log_stream) (*log_stream) << "Creating sorted mer lists\n";
		std::vector<mems_sml_t> handles(n);
		Context::check(mems_sml_create_batch(Context::get(), (int)n, ptrs.data(), lens.data(), seed, handles.data()));
		for (size_t i = 0; i < n; ++i) {
			DNAMemorySML* sml = new DNAMemorySML();
			sml->adopt(handles[i]);
			sml_table.push_back(sml);
		}
	}
	// LoadSMLs (MatchList.h:262-349): one .sml file per sequence (sml_filename); a file that loads and carries the
	// wanted seed is used, any other is (re)created with DNAFileSML::Create.  `solid` has no effect, as in the
	// reference: it assigns getSolidSeed() to a shadowing local that dies at once (MatchList.h:275-276).
	void LoadSMLs(unsigned mer_size, std::ostream* log_stream, int seed_rank = 0, bool solid = false, bool force_create = false) {
		if (mer_size == 0) {
			mer_size = GetDefaultMerSize(seq_table);
			if (log_stream) (*log_stream) << "Using weight " << mer_size << " mers for initial seeds\n";
		}
		(void)solid;
		const uint64_t default_seed = (uint64_t)getSeed((int)mer_size, seed_rank);
		for (size_t i = 0; i < seq_table.size(); ++i) {
			DNAFileSML* file_sml = new DNAFileSML(sml_filename[i]);
			sml_table.push_back(file_sml);
			bool usable = !force_create;
			if (usable) {
				try {
					file_sml->LoadFile(sml_filename[i]);
				} catch (const MemsException&) {
					usable = false;
				}
			}
			if (usable && file_sml->Seed() != default_seed) {
				if (log_stream) (*log_stream) << "Default seed mismatch.  A new sorted mer list will be created.\n";
				usable = false;
			}
			if (usable) {
				if (log_stream) (*log_stream) << "Sorted mer list loaded successfully\n";
				continue;
			}
			if (log_stream) (*log_stream) << "Creating sorted mer list\n";
			file_sml->Create(*seq_table[i], default_seed);
		}
	}
	void Clear() {  // MatchList.h:438-457
		for (auto* s : seq_table) delete s;
		for (auto* s : sml_table) delete s;
		for (auto* m : *this) m->Free();
		seq_table.clear();
		sml_table.clear();
		this->clear();
		seq_filename.clear();
		sml_filename.clear();
	}
	void MultiplicityFilter(unsigned mult) {  // MatchList.h:637-650
		size_t cur = 0;
		for (size_t i = 0; i < this->size(); ++i) {
			if ((*this)[i]->Multiplicity() == mult) (*this)[cur++] = (*this)[i];
			else (*this)[i]->Free();
		}
		this->resize(cur);
	}
	void LengthFilter(uint64_t length) {  // MatchList.h:652-664
		size_t cur = 0;
		for (size_t i = 0; i < this->size(); ++i) {
			if ((*this)[i]->Length() >= length) (*this)[cur++] = (*this)[i];
			else (*this)[i]->Free();
		}
		this->resize(cur);
	}
};

typedef GenericMatchList<> MatchList;

// .mums match files, format version 3 (ReadList / WriteList, MatchList.h:498-634).  The reference writes each
// match's heap address as its id; any unique number round-trips, so the match's index + 1 is written here.
template <class MatchListType>
void WriteList(const MatchListType& mlist, std::ostream& match_file) {
	if (mlist.size() == 0) return;
	const unsigned seq_count = (*mlist.begin())->SeqCount();
	match_file << "FormatVersion" << '\t' << 3 << "\n";
	match_file << "SequenceCount" << '\t' << seq_count << "\n";
	for (unsigned seqI = 0; seqI < seq_count; seqI++) {
		match_file << "Sequence" << seqI << "File" << '\t';
		if (mlist.seq_filename.size() > seqI) match_file << mlist.seq_filename[seqI];
		else match_file << "null";
		match_file << "\n";
		match_file << "Sequence" << seqI << "Length" << '\t';
		if (mlist.seq_table.size() > seqI) match_file << mlist.seq_table[seqI]->length();
		else match_file << "0";
		match_file << "\n";
	}
	match_file << "MatchCount" << '\t' << mlist.size() << std::endl;
	uint64_t id = 0;
	for (const Match* m : mlist) match_file << *m << '\t' << ++id << '\t' << 0 << '\t' << 0 << std::endl;
}

template <class MatchListType>
void ReadList(MatchListType& mlist, std::istream& match_file) {
	std::string tag;
	unsigned seq_count = 0;
	match_file >> tag;
	if (tag != "FormatVersion") throw MemsException(MEMS_ERR_INVALID, "InvalidFileFormat");
	match_file >> tag;
	if (tag != "3") throw MemsException(MEMS_ERR_INVALID, "InvalidFileFormat");
	match_file >> tag;
	if (tag != "SequenceCount") throw MemsException(MEMS_ERR_INVALID, "InvalidFileFormat");
	match_file >> seq_count;
	if (seq_count < 2) throw MemsException(MEMS_ERR_INVALID, "InvalidFileFormat");
	for (unsigned seqI = 0; seqI < seq_count; seqI++) {
		match_file >> tag;  // name tag
		std::getline(match_file, tag);
		mlist.seq_filename.push_back(tag.empty() ? tag : tag.substr(1));  // skip the tab
		uint64_t seq_len;
		match_file >> tag >> seq_len;
	}
	uint64_t match_count = 0;
	match_file >> tag >> match_count;
	std::string line;
	std::getline(match_file, line);
	while (std::getline(match_file, line)) {
		if (line.empty()) continue;
		std::istringstream ls(line);
		uint64_t len;
		ls >> len;
		Match* m = new Match(seq_count);
		m->SetLength(len);
		for (unsigned seqI = 0; seqI < seq_count; seqI++) {
			int64_t start;
			ls >> start;
			m->SetStart(seqI, start);
		}
		mlist.push_back(m);
	}
	if (match_count != mlist.size()) throw MemsException(MEMS_ERR_INVALID, "InvalidFileFormat");
}

}  // namespace mems
