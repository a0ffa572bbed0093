// facade_demo.cpp — the anchoring call sequence of the reference's callers (MatchList::CreateMemorySMLs +
// MemHash::FindMatches, e.g. ProgressiveAligner.cpp:636-653) written against the façade headers.
// Usage: facade_demo <memhash|repeat|pairwise|mums|smlmemhash> <seed_weight> <raw-sequence-file>...
//   prints "len\tstart0\tstart1..." lines ("mums": MemHash, then the .mums file WriteList produces, re-read with ReadList)
//        facade_demo readmums <file.mums>     parses a .mums file and prints its matches
//        facade_demo writesml <seed_weight> <raw-sequence-file> <out.sml>   DNAFileSML::Create
//        facade_demo loadsml <file.sml>       DNAFileSML::LoadFile, prints seed, lengths and the first entries
//        facade_demo smllayout                sizeof / offsets of the façade's SMLHeader
//        facade_demo frompos <seed_weight> <start0,start1,...> <raw-sequence-file>...   MemHash::FindMatchesFromPosition with
//                             LogProgress and SetMatchLog attached: matches on stdout, then "#progress" + text, "#matchlog" + text
//        facade_demo memsfile <seed_weight> <raw-sequence-file>...   MemHash::WriteFile of the search on stdout
//        facade_demo loadmems <file with bare match lines>            MemHash::LoadFile, prints the table + counters
//        facade_demo clone <seed_weight> <raw-sequence-file>          SortedMerList::Clone: the clone outlives the original
//        facade_demo overlaps <file with bare match lines>            EliminateOverlaps (Aligner.cpp:62-180), prints the list it leaves
#include <cstddef>
#include <fstream>
#include <iostream>
#include <iterator>
#include <sstream>
#include <string>

#include "libMems/Aligner.h"
#include "libMems/MemHash.h"

using namespace mems;

int main(int argc, char** argv) {
	if (argc >= 2 && std::string(argv[1]) == "smllayout") {
		std::cout << sizeof(SMLHeader) << " " << offsetof(SMLHeader, version) << " " << offsetof(SMLHeader, alphabet_bits) << " "
		          << offsetof(SMLHeader, seed) << " " << offsetof(SMLHeader, seed_length) << " " << offsetof(SMLHeader, seed_weight)
		          << " " << offsetof(SMLHeader, length) << " " << offsetof(SMLHeader, unique_mers) << " "
		          << offsetof(SMLHeader, word_size) << " " << offsetof(SMLHeader, little_endian) << " " << offsetof(SMLHeader, id)
		          << " " << offsetof(SMLHeader, circular) << " " << offsetof(SMLHeader, translation_table) << " "
		          << offsetof(SMLHeader, description) << " " << sizeof(smlSeqI_t) << "\n";
		return 0;
	}
	if (argc < 3) return 2;
	const std::string mode = argv[1];
	try {
		if (mode == "frompos" || mode == "memsfile") {
			const int first = mode == "frompos" ? 4 : 3;
			if (argc <= first) return 2;
			MatchList ml;
			for (int i = first; i < argc; ++i) {
				std::ifstream f(argv[i], std::ios::binary);
				std::string s((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
				ml.seq_table.push_back(new genome::gnSequence(s));
				ml.seq_filename.push_back(argv[i]);
			}
			ml.CreateMemorySMLs((unsigned)atoi(argv[2]), nullptr);
			MemHash mh;
			mh.SetOutputOrder(MEMS_ORDER_REFERENCE);
			if (mode == "memsfile") {
				mh.FindMatches(ml);
				mh.WriteFile(std::cout);
			} else {
				std::vector<uint64_t> sp;
				std::stringstream list(argv[3]);
				std::string tok;
				while (std::getline(list, tok, ',')) sp.push_back(std::stoull(tok));
				std::ostringstream progress, matchlog, offsets;
				mh.LogProgress(&progress);
				mh.SetMatchLog(&matchlog);
				mh.SetOffsetLog(&offsets);
				mh.FindMatchesFromPosition(ml, sp);
				for (Match* m : ml) std::cout << *m << "\n";
				std::cout << "#counts " << mh.MemCount() << " " << mh.MemCollisionCount() << "\n";
				std::cout << "#progress\n" << progress.str() << "#matchlog\n" << matchlog.str() << "#offsets\n" << offsets.str();
			}
			mh.Clear();
			ml.Clear();
			return 0;
		}
		if (mode == "overlaps") {  // host code only: runs without a GPU
			std::ifstream f(argv[2]);
			MatchList ml;
			std::string line;
			while (std::getline(f, line)) {
				std::stringstream ls(line);
				std::vector<int64_t> v;
				int64_t x;
				while (ls >> x) v.push_back(x);
				if (v.size() < 2) continue;
				Match* m = new Match((unsigned)v.size() - 1);
				m->SetLength((uint64_t)v[0]);
				for (size_t i = 1; i < v.size(); ++i) m->SetStart((unsigned)i - 1, v[i]);
				ml.push_back(m);
			}
			EliminateOverlaps(ml);
			for (Match* m : ml) std::cout << *m << "\n";
			for (Match* m : ml) m->Free();
			return 0;
		}
		if (mode == "loadmems") {
			std::ifstream f(argv[2]);
			MemHash mh;
			mh.LoadFile(f);
			MatchList ml;
			mh.GetMatchList(ml);
			for (Match* m : ml) std::cout << *m << "\n";
			std::cout << "#counts " << mh.MemCount() << " " << mh.MemCollisionCount() << "\n";
			for (Match* m : ml) m->Free();
			return 0;
		}
		if (mode == "clone") {
			if (argc < 4) return 2;
			std::ifstream f(argv[3], std::ios::binary);
			std::string s((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
			genome::gnSequence seq(s);
			DNAMemorySML* a = new DNAMemorySML();
			a->Create(seq, getSeed((unsigned)atoi(argv[2])));
			const bmer first = (*a)[0];
			DNAMemorySML* b = a->Clone();
			delete a;  // the clone keeps the device-resident list alive
			const bmer again = (*b)[0];
			std::cout << (first.position == again.position && first.mer == again.mer && b->SMLLength() + b->SeedLength() - 1 == s.size())
			          << " " << b->GetSeedMer(again.position) << " " << again.mer << "\n";
			delete b;
			return 0;
		}
		if (mode == "writesml") {
			if (argc < 5) return 2;
			std::ifstream f(argv[3], std::ios::binary);
			std::string s((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
			genome::gnSequence seq(s);
			DNAFileSML sml(argv[4]);
			sml.Create(seq, getSeed((unsigned)atoi(argv[2])));
			return 0;
		}
		if (mode == "loadsml") {
			DNAFileSML sml;
			sml.LoadFile(argv[2]);
			std::cout << std::hex << sml.Seed() << std::dec << " " << sml.Length() << " " << sml.SMLLength();
			for (uint64_t i = 0; i < 4 && i < sml.SMLLength(); ++i) std::cout << " " << sml[i].position << ":" << sml[i].mer;
			std::cout << "\n";
			return 0;
		}
	} catch (const MemsException& e) {
		std::cerr << "error " << e.code << ": " << e.what() << "\n";
		return 1;
	}
	if (mode == "readmums") {
		std::ifstream f(argv[2]);
		MatchList ml;
		ReadList(ml, f);
		for (Match* m : ml) std::cout << *m << "\n";
		return 0;
	}
	if (argc < 4) return 2;
	const unsigned weight = (unsigned)atoi(argv[2]);
	try {
		MatchList ml;
		for (int i = 3; i < argc; ++i) {
			std::ifstream f(argv[i], std::ios::binary);
			std::string s((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
			ml.seq_table.push_back(new genome::gnSequence(s));
			ml.seq_filename.push_back(argv[i]);
		}
		if (mode == "smlmemhash") {  // on-disk lists next to the sequences (MatchList::LoadSMLs)
			for (int i = 3; i < argc; ++i) ml.sml_filename.push_back(std::string(argv[i]) + ".sml");
			ml.LoadSMLs(weight, &std::cerr);
		} else {
			ml.CreateMemorySMLs(weight, &std::cerr);
		}
		std::cerr << "seed " << std::hex << ml.sml_table[0]->Seed() << std::dec << " length " << ml.sml_table[0]->SeedLength()
		          << " sml[0] = {" << (*ml.sml_table[0])[0].position << ", " << (*ml.sml_table[0])[0].mer << "}\n";
		MemHash* mh = mode == "repeat" ? new RepeatHash() : (mode == "pairwise" ? new PairwiseMatchFinder() : new MemHash());
		mh->SetOutputOrder(MEMS_ORDER_REFERENCE);
		mh->FindMatches(ml);
		std::cerr << "MemCount " << mh->MemCount() << " MemCollisionCount " << mh->MemCollisionCount() << "\n";
		if (mode == "mums") {
			std::stringstream file;
			WriteList(ml, file);
			std::cout << file.str();
			MatchList back;
			ReadList(back, file);
			if (back.size() != ml.size()) return 3;
			for (size_t i = 0; i < ml.size(); ++i)
				if (!(*back[i] == *ml[i])) return 3;
			for (Match* m : back) m->Free();
		} else {
			for (Match* m : ml) std::cout << *m << "\n";
		}
		mh->Clear();
		delete mh;
		ml.Clear();
	} catch (const MemsException& e) {
		std::cerr << "error " << e.code << ": " << e.what() << "\n";
		return 1;
	}
	return 0;
}
