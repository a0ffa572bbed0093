// oracle/shim: never used by the in-memory path; declared so libMems/gnRAWSequence.h parses.
#pragma once
#include <string>
namespace boost { namespace iostreams {
class mapped_file_source {
public:
	void open(const std::string&, size_t = 0, size_t = 0) {}
	const char* data() const { return 0; }
	size_t size() const { return 0; }
};
} }
