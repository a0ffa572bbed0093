// oracle/shim: minimal boost::filesystem::path for libMems/Files.h (temp-file helpers, never on the hot path).
#pragma once
#include <string>
#include <cstdio>
namespace boost { namespace filesystem {
class path {
public:
	path() {}
	path(const std::string& s) : p(s) {}
	path(const char* s) : p(s) {}
	path branch_path() const { size_t i = p.rfind('/'); return i == std::string::npos ? path("") : path(p.substr(0, i)); }
	path leaf() const { size_t i = p.rfind('/'); return i == std::string::npos ? *this : path(p.substr(i + 1)); }
	std::string string() const { return p; }
	path& operator/=(const std::string& s) { p += "/" + s; return *this; }
private:
	std::string p;
};
inline bool remove(const std::string& s) { return ::remove(s.c_str()) == 0; }
inline bool exists(const path&) { return false; }
} }
