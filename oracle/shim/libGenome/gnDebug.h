// oracle/shim: stand-in for libGenome/gnDebug.h. Test infrastructure.
#pragma once
#include <string>
#include <iostream>
namespace genome {
inline void DebugMsg(const std::string&) {}
inline void ErrorMsg(const std::string& s) { std::cerr << s; }
}
