// oracle/shim: stand-in for libGenome/gnException.h (exception codes + throw macros).
// Test infrastructure.
#pragma once
#include <string>
#include <iostream>
namespace genome {
class gnExceptionCode {
public:
	explicit gnExceptionCode(const char* n) : name(n) {}
	bool operator==(const gnExceptionCode& o) const { return name == o.name; }
	std::string name;
};
class gnException {
public:
	gnException(const char*, int, const char*, gnExceptionCode& c, const char* m = "") : code(c), msg(m) {}
	gnExceptionCode& GetCode() { return code; }
	gnExceptionCode& code;
	std::string msg;
};
inline std::ostream& operator<<(std::ostream& os, const gnException& e) { return os << e.code.name << ": " << e.msg; }
}
#define CREATE_EXCEPTION(E) \
	inline genome::gnExceptionCode& E() { static genome::gnExceptionCode* c = new genome::gnExceptionCode(#E); return *c; }
#define Throw_gnEx(code) throw genome::gnException(__FILE__, __LINE__, __func__, code)
#define Throw_gnExMsg(code, msg) throw genome::gnException(__FILE__, __LINE__, __func__, code, msg)
#define STACK_TRACE_START
#define STACK_TRACE_END
namespace genome {
CREATE_EXCEPTION(IndexOutOfBounds)
CREATE_EXCEPTION(NullPointer)
CREATE_EXCEPTION(SeqIndexOutOfBounds)
CREATE_EXCEPTION(FileNotOpened)
CREATE_EXCEPTION(FileUnreadable)
CREATE_EXCEPTION(IOStreamFailed)
CREATE_EXCEPTION(FragmentIndexOutOfBounds)
CREATE_EXCEPTION(FeatureIndexOutOfBounds)
CREATE_EXCEPTION(HeaderIndexOutOfBounds)
}
