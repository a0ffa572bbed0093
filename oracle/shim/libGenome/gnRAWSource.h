// oracle/shim: stand-in for libGenome/gnRAWSource.h (never called on the hot path). Test infrastructure.
#pragma once
#include "libGenome/gnSequence.h"
namespace genome {
class gnRAWSource {
public:
	static void Write(gnSequence&, const std::string&) {}
};
}
