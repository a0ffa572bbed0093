// oracle/shim: stand-in for libGenome/gnClone.h. Test infrastructure.
#pragma once
#include "libGenome/gnDefs.h"
namespace genome {
class gnClone {
public:
	virtual ~gnClone() {}
	virtual gnClone* Clone() const = 0;
};
}
