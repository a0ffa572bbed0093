// oracle/shim: stand-in for libGenome/gnDefs.h (typedefs only). Test infrastructure.
#pragma once
#include <stdint.h>
#include <limits.h>
#include <limits>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <string>
#include <iostream>
typedef uint8_t uint8;
typedef uint16_t uint16;
typedef uint32_t uint32;
typedef unsigned long long uint64;
typedef int8_t int8;
typedef int16_t int16;
typedef int32_t int32;
typedef long long int64;
typedef unsigned int uint;
typedef float float32;
typedef double float64;
typedef unsigned char boolean;
typedef char gnSeqC;
typedef uint64 gnSeqI;
#ifndef GNSEQI_END
#define GNSEQI_END UINT32_MAX
#endif
#define GNSEQI_ERROR GNSEQI_END
#define GNSEQI_BEGIN 0
#define GNDLLEXPORT
#define ALL_CONTIGS UINT32_MAX
namespace genome {
template <class T> inline T absolut(const T& t) { return t < 0 ? -t : t; }
}
