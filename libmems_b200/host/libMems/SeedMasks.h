// libMems/SeedMasks.h façade — same free functions as the reference (SeedMasks.h:276-401), served by the
// C-ABI's host-side table.
#pragma once
#include <climits>
#include "mems_b200.h"

static const int CODING_SEED = 3;
static const int SOLID_SEED = INT_MAX;
inline int64_t getSolidSeed(int weight) { return (int64_t)mems_get_solid_seed(weight); }
inline int64_t getSeed(int weight, int seed_rank = 0) { return (int64_t)mems_get_seed(weight, seed_rank); }
inline int getSeedLength(int64_t seed) { return mems_get_seed_length((uint64_t)seed); }
inline int getSeedWeight(int64_t seed) { return mems_get_seed_weight((uint64_t)seed); }
inline unsigned getDefaultSeedWeight(uint64_t avg_sequence_length) { return mems_get_default_seed_weight(avg_sequence_length); }
