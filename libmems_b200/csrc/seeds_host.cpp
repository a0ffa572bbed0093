// seeds_host.cpp — spaced-seed pattern table and its helpers (SeedMasks.h:44-401), host arithmetic.
#include <climits>
#include <cmath>
#include <cstdint>

#include "mems_b200.h"

namespace {
// Published seed patterns (Darling et al. 2006) as libMems tabulates them; generated data, see
// tools/gen_seed_table.py.  Row "11" really holds a weight-12 pattern — kept.
const uint64_t kSeedTable[32][6] = {
#include "mems_seed_table.inc"
};
}

extern "C" {

uint64_t mems_get_solid_seed(int weight) {  // getSolidSeed, SeedMasks.h:276-281
	if (weight >= 64) return ~0ull;
	if (weight <= 0) return 0;
	return (1ull << weight) - 1ull;
}

uint64_t mems_get_seed(int weight, int seed_rank) {  // getSeed, SeedMasks.h:298-321
	if (seed_rank == INT_MAX) return mems_get_solid_seed(weight);  // SOLID_SEED
	if (weight > 31) return mems_get_solid_seed(32);
	if (seed_rank > 5) return mems_get_solid_seed(weight);
	if (weight < 0 || seed_rank < 0) return 0;  // out of the reference's table (undefined there)
	return kSeedTable[weight][seed_rank];
}

int mems_get_seed_length(uint64_t seed) {  // getSeedLength, SeedMasks.h:335-350
	if (seed == 0) return 0;
	return 64 - __builtin_clzll(seed) - __builtin_ctzll(seed);
}

int mems_get_seed_weight(uint64_t seed) {  // getSeedWeight, SeedMasks.h:362-373
	return __builtin_popcountll(seed);
}

unsigned mems_get_default_seed_weight(uint64_t avg_seq_len) {  // getDefaultSeedWeight, SeedMasks.h:389-401
	if (avg_seq_len == 0) return 0;
	unsigned w = (unsigned)std::ceil((std::log((double)avg_seq_len) / std::log(2.0)) / 1.5);
	if (!(w & 1u)) ++w;  // even weights can be palindromic
	if (w < 5) w = 0;    // MIN_DNA_SEED_WEIGHT
	if (w > 31) w = 31;  // MAX_DNA_SEED_WEIGHT
	return w;
}

}  // extern "C"
