// kernels_match.cu — placeholder until the match pipeline lands (next commit).
#include "common.cuh"
#include "mems_b200.h"
namespace mems {
void find_matches_on_batch(Batch&, int, int, uint32_t, MatchResult&) {
	throw Error(MEMS_ERR_UNSUPPORTED, "match finding not built yet");
}
}
