// seed_dev.cuh — device-side spaced-seed arithmetic on 2-bit packed DNA.
//
// Layout (identical to SortedMerList::SetSequence/translate32, SortedMerList.cpp:306-317,425-460):
// base p sits in bits 31-2(p%16) .. 30-2(p%16) of uint32 word p/16, i.e. MSB first; A,C,G,T = 0..3.
#pragma once
#include "common.cuh"

namespace mems {

// 64-bit window whose base 0 is base p of the sequence (left-justified, like SortedMerList::GetMer,
// SortedMerList.cpp:321-342, before its mer_mask).  Reads words p/16 .. p/16+2.
template <class WordPtr>
__device__ __forceinline__ uint64_t window64(WordPtr words, uint32_t p) {
	uint32_t i = p >> 4, sh = (p & 15u) * 2u;
	uint32_t a = words[i], b = words[i + 1], c = words[i + 2];
	uint32_t hi = __funnelshift_l(b, a, sh);
	uint32_t lo = __funnelshift_l(c, b, sh);
	return ((uint64_t)hi << 32) | lo;
}

// Bases under the pattern's one-bits, right-justified (2w bits): software PEXT over the pattern's runs.
__device__ __forceinline__ uint64_t extract_fwd(uint64_t win, const SeedDesc& sd) {
	uint64_t f = 0;
	// fully unrolled with a uniform early exit: run descriptors are then read from the kernel-parameter
	// constant bank at fixed offsets (a runtime-indexed parameter array would be spilled to local memory)
#pragma unroll
	for (int r = 0; r < kMaxSeedRuns; ++r) {
		if (r >= sd.n_runs) break;
		f |= (win >> sd.run_net[r]) & sd.run_mask[r];  // one shift + one and-or (LOP3) per 32-bit half
	}
	return f;
}

// Reverse complement of a right-justified w-base value (SortedMerList::RevCompMer, :597-614, without
// its 32-iteration loop): complement, reverse all 64 bits, swap the two bits of every base back.
__device__ __forceinline__ uint64_t revcomp_w(uint64_t fwd, int w) {
	uint64_t x = __brevll(~fwd);
	x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
	return x >> (64 - 2 * w);
}

// Compact canonical key: (min(fwd, rc) << 1) | strand, strand = 1 iff the reverse complement is
// strictly smaller (SortedMerList::GetDnaSeedMer, :764-769: forward wins ties because rc carries bit 0).
// Ordering by this key equals ordering by the reference's 64-bit mer.
__device__ __forceinline__ uint64_t canonical_key(uint64_t fwd, int w) {
	if (w <= 16) {  // the whole mer fits 32 bits: half the work (warp-uniform branch)
		const uint32_t f = (uint32_t)fwd;
		uint32_t x = __brev(~f);
		x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
		const uint32_t rc = w == 16 ? x : (x >> (32 - 2 * w));
		return f <= rc ? ((uint64_t)f << 1) : (((uint64_t)rc << 1) | 1ull);
	}
	uint64_t rc = revcomp_w(fwd, w);
	return fwd <= rc ? (fwd << 1) : ((rc << 1) | 1ull);
}

// Reference 64-bit layout: w-mer left-justified, strand flag in bit 0.
__device__ __forceinline__ uint64_t to_reference_mer(uint64_t ckey, int w) {
	return ((ckey >> 1) << (64 - 2 * w)) | (ckey & 1ull);
}

}  // namespace mems
