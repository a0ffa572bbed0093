// cub_sort_bench.cu — YARDSTICK ONLY (never linked into the product): cub::DeviceRadixSort::SortPairs from the CUDA
// toolkit on the same kind of input as the SML build's sort, timed with CUDA events.  SURVEY.md §7-4 asks for it:
// it tells how far the repository's own onesweep passes (libmems_b200/csrc/radix_sort.cu) are from NVIDIA's.
//   build/cub_sort_bench <n items> <key bits> [iterations]      key type = u32 if bits <= 32 else u64, values u32
// Prints one JSON line: total ms per sort, passes CUB runs for those bits (8-bit digits), ms and GB/s per pass
// (algorithmic bytes 2 * n * (key bytes + 4) per pass, the same accounting as bench.py's roofline).
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <class K>
__global__ void fill(K* keys, uint32_t* vals, uint64_t n, int bits) {
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint64_t x = i * 0x9E3779B97F4A7C15ull + 0x1234567ull;  // splitmix64
	x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
	keys[i] = (K)(bits >= 64 ? x : (x & ((1ull << bits) - 1ull)));
	vals[i] = (uint32_t)i;
}

template <class K>
int run(uint64_t n, int bits, int iters) {
	K *ka, *kb;
	uint32_t *va, *vb;
	CK(cudaMalloc(&ka, n * sizeof(K))); CK(cudaMalloc(&kb, n * sizeof(K)));
	CK(cudaMalloc(&va, n * 4)); CK(cudaMalloc(&vb, n * 4));
	K* src; CK(cudaMalloc(&src, n * sizeof(K)));
	fill<K><<<(unsigned)((n + 255) / 256), 256>>>(src, va, n, bits);
	cub::DoubleBuffer<K> dk(ka, kb);
	cub::DoubleBuffer<uint32_t> dv(va, vb);
	size_t tmp_bytes = 0;
	CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (int)n, 0, bits));
	void* tmp; CK(cudaMalloc(&tmp, tmp_bytes));
	cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
	float total = 0;
	for (int it = 0; it < iters + 3; ++it) {
		CK(cudaMemcpy(dk.Current(), src, n * sizeof(K), cudaMemcpyDeviceToDevice));  // unsorted input every time
		CK(cudaDeviceSynchronize());
		CK(cudaEventRecord(a));
		CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dv, (int)n, 0, bits));
		CK(cudaEventRecord(b));
		CK(cudaEventSynchronize(b));
		float ms; CK(cudaEventElapsedTime(&ms, a, b));
		if (it >= 3) total += ms;
	}
	const double ms = total / iters;
	const int passes = (bits + 7) / 8;
	const double bytes_per_pass = 2.0 * (double)n * (sizeof(K) + 4.0);
	printf("{\"impl\": \"cub::DeviceRadixSort::SortPairs\", \"n\": %llu, \"key_bits\": %d, \"key_bytes\": %d, \"ms_per_sort\": %.4f, "
	       "\"passes\": %d, \"ms_per_pass\": %.4f, \"gbs_per_pass\": %.1f}\n",
	       (unsigned long long)n, bits, (int)sizeof(K), ms, passes, ms / passes, bytes_per_pass / (ms / passes) / 1e6);
	return 0;
}

int main(int argc, char** argv) {
	const uint64_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 40000000ull;
	const int bits = argc > 2 ? atoi(argv[2]) : 31;
	const int iters = argc > 3 ? atoi(argv[3]) : 10;
	return bits <= 32 ? run<uint32_t>(n, bits, iters) : run<uint64_t>(n, bits, iters);
}
