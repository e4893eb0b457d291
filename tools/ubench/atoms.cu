// microbenchmark: cost of shared-memory atomics per warp instruction under different address patterns / forms
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE, int PAT>
__global__ void k(uint32_t *out, const uint32_t *idx, int iters, unsigned long long *cycles) {
	__shared__ uint32_t tab[8192];
	for (int i = threadIdx.x; i < 8192; i += blockDim.x) tab[i] = 0;
	__syncthreads();
	const uint32_t lane = threadIdx.x & 31;
	uint32_t base = (uint32_t) __cvta_generic_to_shared(tab);
	uint32_t a[8];
	for (int j = 0; j < 8; ++j) {
		uint32_t w;
		if (PAT == 0) w = 5;                                  // all lanes same address
		else if (PAT == 1) w = lane + 32 * j;                 // 32 distinct banks
		else if (PAT == 2) w = idx[(threadIdx.x * 8 + j) & 8191] & 8191;   // random
		else if (PAT == 3) w = (lane & 7) * 33 + j * 300;     // 8 distinct addresses, 4 lanes each, distinct banks
		else if (PAT == 5) w = lane + 32 * (idx[(threadIdx.x * 8 + j) & 8191] & 255);   // lane-private bank, random row per lane
		else if (PAT == 6) w = lane + 32 * ((idx[(threadIdx.x * 8 + j) & 8191] >> 3) & 3);   // lane-private bank, 4 distinct rows
		else w = (lane & 3) * 1296 * 4 / 4 + (lane >> 2) + 64 * j;  // 4 slices x 8 distinct
		a[j] = base + 4 * w;
	}
	__syncthreads();
	unsigned long long t0 = clock64();
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			if (MODE == 0) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a[j]) : "memory");
			else if (MODE == 1) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a[j]), "r"(lane + 2) : "memory");
			else if (MODE == 2) { uint32_t r; asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(r) : "r"(a[j]) : "memory"); if (r == 0xFFFFFFFF) out[0] = r; }
			else if (MODE == 3) asm volatile("st.shared.u32 [%0], %1;" ::"r"(a[j]), "r"(lane) : "memory");
			else { uint32_t r; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a[j]) : "memory"); if (r == 0xFFFFFFFF) out[0] = r; }
		}
	}
	__syncthreads();
	unsigned long long t1 = clock64();
	if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
	if (threadIdx.x < 4) out[1 + threadIdx.x] = tab[threadIdx.x];
}
template <int MODE, int PAT> void run(const char *name, uint32_t *out, uint32_t *idx, unsigned long long *cyc, int threads) {
	const int iters = 2000;
	k<MODE, PAT><<<148, threads>>>(out, idx, iters, cyc);
	cudaDeviceSynchronize();
	unsigned long long h[148];
	cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
	double c = (double) h[0] / ((double) iters * 8 * (threads / 32));
	printf("%-44s threads=%4d  cycles per warp-instr (per SM) = %.2f\n", name, threads, c);
}
int main() {
	uint32_t *out, *idx; unsigned long long *cyc;
	cudaMalloc(&out, 64); cudaMalloc(&idx, 8192 * 4); cudaMalloc(&cyc, 148 * 8);
	uint32_t h[8192]; uint32_t s = 12345; for (int i = 0; i < 8192; ++i) { s = s * 1664525u + 1013904223u; h[i] = s >> 8; }
	cudaMemcpy(idx, h, sizeof(h), cudaMemcpyHostToDevice);
	for (int threads : {256, 1024}) {
		run<0, 0>("red +1 (POPC.INC), same address", out, idx, cyc, threads);
		run<0, 1>("red +1 (POPC.INC), 32 distinct banks", out, idx, cyc, threads);
		run<0, 2>("red +1 (POPC.INC), random", out, idx, cyc, threads);
		run<0, 3>("red +1 (POPC.INC), 8 addr x 4 lanes", out, idx, cyc, threads);
		run<1, 0>("red +v (ATOMS.ADD), same address", out, idx, cyc, threads);
		run<1, 1>("red +v (ATOMS.ADD), 32 distinct banks", out, idx, cyc, threads);
		run<1, 2>("red +v (ATOMS.ADD), random", out, idx, cyc, threads);
		run<1, 3>("red +v (ATOMS.ADD), 8 addr x 4 lanes", out, idx, cyc, threads);
		run<0, 5>("red +1 (POPC.INC), lane = bank, random rows", out, idx, cyc, threads);
		run<0, 6>("red +1 (POPC.INC), lane = bank, 4 rows", out, idx, cyc, threads);
		run<1, 5>("red +v (ATOMS.ADD), lane = bank, random rows", out, idx, cyc, threads);
		run<3, 5>("STS, lane = bank, random rows", out, idx, cyc, threads);
		run<4, 5>("LDS, lane = bank, random rows", out, idx, cyc, threads);
		run<2, 1>("atom +1 returning, 32 distinct banks", out, idx, cyc, threads);
		run<3, 1>("STS, 32 distinct banks", out, idx, cyc, threads);
		run<3, 2>("STS, random", out, idx, cyc, threads);
		run<4, 1>("LDS, 32 distinct banks", out, idx, cyc, threads);
		run<4, 2>("LDS, random", out, idx, cyc, threads);
	}
	return 0;
}
