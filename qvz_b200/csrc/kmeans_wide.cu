// kmeans_wide.cu -- one k-means iteration for MORE than QVZ_MAX_K clusters (the reference accepts any uint8_t count,
// include/codebook.h:31; src/main.c:263 `-c`).
//
// Same arithmetic as kmeans.cu (find_distance / assign_cluster / recalculate_means, src/cluster.c:80-187): integer
// distances argmin_k (sum m_k^2 - 2 sum x*m_k) with dp4a, strict '<' in cluster order so the lowest id wins ties, exact
// integer column sums.  What differs is the shape: the distances of 16 clusters at a time live in registers and the
// running best is carried from group to group (the row words are re-read per group, coalesced), and the column sums
// are a second, word-column-tiled pass with shared-memory partial sums.  This is the general path, not the fast one.
#include "qvz_internal.cuh"

#define KW_GROUP 16

__global__ void __launch_bounds__(QVZ_THREADS)
qvz_kmeans_wide_assign_kernel(qvz_layout L, const uint32_t *__restrict__ Xw, uint8_t *__restrict__ cl,
                              const uint32_t *__restrict__ means_w, const uint32_t *__restrict__ means_sq,
                              uint32_t K, const uint32_t *__restrict__ ctl)
{
	if (ctl[QVZ_CTL_DONE]) return;
	extern __shared__ __align__(16) uint32_t mean4[];    // [C4][KW_GROUP] centroid words of the current group
	__shared__ uint32_t msq[KW_GROUP];
	const uint32_t C4 = L.C4, tid = threadIdx.x;
	for (uint64_t base = (uint64_t) blockIdx.x * QVZ_THREADS; base < L.P; base += (uint64_t) gridDim.x * QVZ_THREADS) {
		const uint64_t p = base + tid;                   // P % 256 == 0: every thread of the CTA has a slot
		int bestv = 0x7FFFFFFF;
		uint32_t best = 0;
		for (uint32_t k0 = 0; k0 < K; k0 += KW_GROUP) {
			__syncthreads();
			for (uint32_t i = tid; i < C4 * KW_GROUP; i += QVZ_THREADS) {
				const uint32_t c4 = i / KW_GROUP, k = k0 + (i - c4 * KW_GROUP);
				mean4[i] = k < K ? means_w[k * C4 + c4] : 0u;
			}
			if (tid < KW_GROUP) msq[tid] = k0 + tid < K ? means_sq[k0 + tid] : 0u;
			__syncthreads();
			uint32_t D[KW_GROUP];
#pragma unroll
			for (int k = 0; k < KW_GROUP; ++k) D[k] = 0;
			for (uint32_t c4 = 0; c4 < C4; ++c4) {
				const uint32_t w = Xw[(uint64_t) c4 * L.P + p];
				const uint4 *m = (const uint4 *) (mean4 + c4 * KW_GROUP);
#pragma unroll
				for (int g = 0; g < KW_GROUP / 4; ++g) {
					const uint4 mm = m[g];
					D[4 * g + 0] = __dp4a(w, mm.x, D[4 * g + 0]);
					D[4 * g + 1] = __dp4a(w, mm.y, D[4 * g + 1]);
					D[4 * g + 2] = __dp4a(w, mm.z, D[4 * g + 2]);
					D[4 * g + 3] = __dp4a(w, mm.w, D[4 * g + 3]);
				}
			}
#pragma unroll
			for (int k = 0; k < KW_GROUP; ++k) {
				const int v = (int) msq[k] - 2 * (int) D[k];
				if (k0 + k < K && v < bestv) {           // strict '<', ascending ids: the lowest id wins ties (assign_cluster)
					bestv = v;
					best = k0 + k;
				}
			}
		}
		if (cl[p] != QVZ_NO_LINE) cl[p] = (uint8_t) best;
	}
}

// grid (C4, chunks): accumulator[cluster][col] += byte over a chunk of slots of one word column
__global__ void __launch_bounds__(QVZ_THREADS)
qvz_kmeans_wide_sums_kernel(qvz_layout L, const uint32_t *__restrict__ Xw, const uint8_t *__restrict__ cl,
                            uint32_t K, uint64_t chunk, unsigned long long *__restrict__ sums, const uint32_t *__restrict__ ctl)
{
	if (ctl[QVZ_CTL_DONE]) return;
	extern __shared__ uint32_t acc[];                    // [K][4] byte sums of this chunk, then [K] line counts
	uint32_t *cnt = acc + 4 * K;
	const uint32_t c4 = blockIdx.x, tid = threadIdx.x;
	for (uint32_t i = tid; i < 5 * K; i += QVZ_THREADS) acc[i] = 0;
	__syncthreads();
	const uint64_t p0 = (uint64_t) blockIdx.y * chunk, p1 = p0 + chunk < L.P ? p0 + chunk : L.P;
	for (uint64_t p = p0 + tid; p < p1; p += QVZ_THREADS) {
		const uint32_t k = cl[p];
		if (k == QVZ_NO_LINE) continue;
		const uint32_t w = Xw[(uint64_t) c4 * L.P + p];
#pragma unroll
		for (uint32_t j = 0; j < 4; ++j) atomicAdd(&acc[4 * k + j], (w >> (8 * j)) & 0xFFu);
		if (c4 == 0) atomicAdd(&cnt[k], 1u);
	}
	__syncthreads();
	for (uint32_t i = tid; i < 4 * K; i += QVZ_THREADS) {
		const uint32_t k = i >> 2, col = 4 * c4 + (i & 3);
		if (col < L.C && acc[i]) atomicAdd(&sums[(uint64_t) k * L.C + col], (unsigned long long) acc[i]);
	}
	if (c4 == 0)
		for (uint32_t k = tid; k < K; k += QVZ_THREADS)
			if (cnt[k]) atomicAdd(&sums[(uint64_t) K * L.C + k], (unsigned long long) cnt[k]);
}

int qvz_kmeans_launch_assign_wide(qvz_gpu *h, int64_t *sums_dev) {
	const uint32_t K = h->km_K, C4 = h->L.C4;
	const size_t sum_bytes = ((size_t) K * h->L.C + K) * sizeof(int64_t);
	const uint64_t blocks = h->L.P / QVZ_THREADS;
	const uint64_t cap = (uint64_t) h->sm_count * 8;
	qvz_kmeans_wide_assign_kernel<<<(unsigned) (blocks < cap ? blocks : cap), QVZ_THREADS, (size_t) C4 * KW_GROUP * sizeof(uint32_t), h->stream>>>(
	    h->L, h->Xw, h->cl, h->means_w, h->means_sq, K, h->km_ctl);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	QVZ_CUDA(h, cudaMemsetAsync(sums_dev, 0, sum_bytes, h->stream));
	const uint64_t chunk = 1u << 20;                     // 2^20 slots * 255 < 2^32: the 32-bit partial sums cannot overflow
	dim3 grid(C4, (unsigned) ((h->L.P + chunk - 1) / chunk));
	qvz_kmeans_wide_sums_kernel<<<grid, QVZ_THREADS, (size_t) 5 * K * sizeof(uint32_t), h->stream>>>(
	    h->L, h->Xw, h->cl, K, chunk, (unsigned long long *) sums_dev, h->km_ctl);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}
