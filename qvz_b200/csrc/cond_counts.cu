// cond_counts.cu -- first-order conditional counts per (cluster, column, previous RAW value, value).
//
// Reference: the counting loop of calculate_statistics (src/codebook.c:193-205):
//     pmf_increment(get_cond_pmf(list, 0, 0), x[0]-33)
//     pmf_increment(get_cond_pmf(list, c, x[c-1]-33), x[c]-33)          for c >= 1
// with get_cond_pmf -> pmfs[1 + (c-1)*72 + prev] (src/codebook.c:116-120) and pmf_increment ->
// counts[idx]++, total++ (src/pmf.c:211-214).  Output layout = that pmfs[] order, 72 counters per row.
//
// v1 mapping: one CTA owns one packed word column c4 (= 4 table columns) for a chunk of slots and up to
// QVZ_CC_GROUP clusters; its 72x72 slices live in shared memory as packed 16-bit counter pairs
// (chunk <= 65 280 slots so no counter can overflow); lanes <-> slots so global reads are coalesced;
// shared atomics; non-zero counters are flushed with one global atomic each.
#include "qvz_internal.cuh"

#define QVZ_CC_GROUP 5u                         // clusters per pass: 5 * 4 * 72*72 * 2 B = 207 360 B of shared memory
#define QVZ_CC_SLICE_WORDS (72u * 72u / 2u)      // 2592 words per (cluster, column) slice
#define QVZ_CC_CHUNK 65280u                      // slots per CTA (multiple of 256, < 65536)
#define QVZ_CC_THREADS 1024
#define QVZ_CC_UNROLL 4

__global__ void __launch_bounds__(QVZ_CC_THREADS)
qvz_cond_counts_kernel(qvz_layout L, const uint32_t *__restrict__ Xw, const uint8_t *__restrict__ cl,
                       uint32_t K, uint32_t G, uint32_t *__restrict__ counts)
{
	extern __shared__ uint32_t tab[];            // [G][4][2592]
	const uint32_t c4 = blockIdx.x;
	const uint32_t kbase = blockIdx.z * QVZ_CC_GROUP;
	const uint64_t p0 = (uint64_t) blockIdx.y * QVZ_CC_CHUNK;
	const uint64_t p1 = (p0 + QVZ_CC_CHUNK < L.P) ? p0 + QVZ_CC_CHUNK : L.P;
	const uint32_t words = G * 4 * QVZ_CC_SLICE_WORDS;

	for (uint32_t i = threadIdx.x; i < words; i += QVZ_CC_THREADS) tab[i] = 0;
	__syncthreads();

	const uint32_t *xc = Xw + (uint64_t) c4 * L.P;
	const uint32_t *xp = c4 ? xc - L.P : xc;
	for (uint64_t pb = p0 + threadIdx.x; pb < p1; pb += (uint64_t) QVZ_CC_UNROLL * QVZ_CC_THREADS) {
		uint32_t kk[QVZ_CC_UNROLL], ww[QVZ_CC_UNROLL], pp[QVZ_CC_UNROLL];
#pragma unroll
		for (int u = 0; u < QVZ_CC_UNROLL; ++u) {    // all loads first: 3*UNROLL independent requests in flight
			const uint64_t p = pb + (uint64_t) u * QVZ_CC_THREADS;
			const bool in = p < p1;
			kk[u] = in ? cl[p] : QVZ_NO_LINE;
			ww[u] = in ? xc[p] : 0u;
			pp[u] = (in && c4) ? xp[p] : 0x21212121u;
		}
#pragma unroll
		for (int u = 0; u < QVZ_CC_UNROLL; ++u) {
			const uint32_t g = kk[u] - kbase;        // wraps to a huge value for k < kbase and for 0xFF
			if (g >= G) continue;
			const uint32_t w = ww[u] - 0x21212121u;  // ingest guarantees every real byte >= 33: no borrow
			uint32_t prev = c4 ? ((pp[u] >> 24) - 33u) : 0u;
#pragma unroll
			for (uint32_t j = 0; j < 4; ++j) {
				if (4 * c4 + j < L.C) {
					const uint32_t cur = (w >> (8 * j)) & 0xFFu;
					const uint32_t bin = prev * 72u + cur;
					atomicAdd(&tab[(g * 4 + j) * QVZ_CC_SLICE_WORDS + (bin >> 1)], 1u << (16 * (bin & 1)));
					prev = cur;
				}
			}
		}
	}
	__syncthreads();

	const uint64_t per_cluster = (uint64_t) (1 + 72 * (L.C - 1)) * 72;
	for (uint32_t i = threadIdx.x; i < words; i += QVZ_CC_THREADS) {
		const uint32_t v = tab[i];
		if (!v) continue;
		const uint32_t slice = i / QVZ_CC_SLICE_WORDS, wbin = i - slice * QVZ_CC_SLICE_WORDS;
		const uint32_t g = slice >> 2, j = slice & 3, col = 4 * c4 + j;
		// column 0 only ever sees prev == 0 => bins 0..71 => pmfs[0]; column c >= 1 starts at pmfs[1 + (c-1)*72]
		uint32_t *dst = counts + (uint64_t) (kbase + g) * per_cluster + (col ? (uint64_t) (1 + (col - 1) * 72) * 72 : 0) + 2 * wbin;
		if (v & 0xFFFFu) atomicAdd(dst, v & 0xFFFFu);
		if (v >> 16) atomicAdd(dst + 1, v >> 16);
	}
}

int qvz_cond_counts_launch(qvz_gpu *h, uint32_t *counts_dev) {
	const uint32_t K = h->K;
	const uint32_t G = K < QVZ_CC_GROUP ? K : QVZ_CC_GROUP;
	const size_t smem = (size_t) G * 4 * QVZ_CC_SLICE_WORDS * sizeof(uint32_t);
	QVZ_CUDA(h, cudaMemsetAsync(counts_dev, 0, qvz_gpu_cond_counts_len(K, h->L.C) * sizeof(uint32_t), h->stream));
	QVZ_CUDA(h, cudaFuncSetAttribute(qvz_cond_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	dim3 grid(h->L.C4, (unsigned) ((h->L.P + QVZ_CC_CHUNK - 1) / QVZ_CC_CHUNK), (K + QVZ_CC_GROUP - 1) / QVZ_CC_GROUP);
	// the last cluster group may be partial: G applies to all groups, out-of-range ids are skipped by g >= G
	qvz_cond_counts_kernel<<<grid, QVZ_CC_THREADS, smem, h->stream>>>(h->L, h->Xw, h->cl, K, G, counts_dev);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}
