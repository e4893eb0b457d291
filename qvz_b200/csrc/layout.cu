// layout.cu -- ingest / egress re-layout kernels between the host's line-major images and the
// device-resident run-interleaved packed layout (qvz_internal.cuh).
//
// Replaces the per-line pointer table built by load_file (reference src/lines.c:62-79): instead of
// 16 B of host pointer per line, the rows are packed once into Xw[c4][p] in HBM.
//
// All kernels here work on a range of whole RUNS [r0, r0+nr): a range of runs is a contiguous range of
// lines [r0*Lr, (r0+nr)*Lr), so the host side of the copy is one contiguous piece and abi.cu can pipeline
// piece k's PCIe copy with piece k-1's re-layout through two staging buffers.  Thread t of a launch handles
// run r0 + t % nr, step t / nr: adjacent threads touch adjacent slots (coalesced on the packed side).
#include "qvz_internal.cuh"

struct run_range {
	uint32_t r0, nr;         // runs [r0, r0 + nr)
	uint64_t line0;          // first line of the range = r0 * Lr: row 0 of the staging buffer
};

__device__ __forceinline__ bool range_slot(const qvz_layout &L, const run_range &R, uint64_t t, uint64_t &p, uint64_t &line) {
	if (t >= (uint64_t) R.nr * L.Lr) return false;
	const uint32_t i = (uint32_t) (t / R.nr), r = R.r0 + (uint32_t) (t - (uint64_t) i * R.nr);
	p = (uint64_t) i * L.T + r;
	line = (uint64_t) r * L.Lr + i;
	return true;
}

// Reads the line byte-wise (the 32 B sectors are served from L1 after the first touch), validates the
// symbol range the reference silently assumes (src/pmf.c:372-381 indexes a 72-entry table with byte-33),
// writes packed words coalesced across the warp.  Slots past the last line get zero words and id 0xFF.
__global__ void __launch_bounds__(QVZ_THREADS)
qvz_ingest_kernel(qvz_layout L, run_range R, const uint8_t *__restrict__ stage, uint32_t row_stride,
                  uint32_t *__restrict__ Xw, uint8_t *__restrict__ Xb, uint8_t *__restrict__ cl, int *__restrict__ flags)
{
	uint64_t p, line;
	if (!range_slot(L, R, (uint64_t) blockIdx.x * QVZ_THREADS + threadIdx.x, p, line)) return;
	const bool valid = line < L.n_lines;
	const uint8_t *src = stage + (line - R.line0) * (uint64_t) row_stride;
	bool bad = false;
	uint32_t mx = 33;
	for (uint32_t c4 = 0; c4 < L.C4; ++c4) {
		uint32_t w = 0;
		if (valid) {
#pragma unroll
			for (uint32_t j = 0; j < 4; ++j) {
				const uint32_t c = 4 * c4 + j;
				if (c < L.C) {
					const uint32_t b = __ldg(src + c);
					bad |= (b < 33u) | (b >= 33u + QVZ_ALPHABET);
					mx = max(mx, b);
					w |= b << (8 * j);
				}
			}
		}
		Xw[(uint64_t) c4 * L.P + p] = w;
		if (Xb) {                                    // the same bytes as one plane per column (zero = no line)
#pragma unroll
			for (uint32_t j = 0; j < 4; ++j)
				if (4 * c4 + j < L.C) Xb[(uint64_t) (4 * c4 + j) * L.P + p] = (uint8_t) (w >> (8 * j));
		}
	}
	cl[p] = valid ? 0 : QVZ_NO_LINE;
	if (bad) atomicOr(&flags[0], 1);
	mx = __reduce_max_sync(__activemask(), mx);
	if ((threadIdx.x & 31) == 0) atomicMax(&flags[5], (int) mx - 33);
}

__global__ void __launch_bounds__(QVZ_THREADS)
qvz_ids_to_lines_kernel(qvz_layout L, run_range R, const uint8_t *__restrict__ cl, uint8_t *__restrict__ stage)
{
	uint64_t p, line;
	if (!range_slot(L, R, (uint64_t) blockIdx.x * QVZ_THREADS + threadIdx.x, p, line)) return;
	if (line < L.n_lines) stage[line - R.line0] = cl[p];
}

__global__ void __launch_bounds__(QVZ_THREADS)
qvz_ids_from_lines_kernel(qvz_layout L, run_range R, const uint8_t *__restrict__ stage, uint8_t *__restrict__ cl,
                          uint32_t K, int *__restrict__ flags)
{
	uint64_t p, line;
	if (!range_slot(L, R, (uint64_t) blockIdx.x * QVZ_THREADS + threadIdx.x, p, line)) return;
	uint32_t id = QVZ_NO_LINE;
	if (line < L.n_lines) {
		id = stage[line - R.line0];
		if (id >= K) {                               // the kernels index their tables with the id: refuse it here
			atomicOr(&flags[6], 1);
			id = 0;
		}
	}
	cl[p] = (uint8_t) id;
}

// Packed words [C4][P] -> line-major bytes (symbol stream, or the `-u` image with '\n' per line).
__global__ void __launch_bounds__(QVZ_THREADS)
qvz_words_to_lines_kernel(qvz_layout L, run_range R, const uint32_t *__restrict__ Yw, uint8_t *__restrict__ stage,
                          uint32_t out_stride, int add_newline)
{
	uint64_t p, line;
	if (!range_slot(L, R, (uint64_t) blockIdx.x * QVZ_THREADS + threadIdx.x, p, line)) return;
	if (line >= L.n_lines) return;
	uint8_t *dst = stage + (line - R.line0) * (uint64_t) out_stride;
	for (uint32_t c4 = 0; c4 < L.C4; ++c4) {
		const uint32_t w = Yw[(uint64_t) c4 * L.P + p];
#pragma unroll
		for (uint32_t j = 0; j < 4; ++j) {
			const uint32_t c = 4 * c4 + j;
			if (c < L.C) dst[c] = (uint8_t) (w >> (8 * j));
		}
	}
	if (add_newline) dst[L.C] = '\n';
}

__global__ void __launch_bounds__(QVZ_THREADS)
qvz_doubles_to_lines_kernel(qvz_layout L, run_range R, const double *__restrict__ Ep, double *__restrict__ stage)
{
	uint64_t p, line;
	if (!range_slot(L, R, (uint64_t) blockIdx.x * QVZ_THREADS + threadIdx.x, p, line)) return;
	if (line < L.n_lines) stage[line - R.line0] = Ep[p];
}

static inline unsigned range_blocks(const qvz_gpu *h, uint32_t nr) {
	return (unsigned) (((uint64_t) nr * h->L.Lr + QVZ_THREADS - 1) / QVZ_THREADS);
}

static inline run_range make_range(const qvz_gpu *h, uint32_t r0, uint32_t nr) {
	run_range R;
	R.r0 = r0;
	R.nr = nr;
	R.line0 = (uint64_t) r0 * h->L.Lr;
	return R;
}

int qvz_layout_ingest(qvz_gpu *h, uint32_t r0, uint32_t nr, const uint8_t *stage_dev, uint32_t row_stride) {
	qvz_ingest_kernel<<<range_blocks(h, nr), QVZ_THREADS, 0, h->stream>>>(h->L, make_range(h, r0, nr), stage_dev, row_stride,
	                                                                       h->Xw, nullptr /* byte planes are made on first use: cond_counts.cu */, h->cl, h->flags);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}

int qvz_layout_ids_to_lines(qvz_gpu *h, uint32_t r0, uint32_t nr, uint8_t *stage_dev) {
	qvz_ids_to_lines_kernel<<<range_blocks(h, nr), QVZ_THREADS, 0, h->stream>>>(h->L, make_range(h, r0, nr), h->cl, stage_dev);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}

int qvz_layout_ids_from_lines(qvz_gpu *h, uint32_t r0, uint32_t nr, const uint8_t *stage_dev, uint32_t K) {
	qvz_ids_from_lines_kernel<<<range_blocks(h, nr), QVZ_THREADS, 0, h->stream>>>(h->L, make_range(h, r0, nr), stage_dev, h->cl, K, h->flags);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}

int qvz_layout_words_to_lines(qvz_gpu *h, uint32_t r0, uint32_t nr, const uint32_t *Yw, uint8_t *stage_dev,
                              uint32_t out_stride, int add_newline) {
	qvz_words_to_lines_kernel<<<range_blocks(h, nr), QVZ_THREADS, 0, h->stream>>>(h->L, make_range(h, r0, nr), Yw, stage_dev,
	                                                                               out_stride, add_newline);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}

int qvz_layout_doubles_to_lines(qvz_gpu *h, uint32_t r0, uint32_t nr, const double *Ep, double *stage_dev) {
	qvz_doubles_to_lines_kernel<<<range_blocks(h, nr), QVZ_THREADS, 0, h->stream>>>(h->L, make_range(h, r0, nr), Ep, stage_dev);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}
