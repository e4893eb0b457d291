// layout.cu -- ingest / egress re-layout kernels between the host's line-major images and the
// device-resident run-interleaved packed layout (qvz_internal.cuh).
//
// Replaces the per-line pointer table built by load_file (reference src/lines.c:62-79): instead of
// 16 B of host pointer per line, the rows are packed once into Xw[c4][p] in HBM.
#include "qvz_internal.cuh"

// One thread per slot.  Reads its line byte-wise (the 32 B sectors are served from L1 after the first
// touch), validates the symbol range the reference silently assumes (src/pmf.c:372-381 indexes a
// 72-entry table with byte-33), writes packed words coalesced across the warp.
__global__ void __launch_bounds__(QVZ_THREADS)
qvz_ingest_kernel(qvz_layout L, const uint8_t *__restrict__ raw, uint32_t row_stride,
                  uint32_t *__restrict__ Xw, uint8_t *__restrict__ cl, int *__restrict__ flags)
{
	uint64_t p = (uint64_t) blockIdx.x * QVZ_THREADS + threadIdx.x;
	if (p >= L.P) return;
	uint64_t line = qvz_slot_line(L, p);
	bool valid = line < L.n_lines;
	const uint8_t *src = raw + line * (uint64_t) row_stride;
	bool bad = false;
	uint32_t mx = 33;
	for (uint32_t c4 = 0; c4 < L.C4; ++c4) {
		uint32_t w = 0;
		if (valid) {
#pragma unroll
			for (uint32_t j = 0; j < 4; ++j) {
				uint32_t c = 4 * c4 + j;
				if (c < L.C) {
					uint32_t b = __ldg(src + c);
					bad |= (b < 33u) | (b >= 33u + QVZ_ALPHABET);
					mx = max(mx, b);
					w |= b << (8 * j);
				}
			}
		}
		Xw[(uint64_t) c4 * L.P + p] = w;
	}
	cl[p] = valid ? 0 : QVZ_NO_LINE;
	if (bad) atomicOr(&flags[0], 1);
	mx = __reduce_max_sync(__activemask(), mx);
	if ((threadIdx.x & 31) == 0) atomicMax(&flags[5], (int) mx - 33);
}

__global__ void __launch_bounds__(QVZ_THREADS)
qvz_ids_to_lines_kernel(qvz_layout L, const uint8_t *__restrict__ cl, uint8_t *__restrict__ ids)
{
	uint64_t p = (uint64_t) blockIdx.x * QVZ_THREADS + threadIdx.x;
	if (p >= L.P) return;
	uint64_t line = qvz_slot_line(L, p);
	if (line < L.n_lines) ids[line] = cl[p];
}

__global__ void __launch_bounds__(QVZ_THREADS)
qvz_ids_from_lines_kernel(qvz_layout L, const uint8_t *__restrict__ ids, uint8_t *__restrict__ cl)
{
	uint64_t p = (uint64_t) blockIdx.x * QVZ_THREADS + threadIdx.x;
	if (p >= L.P) return;
	uint64_t line = qvz_slot_line(L, p);
	cl[p] = line < L.n_lines ? ids[line] : QVZ_NO_LINE;
}

// Packed words [C4][P] -> line-major bytes (symbol stream, or the `-u` image with '\n' per line).
__global__ void __launch_bounds__(QVZ_THREADS)
qvz_words_to_lines_kernel(qvz_layout L, const uint32_t *__restrict__ Yw, uint8_t *__restrict__ out,
                          uint32_t out_stride, int add_newline)
{
	uint64_t p = (uint64_t) blockIdx.x * QVZ_THREADS + threadIdx.x;
	if (p >= L.P) return;
	uint64_t line = qvz_slot_line(L, p);
	if (line >= L.n_lines) return;
	uint8_t *dst = out + line * (uint64_t) out_stride;
	for (uint32_t c4 = 0; c4 < L.C4; ++c4) {
		uint32_t w = Yw[(uint64_t) c4 * L.P + p];
#pragma unroll
		for (uint32_t j = 0; j < 4; ++j) {
			uint32_t c = 4 * c4 + j;
			if (c < L.C) dst[c] = (uint8_t) (w >> (8 * j));
		}
	}
	if (add_newline) dst[L.C] = '\n';
}

__global__ void __launch_bounds__(QVZ_THREADS)
qvz_doubles_to_lines_kernel(qvz_layout L, const double *__restrict__ Ep, double *__restrict__ out)
{
	uint64_t p = (uint64_t) blockIdx.x * QVZ_THREADS + threadIdx.x;
	if (p >= L.P) return;
	uint64_t line = qvz_slot_line(L, p);
	if (line < L.n_lines) out[line] = Ep[p];
}

static inline unsigned slot_blocks(const qvz_gpu *h) {
	return (unsigned) ((h->L.P + QVZ_THREADS - 1) / QVZ_THREADS);
}

int qvz_layout_ingest(qvz_gpu *h, const uint8_t *raw_dev, uint32_t row_stride) {
	qvz_ingest_kernel<<<slot_blocks(h), QVZ_THREADS, 0, h->stream>>>(h->L, raw_dev, row_stride, h->Xw,
	                                                                 h->cl, h->flags);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}

int qvz_layout_ids_to_lines(qvz_gpu *h, uint8_t *ids_dev) {
	qvz_ids_to_lines_kernel<<<slot_blocks(h), QVZ_THREADS, 0, h->stream>>>(h->L, h->cl, ids_dev);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}

int qvz_layout_ids_from_lines(qvz_gpu *h, const uint8_t *ids_dev) {
	qvz_ids_from_lines_kernel<<<slot_blocks(h), QVZ_THREADS, 0, h->stream>>>(h->L, ids_dev, h->cl);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}

int qvz_layout_words_to_lines(qvz_gpu *h, const uint32_t *Yw, uint8_t *out_dev, uint32_t out_stride,
                              int add_newline) {
	qvz_words_to_lines_kernel<<<slot_blocks(h), QVZ_THREADS, 0, h->stream>>>(h->L, Yw, out_dev, out_stride,
	                                                                         add_newline);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}

int qvz_layout_doubles_to_lines(qvz_gpu *h, const double *Ep, double *out_dev) {
	qvz_doubles_to_lines_kernel<<<slot_blocks(h), QVZ_THREADS, 0, h->stream>>>(h->L, Ep, out_dev);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}
