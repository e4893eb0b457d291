// quantize.cu -- the per-line quantization walk.
//
// Reference: the per-line loop of start_qv_compression (src/qv_compressor.c:76-135):
//     q = choose_quantizer(qlist, well, col, prev_qv, &idx)    src/codebook.c:162-171
//           ctx = input_alphabets[col]->indexes[prev_qv];  draw = well_1024a_bits(well, 7)  (src/well.c:33-46)
//           hi  = draw >= qratio[col][ctx];  idx = 2*ctx + hi
//     qv = q->q[data];  q_state = q->output_alphabet->indexes[qv];  error (+)= dist[data + 72*qv]
// minus the arithmetic-coder calls (:86, :96, :117), which the host makes afterwards from the emitted
// (state, hi) stream -- legal because the coder never feeds back into the quantizer choice.
//
// One thread walks one run of Lr consecutive lines, so its draws are one contiguous piece of the
// reference's WELL stream: it starts from the jump-ahead state of well.cu and then runs the reference's
// own bit server (refill when fewer than 7 bits are left).  All threads are at the same column at the
// same time, adjacent threads hold adjacent slots => packed words are read and written coalesced.
//
// Tables (built by qvz_quantize_compose_kernel from `struct qvz_flat_tables`):
//   W[k][col][prev_qv][data] = qv_lo | qv_hi << 8 | state_lo << 16 | state_hi << 24
//       the composition ctx_of -> (qmap, smap) for BOTH quantizers of the context, indexed by the previous
//       quantized VALUE: one dependent load per symbol gives both candidates;
//   R[k][col][prev_qv]       = qratio, or 0xFF where the reference would hit its assert (codebook.c:164);
//       loaded in parallel with W (same dependence), the lo/hi choice is a byte select afterwards.
// The hot part of W per column is a few KB (a band around prev ~ data), so it lives in L1: the kernel
// asks for a shared-memory carve-out that leaves ~100 KB of L1 and streams the row words past it
// (ld.global.nc.L1::no_allocate / st.global.L1::no_allocate).
#include "qvz_internal.cuh"

#define QZ_THREADS QVZ_THREADS
#define QZ_BLOCKS_PER_SM 4

__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p) {
	uint32_t v;
	asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ void st_stream_u32(uint32_t *p, uint32_t v) {
	asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v));
}

template <bool TOEPLITZ>
__global__ void __launch_bounds__(QZ_THREADS, QZ_BLOCKS_PER_SM)
qvz_quantize_kernel(qvz_layout L, const uint32_t *__restrict__ Xw, const uint8_t *__restrict__ cl,
                    const uint32_t *__restrict__ W, const uint8_t *__restrict__ R,
                    const double *__restrict__ D, const uint32_t *__restrict__ run_states,
                    uint32_t *__restrict__ Yw, uint32_t *__restrict__ Qw, double *__restrict__ Ep,
                    int *__restrict__ flags)
{
	__shared__ uint32_t ws[32 * QZ_THREADS];         // WELL state, word k of thread t at ws[k*256 + t]
	__shared__ double dd[QVZ_ALPHABET];              // distortion as a function of |x - y| (TOEPLITZ only)
	const uint32_t t = threadIdx.x;
	const uint64_t r = (uint64_t) blockIdx.x * QZ_THREADS + t;      // run index, < T (T % 256 == 0)
#pragma unroll
	for (int k = 0; k < 32; ++k) ws[k * QZ_THREADS + t] = run_states[r * 32 + k];
	if (TOEPLITZ) {
		if (t < QVZ_ALPHABET) dd[t] = D[t];          // D[x + 72*0] = f(|x|)
		__syncthreads();
	}
	uint32_t n = 0;                                  // ring index: uniform, every thread steps in lockstep
	uint32_t bits = 0, left = 0;
	bool missing = false;
	const uint32_t C = L.C, C4 = L.C4;

	for (uint32_t i = 0; i < L.Lr; ++i) {
		const uint64_t p = (uint64_t) i * L.T + r;
		const uint32_t kraw = cl[p];
		const bool valid = kraw != QVZ_NO_LINE;
		const uint32_t k = valid ? kraw : 0;
		const uint32_t *Wc = W + (uint64_t) k * C * (72u * 72u);     // advances by one column table per symbol
		const uint8_t *Rc = R + (uint64_t) k * C * 72u;
		const uint32_t *xp = Xw + p;
		uint32_t *yp = Yw + p;
		uint32_t *qp = Qw ? Qw + p : nullptr;
		uint32_t prev = 0;
		double err = 0.0;                            // 0.0 + d == d exactly: same bits as "error = d" at column 0
		uint32_t wnext = ld_stream_u32(xp);
		for (uint32_t c4 = 0; c4 < C4; ++c4) {
			const uint32_t w = valid ? wnext - 0x21212121u : 0u;
			if (c4 + 1 < C4) wnext = ld_stream_u32(xp + (uint64_t) (c4 + 1) * L.P);
			uint32_t outw = 0, qvw = 0;
#pragma unroll
			for (uint32_t j = 0; j < 4; ++j) {
				if (4 * c4 + j < C) {
					if (left < 7) {                  // well_1024a_bits refill (src/well.c:37-40)
						const uint32_t z0 = ws[((n + 31) & 31) * QZ_THREADS + t];
						const uint32_t a = ws[((n + 3) & 31) * QZ_THREADS + t];
						const uint32_t b = ws[((n + 24) & 31) * QZ_THREADS + t];
						const uint32_t c = ws[((n + 10) & 31) * QZ_THREADS + t];
						const uint32_t z1 = ws[n * QZ_THREADS + t] ^ (a ^ (a >> 8));
						const uint32_t z2 = (b ^ (b << 19)) ^ (c ^ (c << 14));
						ws[n * QZ_THREADS + t] = z1 ^ z2;
						n = (n + 31) & 31;
						bits = (z0 ^ (z0 << 11)) ^ (z1 ^ (z1 << 7)) ^ (z2 ^ (z2 << 13));
						ws[n * QZ_THREADS + t] = bits;
						left = 32;
					}
					const uint32_t draw = bits & 127u;
					bits >>= 7;
					left -= 7;
					const uint32_t data = (w >> (8 * j)) & 0xFFu;
					// the two loads below depend only on prev (and on data, known long before)
					const uint32_t e = __ldg(Wc + prev * 72u + data);
					const uint32_t ratio = __ldg(Rc + prev);
					missing |= valid && (ratio == 0xFFu);
					const uint32_t hi = draw >= ratio;
					const uint32_t qv = __byte_perm(e, 0, hi) & 0xFFu;           // byte hi of e
					const uint32_t st = __byte_perm(e, 0, 2 + hi) & 0xFFu;       // byte 2+hi of e
					outw |= (st | (hi << 7)) << (8 * j);
					qvw |= (qv + 33u) << (8 * j);
					if (TOEPLITZ) {
						const int df = (int) data - (int) qv;
						err += dd[df < 0 ? -df : df];
					} else {
						err += __ldg(&D[data + 72u * qv]);
					}
					prev = qv;
					Wc += 72u * 72u;
					Rc += 72u;
				}
			}
			st_stream_u32(yp + (uint64_t) c4 * L.P, outw);
			if (qp) st_stream_u32(qp + (uint64_t) c4 * L.P, qvw);
		}
		if (Ep) Ep[p] = err / (double) C;
	}
	if (missing) atomicOr(&flags[2], 1);
}

// Compose the flat tables on the device: one thread per (kc, prev value, data value).
__global__ void __launch_bounds__(256)
qvz_quantize_compose_kernel(uint32_t KC, const uint32_t *__restrict__ nctx, const uint8_t *__restrict__ ctx_of,
                            const unsigned long long *__restrict__ q_off, const uint8_t *__restrict__ qratio,
                            const uint8_t *__restrict__ qmap, const uint8_t *__restrict__ smap,
                            uint32_t *__restrict__ W, uint8_t *__restrict__ R, int *__restrict__ flags)
{
	const uint64_t idx = (uint64_t) blockIdx.x * 256 + threadIdx.x;
	if (idx >= (uint64_t) KC * 72 * 72) return;
	const uint32_t x = idx % 72, v = (idx / 72) % 72;
	const uint64_t kc = idx / (72 * 72);
	const uint32_t ctx = ctx_of[kc * 72 + v];
	uint32_t e = 0;
	if (ctx != QVZ_CTX_ABSENT) {
		if (ctx >= nctx[kc] || nctx[kc] > 72) {
			atomicOr(&flags[3], 1);
		} else {
			const uint64_t q = q_off[kc] + 2 * ctx;
			const uint32_t lo = qmap[q * 72 + x], hi = qmap[(q + 1) * 72 + x];
			if (lo >= 72 || hi >= 72) {
				atomicOr(&flags[3], 1);
			} else {
				e = lo | (hi << 8) | ((uint32_t) (smap[q * 72 + lo] & 0x7F) << 16) | ((uint32_t) (smap[(q + 1) * 72 + hi] & 0x7F) << 24);
			}
			if (x == 0) R[kc * 72 + v] = qratio[q_off[kc] / 2 + ctx];
		}
	} else if (x == 0) {
		R[kc * 72 + v] = 0xFF;
	}
	W[idx] = e;
}

int qvz_quantize_compose(qvz_gpu *h, uint32_t KC, const uint32_t *nctx, const uint8_t *ctx_of, const uint64_t *q_off,
                         const uint8_t *qratio, const uint8_t *qmap, const uint8_t *smap) {
	const uint64_t total = (uint64_t) KC * 72 * 72;
	qvz_quantize_compose_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, h->stream>>>(
	    KC, nctx, ctx_of, (const unsigned long long *) q_off, qratio, qmap, smap, h->W, h->R, h->flags);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}

int qvz_quantize_launch(qvz_gpu *h, int want_qv, int want_err, int toeplitz) {
	const unsigned grid = h->L.T / QZ_THREADS;
	// leave ~100 KB of L1 for the table band: 4 CTAs x 33 KB of shared memory
	const int carve = 60;
	if (toeplitz) {
		cudaFuncSetAttribute(qvz_quantize_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
		qvz_quantize_kernel<true><<<grid, QZ_THREADS, 0, h->stream>>>(h->L, h->Xw, h->cl, h->W, h->R, h->D, h->run_states, h->Yw,
		                                                              want_qv ? h->Qw : nullptr, want_err ? h->Ep : nullptr, h->flags);
	} else {
		cudaFuncSetAttribute(qvz_quantize_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
		qvz_quantize_kernel<false><<<grid, QZ_THREADS, 0, h->stream>>>(h->L, h->Xw, h->cl, h->W, h->R, h->D, h->run_states, h->Yw,
		                                                               want_qv ? h->Qw : nullptr, want_err ? h->Ep : nullptr, h->flags);
	}
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}
