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
// Tables: U[k][col][prev_qv][hi][data] = state | qv << 8 is the composition
// ctx_of -> (qmap, smap) of `struct qvz_flat_tables`, indexed by the previous quantized VALUE so that the
// context lookup costs no extra dependent load; R[k][col][prev_qv] = qratio or 0xFF if the reference
// would hit its assert (src/codebook.c:164).
#include "qvz_internal.cuh"

__global__ void __launch_bounds__(QVZ_THREADS)
qvz_quantize_kernel(qvz_layout L, const uint32_t *__restrict__ Xw, const uint8_t *__restrict__ cl,
                    const uint16_t *__restrict__ U, const uint8_t *__restrict__ R,
                    const double *__restrict__ D, const uint32_t *__restrict__ run_states,
                    uint32_t *__restrict__ Yw, uint32_t *__restrict__ Qw, double *__restrict__ Ep,
                    int *__restrict__ flags)
{
	__shared__ uint32_t ws[32 * QVZ_THREADS];        // WELL state, word k of thread t at ws[k*256 + t]
	const uint32_t t = threadIdx.x;
	const uint64_t r = (uint64_t) blockIdx.x * QVZ_THREADS + t;     // run index, < T (T % 256 == 0)
#pragma unroll
	for (int k = 0; k < 32; ++k) ws[k * QVZ_THREADS + t] = run_states[r * 32 + k];
	uint32_t n = 0;                                  // ring index: uniform, every thread steps in lockstep
	uint32_t bits = 0, left = 0;
	bool missing = false;

	for (uint32_t i = 0; i < L.Lr; ++i) {
		const uint64_t p = (uint64_t) i * L.T + r;
		const uint32_t kraw = cl[p];
		const bool valid = kraw != QVZ_NO_LINE;
		const uint32_t k = valid ? kraw : 0;
		uint32_t prev = 0;
		double err = 0.0;
		for (uint32_t c4 = 0; c4 < L.C4; ++c4) {
			const uint32_t w = valid ? Xw[(uint64_t) c4 * L.P + p] - 0x21212121u : 0u;
			uint32_t outw = 0, qvw = 0;
#pragma unroll
			for (uint32_t j = 0; j < 4; ++j) {
				const uint32_t col = 4 * c4 + j;
				if (col < L.C) {
					if (left < 7) {                  // well_1024a_bits refill (src/well.c:37-40)
						const uint32_t z0 = ws[((n + 31) & 31) * QVZ_THREADS + t];
						const uint32_t a = ws[((n + 3) & 31) * QVZ_THREADS + t];
						const uint32_t b = ws[((n + 24) & 31) * QVZ_THREADS + t];
						const uint32_t c = ws[((n + 10) & 31) * QVZ_THREADS + t];
						const uint32_t z1 = ws[n * QVZ_THREADS + t] ^ (a ^ (a >> 8));
						const uint32_t z2 = (b ^ (b << 19)) ^ (c ^ (c << 14));
						ws[n * QVZ_THREADS + t] = z1 ^ z2;
						n = (n + 31) & 31;
						bits = (z0 ^ (z0 << 11)) ^ (z1 ^ (z1 << 7)) ^ (z2 ^ (z2 << 13));
						ws[n * QVZ_THREADS + t] = bits;
						left = 32;
					}
					const uint32_t draw = bits & 127u;
					bits >>= 7;
					left -= 7;
					const uint32_t data = (w >> (8 * j)) & 0xFFu;
					const uint32_t row = ((k * L.C + col) * 72u + prev);
					const uint32_t ratio = __ldg(&R[row]);
					missing |= valid && (ratio == 0xFFu);
					const uint32_t hi = draw >= ratio;
					const uint32_t e = __ldg(&U[(uint64_t) (row * 2u + hi) * 72u + data]);
					const uint32_t qv = e >> 8;
					outw |= ((e & 0xFFu) | (hi << 7)) << (8 * j);
					qvw |= (qv + 33u) << (8 * j);
					const double d = __ldg(&D[data + 72u * qv]);
					err = (col == 0) ? d : err + d;  // assignment at column 0, += after (qv_compressor.c:97,118)
					prev = qv;
				}
			}
			Yw[(uint64_t) c4 * L.P + p] = outw;
			if (Qw) Qw[(uint64_t) c4 * L.P + p] = qvw;
		}
		if (Ep) Ep[p] = err / (double) L.C;
	}
	if (missing) atomicOr(&flags[2], 1);
}

int qvz_quantize_launch(qvz_gpu *h, int want_qv, int want_err) {
	const unsigned grid = h->L.T / QVZ_THREADS;
	qvz_quantize_kernel<<<grid, QVZ_THREADS, 0, h->stream>>>(h->L, h->Xw, h->cl, h->U, h->R, h->D,
	                                                         h->run_states, h->Yw, want_qv ? h->Qw : nullptr,
	                                                         want_err ? h->Ep : nullptr, h->flags);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}
