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
// reference's WELL stream: it starts from the jump-ahead state of well.cu and serves 4 draws per word,
// low bits first, exactly like the reference's bit server.  Every run starts on a word boundary, so the
// refill position inside a line is known statically: line i of a run starts at draw i*C, i.e. at
// sub-draw PH = (i*C) & 3 of a word, and a full 4-symbol data word refills before symbol (4-PH)&3.
// All threads are at the same column at the same time, adjacent threads hold adjacent slots => packed
// words are read and written coalesced.
//
// Tables (built by qvz_quantize_compose_kernel from `struct qvz_flat_tables`):
//   W[k][col][prev_qv][data] = qv_lo | qv_hi << 8 | state_lo << 16 | (state_hi | 0x80) << 24
//       the composition ctx_of -> (qmap, smap) for BOTH quantizers of the context, indexed by the previous
//       quantized VALUE: one dependent load per symbol gives both candidates; the output byte
//       (state | hi << 7) is byte 2+hi of the entry;
//   R[k][col][prev_qv]       = qratio, or 0xFF where the reference would hit its assert (codebook.c:164);
//       loaded in parallel with W (same dependence), the lo/hi choice is a byte select afterwards.
// The hot part of W per column is a few KB (a band around prev ~ data), so it lives in L1: the kernel
// asks for a shared-memory carve-out that leaves ~100 KB of L1 and streams the row words past it
// (ld.global.nc.L1::no_allocate / st.global.L1::no_allocate).
//
// Distortion (src/qv_compressor.c:97,118,127): DMODE 2 = the matrix is a function of |x-y| with integer
// values (-d M, -d A): the per-line sum is accumulated in uint32 (every partial sum of the reference's
// double additions is the same exact integer); DMODE 1 = function of |x-y| (-d L): doubles from shared
// memory added in column order (bit-identical to the reference's sequence of additions);
// DMODE 0 = arbitrary 72x72 matrix (-D file) read from global memory.
#include <stdlib.h>

#include <type_traits>

#include "qvz_internal.cuh"

#define QZ_THREADS QVZ_THREADS
#define QZ_BLOCKS_PER_SM 4
#define QZ_WS_WORDS (32 * QZ_THREADS)

__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p) {
	uint32_t v;
	asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ void st_stream_u32(uint32_t *p, uint32_t v) {
	asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v));
}

struct qz_ctx {
	const uint32_t *W;
	const uint8_t *R;
	const double *D;
	uint32_t *ws;          // this thread's column of the WELL ring: ws[k * QZ_THREADS]
	const double *dd;      // shared |x-y| tables
	const uint32_t *di;
	uint32_t n10;          // ring index << 10 (bytes: 4 * QZ_THREADS per ring slot), uniform
	uint32_t bits;         // undrawn part of the current WELL word
	uint32_t prev, wbase, rbase, maxratio;
	uint32_t erri;
	double errd;
};

// well_1024a (src/well.c:8-24) on the shared-memory ring
__device__ __forceinline__ void qz_refill(qz_ctx &q) {
	char *base = (char *) q.ws;
	const uint32_t n10 = q.n10;
	const uint32_t z0 = *(uint32_t *) (base + ((n10 + 31u * 1024u) & 0x7C00u));
	const uint32_t a = *(uint32_t *) (base + ((n10 + 3u * 1024u) & 0x7C00u));
	const uint32_t b = *(uint32_t *) (base + ((n10 + 24u * 1024u) & 0x7C00u));
	const uint32_t c = *(uint32_t *) (base + ((n10 + 10u * 1024u) & 0x7C00u));
	const uint32_t z1 = *(uint32_t *) (base + n10) ^ (a ^ (a >> 8));
	const uint32_t z2 = (b ^ (b << 19)) ^ (c ^ (c << 14));
	*(uint32_t *) (base + n10) = z1 ^ z2;
	q.n10 = (n10 + 31u * 1024u) & 0x7C00u;
	q.bits = (z0 ^ (z0 << 11)) ^ (z1 ^ (z1 << 7)) ^ (z2 ^ (z2 << 13));
	*(uint32_t *) (base + q.n10) = q.bits;
}

// one symbol: byte J of the data word w; returns with outw / qvw byte J filled in
template <int DMODE, bool WANT_QV, int J>
__device__ __forceinline__ void qz_symbol(qz_ctx &q, uint32_t w, uint32_t &outw, uint32_t &qvw) {
	const uint32_t draw = q.bits & 127u;
	q.bits >>= 7;
	const uint32_t data = __byte_perm(w, 0, 0x4440 + J);
	const uint32_t e = __ldg(q.W + (q.wbase + data + q.prev * 72u));
	const uint32_t ratio = __ldg(q.R + (q.rbase + q.prev));
	q.maxratio = max(q.maxratio, ratio);
	const uint32_t hi = draw >= ratio;
	const uint32_t qv = __byte_perm(e, 0, 0x4440 + hi);                    // byte hi of e
	// insert byte 2+hi of e (state | hi<<7) at byte J of outw; byte hi of e (qv) at byte J of qvw
	outw = __byte_perm(outw, e, (0x3210 & ~(0xF << (4 * J))) + ((6 + hi) << (4 * J)));
	if (WANT_QV) qvw = __byte_perm(qvw, e, (0x3210 & ~(0xF << (4 * J))) + ((4 + hi) << (4 * J)));
	if (DMODE == 2) {
		q.erri += q.di[abs((int) data - (int) qv)];
	} else if (DMODE == 1) {
		q.errd += q.dd[abs((int) data - (int) qv)];
	} else {
		q.errd += __ldg(&q.D[data + 72u * qv]);
	}
	q.prev = qv;
	q.wbase += 72u * 72u;
	q.rbase += 72u;
}

// one full data word; PH = sub-draw position of the word's first symbol inside its WELL word
template <int DMODE, bool WANT_QV, int PH>
__device__ __forceinline__ void qz_word(qz_ctx &q, uint32_t w, uint32_t &outw, uint32_t &qvw) {
	if (PH == 0) qz_refill(q);
	qz_symbol<DMODE, WANT_QV, 0>(q, w, outw, qvw);
	if (PH == 3) qz_refill(q);
	qz_symbol<DMODE, WANT_QV, 1>(q, w, outw, qvw);
	if (PH == 2) qz_refill(q);
	qz_symbol<DMODE, WANT_QV, 2>(q, w, outw, qvw);
	if (PH == 1) qz_refill(q);
	qz_symbol<DMODE, WANT_QV, 3>(q, w, outw, qvw);
}

template <int DMODE, bool WANT_QV, int PH>
__device__ __forceinline__ void qz_full_words(qz_ctx &q, const qvz_layout &L, bool valid, const uint32_t *xp,
                                              uint32_t *yp, uint32_t *qp, uint32_t full, uint32_t &wnext) {
	for (uint32_t c4 = 0; c4 < full; ++c4) {
		const uint32_t w = valid ? wnext - 0x21212121u : 0u;
		if (c4 + 1 < L.C4) wnext = ld_stream_u32(xp + (uint64_t) (c4 + 1) * L.P);
		uint32_t outw = 0, qvw = 0;
		qz_word<DMODE, WANT_QV, PH>(q, w, outw, qvw);
		st_stream_u32(yp + (uint64_t) c4 * L.P, outw);
		if (WANT_QV) st_stream_u32(qp + (uint64_t) c4 * L.P, qvw + 0x21212121u);
	}
}

template <int DMODE, bool WANT_QV>
__global__ void __launch_bounds__(QZ_THREADS, QZ_BLOCKS_PER_SM)
qvz_quantize_kernel(qvz_layout L, const uint32_t *__restrict__ Xw, const uint8_t *__restrict__ cl,
                    const uint32_t *__restrict__ W, const uint8_t *__restrict__ R,
                    const double *__restrict__ D, const uint32_t *__restrict__ run_states,
                    uint32_t *__restrict__ Yw, uint32_t *__restrict__ Qw, double *__restrict__ Ep,
                    int *__restrict__ flags)
{
	extern __shared__ __align__(16) uint32_t smem[];  // WELL ring [32][256], then the |x-y| distortion table
	const uint32_t t = threadIdx.x;
	const uint64_t r = (uint64_t) blockIdx.x * QZ_THREADS + t;      // run index, < T (T % 256 == 0)
#pragma unroll
	for (int k = 0; k < 32; ++k) smem[k * QZ_THREADS + t] = run_states[r * 32 + k];
	double *dd = (double *) (smem + QZ_WS_WORDS);
	uint32_t *di = smem + QZ_WS_WORDS;
	if (DMODE == 1 && t < QVZ_ALPHABET) dd[t] = D[t];               // D[x + 72*0] = f(|x - 0|)
	if (DMODE == 2 && t < QVZ_ALPHABET) di[t] = (uint32_t) D[t];
	if (DMODE != 0) __syncthreads();

	qz_ctx q;
	q.W = W;
	q.R = R;
	q.D = D;
	q.ws = smem + t;
	q.dd = dd;
	q.di = di;
	q.n10 = 0;
	q.bits = 0;
	bool missing = false;
	const uint32_t C = L.C, full = C >> 2, rem = C & 3;

	for (uint32_t i = 0; i < L.Lr; ++i) {
		const uint64_t p = (uint64_t) i * L.T + r;
		const uint32_t kraw = cl[p];
		const bool valid = kraw != QVZ_NO_LINE;
		const uint32_t k = valid ? kraw : 0;
		q.wbase = k * C * (72u * 72u);
		q.rbase = k * C * 72u;
		q.prev = 0;
		q.maxratio = 0;
		q.erri = 0;
		q.errd = 0.0;                                // 0.0 + d == d exactly: same bits as "error = d" at column 0
		const uint32_t *xp = Xw + p;
		uint32_t *yp = Yw + p;
		uint32_t *qp = WANT_QV ? Qw + p : nullptr;
		uint32_t wnext = ld_stream_u32(xp);
		const uint32_t ph = (i * C) & 3;             // uniform: where in its WELL word this line starts
		switch (ph) {
		case 0: qz_full_words<DMODE, WANT_QV, 0>(q, L, valid, xp, yp, qp, full, wnext); break;
		case 1: qz_full_words<DMODE, WANT_QV, 1>(q, L, valid, xp, yp, qp, full, wnext); break;
		case 2: qz_full_words<DMODE, WANT_QV, 2>(q, L, valid, xp, yp, qp, full, wnext); break;
		default: qz_full_words<DMODE, WANT_QV, 3>(q, L, valid, xp, yp, qp, full, wnext); break;
		}
		if (rem) {                                   // last, partial word of the line
			const uint32_t w = valid ? wnext - 0x21212121u : 0u;
			uint32_t outw = 0, qvw = 0;
			if (ph == 0) qz_refill(q);
			qz_symbol<DMODE, WANT_QV, 0>(q, w, outw, qvw);
			if (rem > 1) {
				if (ph == 3) qz_refill(q);
				qz_symbol<DMODE, WANT_QV, 1>(q, w, outw, qvw);
			}
			if (rem > 2) {
				if (ph == 2) qz_refill(q);
				qz_symbol<DMODE, WANT_QV, 2>(q, w, outw, qvw);
			}
			st_stream_u32(yp + (uint64_t) full * L.P, outw);
			if (WANT_QV) st_stream_u32(qp + (uint64_t) full * L.P, qvw + 0x21212121u);
		}
		missing |= valid && (q.maxratio == 0xFFu);
		if (Ep) Ep[p] = (DMODE == 2 ? (double) q.erri : q.errd) / (double) C;
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
				e = lo | (hi << 8) | ((uint32_t) (smap[q * 72 + lo] & 0x7F) << 16) |
				    ((uint32_t) ((smap[(q + 1) * 72 + hi] & 0x7F) | 0x80) << 24);
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

template <int DMODE, bool WANT_QV>
static void launch_quantize(qvz_gpu *h, int want_err) {
	auto kern = qvz_quantize_kernel<DMODE, WANT_QV>;
	const size_t smem = QZ_WS_WORDS * sizeof(uint32_t) + QVZ_ALPHABET * sizeof(double);
	// leave ~100 KB of L1 for the table band: 4 CTAs x ~33 KB of shared memory
	cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 60);
	kern<<<h->L.T / QZ_THREADS, QZ_THREADS, smem, h->stream>>>(h->L, h->Xw, h->cl, h->W, h->R, h->D, h->run_states, h->Yw,
	                                                           WANT_QV ? h->Qw : nullptr, want_err ? h->Ep : nullptr, h->flags);
}

// dmode: 0 arbitrary matrix, 1 function of |x-y|, 2 integer-valued function of |x-y|
int qvz_quantize_launch(qvz_gpu *h, int want_qv, int want_err, int dmode) {
	if (want_qv) {
		if (dmode == 2) launch_quantize<2, true>(h, want_err);
		else if (dmode == 1) launch_quantize<1, true>(h, want_err);
		else launch_quantize<0, true>(h, want_err);
	} else {
		if (dmode == 2) launch_quantize<2, false>(h, want_err);
		else if (dmode == 1) launch_quantize<1, false>(h, want_err);
		else launch_quantize<0, false>(h, want_err);
	}
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}

// =====================================================================================================
// Batched walk (the fast path).
//
// The line-major walk above keeps one WELL state per line in flight and its table loads go to L1/L2.  The
// batched path splits the work in two kernels:
//
//   qvz_draws_kernel      thread <-> run.  The WELL1024a state lives in 32 REGISTERS: the generator is
//                         unrolled over one full turn of the ring (32 steps, n returns to 0) so that every
//                         state index is a compile-time constant; each word is expanded to one byte per
//                         7-bit draw in the reference's bit server order (4 draws per word, low bits first,
//                         src/well.c:33-46) and stored in SEQUENCE order, Dw[w*T + run] (coalesced: adjacent
//                         threads are adjacent runs).  Line i of a run starts at draw i*C: the walk realigns.
//   qvz_quantize_batched  one CTA walks QB_LINES slots column-synchronously.  The (cluster, column) tables
//                         are COMPACTED to the A x A box of values that can occur (A-1 = largest symbol /
//                         quantized value present) and staged per column group in shared memory by TMA bulk
//                         copies (double buffered, mbarrier completion), so every lookup is one LDS.64;
//                         each thread carries QB_LPT independent lines (ILP across the dependent
//                         prev -> lookup -> prev chains).
//
// Compact table image, one per column (G[col] = two planes of K*A*A words: lo quantizers, then hi quantizers):
//   plane[hi][k][prev][data] = variant of that quantizer (the lo/hi choice is known BEFORE the lookup, so only
//   the chosen 4-byte variant is loaded: one shared-memory wavefront when the lanes' words fall in distinct banks),   variant =
//       byte 0  qv       the quantized value = the next column's context value (q->q[data])
//       byte 1  |data-qv| index of the distortion table (72 = poison: NaN / 2^31, see below)
//       byte 2  state | hi<<7      the output symbol (q->output_alphabet->indexes[qv], quantizer choice)
//       byte 3  qratio of the NEXT column's context qv: the draw comparison of the next symbol is
//               (draw<<24 | 0xFFFFFF) >= variant, with no field extraction at all.
//   A context the reference would assert on (src/codebook.c:164) has poisoned entries: the line's error sum
//   becomes NaN (double modes) or >= 2^31 (integer mode) and is reported once per line instead of being
//   checked per symbol.
// =====================================================================================================
#ifndef QB_THREADS
#define QB_THREADS 1024                 // threads per CTA of the walk (tuning: -DQB_THREADS=512 -DQB_CTAS=2)
#endif
#ifndef QB_CTAS
#define QB_CTAS 1                       // resident walk CTAs per SM
#endif
#define QB_LPT 4
#define QB_LINES (QB_THREADS * QB_LPT)
static_assert(QVZ_RUN_ALIGN % QB_LINES == 0, "a step (T slots) must be a whole number of walk batches");
#define QB_POISON 72u

// ---- draw generator ------------------------------------------------------------------------------------
// one well_1024a step (src/well.c:8-24) at ring position n = (32 - T) & 31, state in registers
template <int T>
__device__ __forceinline__ uint32_t well_step_reg(uint32_t (&s)[32]) {
	constexpr int n = (32 - T) & 31;
	const uint32_t z0 = s[(n + 31) & 31], a = s[(n + 3) & 31], b = s[(n + 24) & 31], c = s[(n + 10) & 31];
	const uint32_t z1 = s[n] ^ (a ^ (a >> 8));
	const uint32_t z2 = (b ^ (b << 19)) ^ (c ^ (c << 14));
	s[n] = z1 ^ z2;
	const uint32_t out = (z0 ^ (z0 << 11)) ^ (z1 ^ (z1 << 7)) ^ (z2 ^ (z2 << 13));
	s[(n + 31) & 31] = out;
	return out;
}

// the 4 draws of one word (bits 0-6, 7-13, 14-20, 21-27; top 4 bits dropped) -> one byte each
__device__ __forceinline__ uint32_t well_expand_draws(uint32_t w) {
	w &= 0x0FFFFFFFu;
	const uint32_t u = (w & 0x0FFFC000u) * 3u + w;       // draws 2, 3 move up by 2 bits
	return u + (u & 0x3F803F80u);                        // draws 1, 3 move up by 1 more bit
}

// One turn of the ring: 32 steps, 32 draw words.  Dw holds the draws of a run in SEQUENCE order -- word w of run r
// (draws 4w .. 4w+3 after the run's first draw) at Dw[w*T + r] -- so the generator does no per-line bookkeeping at
// all (adjacent threads = adjacent runs: coalesced stores); the walk realigns on read, where a line starts at draw
// i*C of its run, i.e. at byte (i*C)&3 of word (i*C)>>2.  GUARD = the last, partial turn of the run.
template <int T, bool GUARD>
__device__ __forceinline__ void well_turn_seq(uint32_t (&s)[32], uint32_t *&dp, uint64_t stride, uint32_t left) {
	if constexpr (T < 32) {
		const uint32_t w = well_expand_draws(well_step_reg<T>(s));
		if (!GUARD || (uint32_t) T < left) st_stream_u32(dp, w);
		dp += stride;
		well_turn_seq<T + 1, GUARD>(s, dp, stride, left);
	}
}

__global__ void __launch_bounds__(QZ_THREADS, 4)
qvz_draws_kernel(qvz_layout L, const uint32_t *__restrict__ run_states, uint32_t *__restrict__ Dw)
{
	const uint64_t r = (uint64_t) blockIdx.x * QZ_THREADS + threadIdx.x;
	uint32_t s[32];
#pragma unroll
	for (int k = 0; k < 32; ++k) s[k] = run_states[r * 32 + k];
	const uint32_t words = (uint32_t) ((uint64_t) L.Lr * L.C / 4);       // Lr % 4 == 0: whole words
	uint32_t *dp = Dw + r;
	for (uint32_t t = 0; t < words / 32; ++t) well_turn_seq<0, false>(s, dp, L.T, 0);
	if (words & 31) well_turn_seq<0, true>(s, dp, L.T, words & 31);
}

// largest quantized value any present context can emit for a data value <= smax  -> flags[4]
__global__ void __launch_bounds__(256)
qvz_quantize_vmax_kernel(uint64_t entries, const uint32_t *__restrict__ W, const uint8_t *__restrict__ R,
                         uint32_t smax, int *__restrict__ vmax)
{
	const uint64_t idx = (uint64_t) blockIdx.x * 256 + threadIdx.x;
	uint32_t m = 0;
	if (idx < entries) {
		const uint32_t x = idx % 72;
		if (x <= smax && R[idx / 72] != 0xFF) {
			const uint32_t e = W[idx];
			m = max(e & 0xFFu, (e >> 8) & 0xFFu);
		}
	}
	m = __reduce_max_sync(0xFFFFFFFFu, m);
	if ((threadIdx.x & 31) == 0 && m) atomicMax(vmax, (int) m);
}

// full 72x72 tables -> one contiguous image per COLUMN: G[col] = entry[K][A][A] (see the format above)
__global__ void __launch_bounds__(256)
qvz_quantize_compact_kernel(uint32_t K, uint32_t C, uint32_t A, const uint32_t *__restrict__ W,
                            const uint8_t *__restrict__ R, uint32_t *__restrict__ G)
{
	const uint64_t idx = (uint64_t) blockIdx.x * 256 + threadIdx.x;
	if (idx >= (uint64_t) K * C * A * A) return;
	const uint32_t x = idx % A, v = (idx / A) % A;
	const uint64_t kc = idx / ((uint64_t) A * A);
	const uint32_t k = kc / C, col = kc - (uint64_t) k * C;
	uint32_t var[2];
	if (R[kc * 72 + v] == 0xFF) {                    // no such context: poison
		var[0] = (QB_POISON << 8) | (0x7Fu << 16);
		var[1] = (QB_POISON << 8) | (0xFFu << 16);
	} else {
		const uint32_t e = W[(kc * 72 + v) * 72 + x];
#pragma unroll
		for (uint32_t hi = 0; hi < 2; ++hi) {
			const uint32_t qv = (e >> (8 * hi)) & 0xFFu, st = (e >> (16 + 8 * hi)) & 0xFFu;
			uint32_t nr = 0;
			if (col + 1 < C) {
				nr = R[(kc + 1) * 72 + qv];
				if (nr == 0xFF) nr = 1;              // the poisoned row of the next column reports it
			}
			nr = (nr - 1u) & 0xFFu;
			const uint32_t d = x > qv ? x - qv : qv - x;
			var[hi] = qv | (d << 8) | (st << 16) | (nr << 24);
		}
	}
	uint32_t *g = (uint32_t *) G + (size_t) col * 2 * K * A * A + ((size_t) k * A + v) * A + x;
	g[0] = var[0];
	g[(size_t) K * A * A] = var[1];
}

// ---- TMA bulk copy + mbarrier (sm_90+/sm_100a): one elected thread moves a whole column-group image
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "WAIT_%=:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra DONE_%=;\n"
	    "bra WAIT_%=;\n"
	    "DONE_%=:\n"
	    "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
	double v;
	asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}

// S = columns staged per barrier (4, 2 or 1: the largest whose double buffer fits shared memory)
// ONEK = one cluster (the reference's default): no per-line table offset
template <int DMODE, bool WANT_QV, int S, bool ONEK>
__global__ void __launch_bounds__(QB_THREADS, QB_CTAS)
qvz_quantize_batched_kernel(qvz_layout L, const uint32_t *__restrict__ Xw, const uint32_t *__restrict__ Dw,
                            const uint8_t *__restrict__ cl, const uint8_t *__restrict__ G,
                            const uint8_t *__restrict__ R, const double *__restrict__ D, uint32_t K,
                            uint32_t A, uint32_t *__restrict__ Yw, uint32_t *__restrict__ Qw,
                            double *__restrict__ Ep, int *__restrict__ flags)
{
	extern __shared__ __align__(16) uint32_t smem[];
	// [dd: 73 doubles (or 73 words)][2 mbarriers][buffer 0][buffer 1]; a buffer = S consecutive column images of G
	constexpr uint32_t DD_WORDS = 2 * (QVZ_ALPHABET + 2);
	double *dd = (double *) smem;
	uint32_t *di = smem;
	uint64_t *full = (uint64_t *) (smem + DD_WORDS);
	const uint32_t col_bytes = K * A * A * 8;            // multiple of 32 (A is even)
	const uint32_t buf_bytes = S * col_bytes;
	const uint32_t buf0 = smem_u32(smem + DD_WORDS + 4);
	const uint32_t dd_addr = smem_u32(smem);
	const uint32_t tid = threadIdx.x;
	if (DMODE == 1 && tid <= QVZ_ALPHABET) dd[tid] = tid < QVZ_ALPHABET ? D[tid] : __longlong_as_double(0x7FF8000000000000ll);
	if (DMODE == 2 && tid <= QVZ_ALPHABET) di[tid] = tid < QVZ_ALPHABET ? (uint32_t) D[tid] : 0x80000000u;
	if (tid == 0) {
		mbar_init(&full[0], 1);
		mbar_init(&full[1], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	const uint32_t C = L.C, C4 = L.C4;
	const uint32_t A4 = A * 4;
	const uint32_t plane = K * A * A4;                   // bytes from a lo variant to its hi variant
	uint32_t gcount = 0;                             // column groups consumed so far (uniform): buffer = gcount & 1
	auto stage = [&](uint32_t col0, uint32_t g) {    // thread 0: columns col0 .. min(col0+S, C)-1 -> buffer g & 1
		const uint32_t ncol = (C - col0 < (uint32_t) S) ? C - col0 : (uint32_t) S;
		const uint32_t bytes = ncol * col_bytes;
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of this buffer are done
		mbar_expect_tx(&full[g & 1], bytes);
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		             ::"r"(buf0 + (g & 1) * buf_bytes), "l"(G + (uint64_t) col0 * col_bytes), "r"(bytes), "r"(smem_u32(&full[g & 1])) : "memory");
	};

	bool missing = false;
	// A batch = QB_LINES consecutive slots of ONE step i (slot = i*T + run; T % QB_LINES == 0, abi.cu): all its lines start
	// at the same draw i*C of their runs, so the position in the draw stream is uniform across the CTA.
	const uint32_t bps = L.T / QB_LINES;                 // batches per step
	const uint64_t nbatch = (uint64_t) bps * L.Lr;
	for (uint64_t batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
		const uint32_t step = (uint32_t) (batch / bps), boff = (uint32_t) (batch - (uint64_t) step * bps) * QB_LINES;
		const uint64_t pbase = (uint64_t) step * L.T + boff + tid;
		const uint32_t d0 = step * C;                    // first draw of these lines within their runs
		const uint32_t dsh = 8 * (d0 & 3);               // ... = byte d0 & 3 of word d0 >> 2
		uint32_t koff[QB_LPT], vprev[QB_LPT], erri[QB_LPT];
		double errd[QB_LPT];
		bool valid[QB_LPT];
#pragma unroll
		for (int j = 0; j < QB_LPT; ++j) {
			const uint32_t kraw = cl[pbase + j * QB_THREADS];
			valid[j] = kraw != QVZ_NO_LINE;
			const uint32_t k = valid[j] ? kraw : 0;      // a slot without a line walks cluster 0's tables on zero data
			koff[j] = ONEK ? 0u : k * A * A4;
			// column 0: previous value 0 (src/qv_compressor.c:89), its ratio is the first byte of the cluster's R rows
			vprev[j] = ((uint32_t) __ldg(R + (size_t) k * C * 72) - 1u) << 24;
			erri[j] = 0;
			errd[j] = 0.0;
		}
		if (tid == 0) stage(0, gcount);
		uint32_t xn[QB_LPT], da[QB_LPT], db[QB_LPT];     // draw words w, w+1 of the current step: da is the lower one in even steps, db in odd steps
		const uint32_t *xr = Xw + pbase;                 // running pointer: one word column (P slots) per step
		const uint32_t *drw = Dw + (uint64_t) (d0 >> 2) * L.T + (boff + tid);      // one draw word (T runs) per step
		uint32_t *yr = Yw + pbase, *qr = WANT_QV ? Qw + pbase : nullptr;
#pragma unroll
		for (int j = 0; j < QB_LPT; ++j) {
			xn[j] = ld_stream_u32(xr + j * QB_THREADS);
			da[j] = ld_stream_u32(drw + j * QB_THREADS);
			db[j] = ld_stream_u32(drw + L.T + j * QB_THREADS);      // Dw has spare word rows at the end (abi.cu)
		}
		drw += 2 * (uint64_t) L.T;
		// one data word (4 columns) of the QB_LPT lines of this thread; TAIL = the last, partial word
		// ODD alternates from word to word: the draw word that was the lower one is dead after the realignment and
		// receives the load for the next step, so no loaded value is ever copied (a copy would wait for the load)
		auto word = [&](uint32_t c4, auto tail_tag, auto odd_tag) {
			constexpr bool TAIL = decltype(tail_tag)::value;
			constexpr bool ODD = decltype(odd_tag)::value;
			uint32_t (&dlo)[QB_LPT] = ODD ? db : da;
			uint32_t (&dhi)[QB_LPT] = ODD ? da : db;
			uint32_t x[QB_LPT], dr[QB_LPT], outw[QB_LPT], qvw[QB_LPT];
#pragma unroll
			for (int j = 0; j < QB_LPT; ++j) {
				// raw ASCII bytes index the tables directly: the -33 is folded into the table base below.  A slot without
				// a line (zero words) walks symbol 0 so that every lookup stays inside the tables; nothing of it is kept.
				x[j] = valid[j] ? xn[j] : 0x21212121u;
				dr[j] = __funnelshift_r(dlo[j], dhi[j], dsh);    // draws d0 + 4*c4 .. +3 of the run
				outw[j] = 0;
				qvw[j] = 0;
			}
			if (!TAIL) {                                 // next word's rows and draws: in flight during this word
				xr += L.P;
#pragma unroll
				for (int j = 0; j < QB_LPT; ++j) {
					xn[j] = ld_stream_u32(xr + j * QB_THREADS);
					dlo[j] = ld_stream_u32(drw + j * QB_THREADS);
				}
				drw += L.T;
			}
#pragma unroll
			for (int g = 0; g < 4 / S; ++g) {            // the column groups (= staged images) inside this word
				const uint32_t col0 = 4 * c4 + g * S;
				if (!TAIL || col0 < C) {
					__syncthreads();                     // everyone is done with the previous group: its buffer is free
					if (tid == 0 && col0 + S < C) stage(col0 + S, gcount + 1);
					mbar_wait(&full[gcount & 1], (gcount >> 1) & 1);   // this group's image has landed
					const uint32_t tabg = buf0 + (gcount & 1) * buf_bytes;
					gcount += 1;
#pragma unroll
					for (int sidx = 0; sidx < S; ++sidx) {
						const int b = g * S + sidx;
						if (!TAIL || col0 + sidx < C) {
							const uint32_t tab = tabg + sidx * col_bytes - 33u * 4u;       // indexed by the raw byte ('!' + value)
#pragma unroll
							for (int j = 0; j < QB_LPT; ++j) {
								// draw >= qratio  <=>  (int)(draw << 24) > (int)((qratio-1) << 24 | low bits of the previous variant)
								const int drawt = (int) __byte_perm(dr[j], 0, 0x0444 + (b << 12));
								const uint32_t base = (drawt > (int) vprev[j]) ? tab + plane : tab;
								const uint32_t data = __byte_perm(x[j], 0, 0x4440 + b);
								const uint32_t row = ONEK ? (vprev[j] & 0x7Fu) * A4 : (vprev[j] & 0x7Fu) * A4 + koff[j];
								const uint32_t v = lds_u32(base + row + data * 4);
								outw[j] = __byte_perm(outw[j], v, (0x3210 & ~(0xF << (4 * b))) + (6 << (4 * b)));
								if (WANT_QV) qvw[j] = __byte_perm(qvw[j], v, (0x3210 & ~(0xF << (4 * b))) + (4 << (4 * b)));
								const uint32_t didx = __byte_perm(v, 0, 0x4441);
								if (DMODE == 2) erri[j] += lds_u32(dd_addr + didx * 4);
								else if (DMODE == 1) errd[j] += lds_f64(dd_addr + didx * 8);
								else if (valid[j]) {
									if (didx == QB_POISON) missing = true;
									errd[j] += __ldg(&D[(data - 33u) + 72u * (v & 0x7Fu)]);
								}
								vprev[j] = v;
							}
						}
					}
				}
			}
#pragma unroll
			for (int j = 0; j < QB_LPT; ++j) st_stream_u32(yr + j * QB_THREADS, outw[j]);
			yr += L.P;
			if (WANT_QV) {
#pragma unroll
				for (int j = 0; j < QB_LPT; ++j) st_stream_u32(qr + j * QB_THREADS, (qvw[j] & 0x7F7F7F7Fu) + 0x21212121u);
				qr += L.P;
			}
		};
		uint32_t c4 = 0;
		for (; c4 + 2 < C4; c4 += 2) {
			word(c4, std::false_type{}, std::false_type{});
			word(c4 + 1, std::false_type{}, std::true_type{});
		}
		if (c4 + 1 < C4) {
			word(c4, std::false_type{}, std::false_type{});
			word(c4 + 1, std::true_type{}, std::true_type{});
		} else word(c4, std::true_type{}, std::false_type{});
#pragma unroll
		for (int j = 0; j < QB_LPT; ++j) {
			if (DMODE == 2) missing |= valid[j] && erri[j] >= 0x80000000u;
			if (DMODE == 1) missing |= valid[j] && errd[j] != errd[j];
			Ep[pbase + j * QB_THREADS] = (DMODE == 2 ? (double) erri[j] : errd[j]) / (double) C;
		}
		__syncthreads();                             // the next batch restages into the other buffer's predecessor
	}
	if (missing) atomicOr(&flags[2], 1);
}

static size_t batched_smem(uint32_t K, uint32_t A, uint32_t S) {
	return 2 * (QVZ_ALPHABET + 2) * sizeof(uint32_t) + 16 + 2 * (size_t) S * ((size_t) K * A * A * 8);
}

// columns staged per barrier: the largest of 4, 2, 1 whose double buffer fits; 0 = batched path unusable
uint32_t qvz_quantize_batched_group(uint32_t K, uint32_t A) {
	for (uint32_t S = 4; S >= 1; S >>= 1)
		if (batched_smem(K, A, S) <= (QB_CTAS == 1 ? 200 : 108) * 1024) return S;
	return 0;
}

int qvz_quantize_draws(qvz_gpu *h) {
	qvz_draws_kernel<<<h->L.T / QZ_THREADS, QZ_THREADS, 0, h->stream>>>(h->L, h->run_states, h->Dw);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}

int qvz_quantize_vmax(qvz_gpu *h, uint32_t KC, uint32_t smax) {
	const uint64_t entries = (uint64_t) KC * 72 * 72;
	qvz_quantize_vmax_kernel<<<(unsigned) ((entries + 255) / 256), 256, 0, h->stream>>>(entries, h->W, h->R, smax, h->flags + 4);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}

int qvz_quantize_compact(qvz_gpu *h, uint32_t K, uint32_t C, uint32_t A) {
	const uint64_t total = (uint64_t) K * C * A * A;
	qvz_quantize_compact_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, h->stream>>>(K, C, A, h->W, h->R, (uint32_t *) h->G);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}

template <int DMODE, bool WANT_QV, int S, bool ONEK>
static void launch_batched(qvz_gpu *h, uint32_t K, uint32_t A) {
	auto kern = qvz_quantize_batched_kernel<DMODE, WANT_QV, S, ONEK>;
	const size_t smem = batched_smem(K, A, S);
	cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	const uint64_t nbatch = (uint64_t) (h->L.T / QB_LINES) * h->L.Lr;     // T % QB_LINES == 0 (QVZ_RUN_ALIGN)
	const uint64_t resident = (uint64_t) h->sm_count * QB_CTAS;
	const unsigned grid = (unsigned) (nbatch < resident ? nbatch : resident);
	kern<<<grid, QB_THREADS, smem, h->stream>>>(h->L, h->Xw, h->Dw, h->cl, h->G, h->R, h->D, K, A, h->Yw,
	                                            WANT_QV ? h->Qw : nullptr, h->Ep, h->flags);
}

template <int DMODE, bool WANT_QV>
static void launch_batched_s(qvz_gpu *h, uint32_t K, uint32_t A, uint32_t S) {
	if (K == 1 && S == 4) launch_batched<DMODE, WANT_QV, 4, true>(h, K, A);
	else if (K == 1 && S == 2) launch_batched<DMODE, WANT_QV, 2, true>(h, K, A);
	else if (S == 4) launch_batched<DMODE, WANT_QV, 4, false>(h, K, A);
	else if (S == 2) launch_batched<DMODE, WANT_QV, 2, false>(h, K, A);
	else launch_batched<DMODE, WANT_QV, 1, false>(h, K, A);
}

int qvz_quantize_launch_batched(qvz_gpu *h, uint32_t K, uint32_t A, int want_qv, int dmode) {
	const uint32_t S = qvz_quantize_batched_group(K, A);
	if (want_qv) {
		if (dmode == 2) launch_batched_s<2, true>(h, K, A, S);
		else if (dmode == 1) launch_batched_s<1, true>(h, K, A, S);
		else launch_batched_s<0, true>(h, K, A, S);
	} else {
		if (dmode == 2) launch_batched_s<2, false>(h, K, A, S);
		else if (dmode == 1) launch_batched_s<1, false>(h, K, A, S);
		else launch_batched_s<0, false>(h, K, A, S);
	}
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}
