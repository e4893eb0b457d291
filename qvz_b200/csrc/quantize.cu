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
//   qvz_quantize_batched  one CTA walks QB_LINES slots column-synchronously.  The tables of a column are ONE
//                         image of `rows` x A entries per plane, staged in shared memory by TMA bulk copies
//                         (a ring of buffers, full/empty mbarriers); each thread carries QB_LPT independent lines
//                         (ILP across the dependent prev -> lookup -> prev chains).  When no reachable context mixes
//                         its two quantizers (DRAWS = false) the draws are neither generated nor read.
//
// Table image of a column (G[col] = two planes, lo quantizers then hi quantizers, of rows x A words):
//   a ROW is one (cluster, context) pair that some line of the resident rows can reach in this column (row 0 is
//   the poison row); rows are numbered per column by qvz_quantize_rows_kernel from a forward reachability pass
//   over the tables, seeded with the data values that occur per (cluster, column) when the counting stage has
//   seen them (cond_counts.cu: support masks) -- a column's image holds what the walk can touch, not 72 x 72 x K.
//   plane[hi][row][data] = variant of that quantizer for that input (the lo/hi choice is known BEFORE the
//   lookup, so only the chosen 4-byte variant is loaded),   variant =
//       byte 0  qv       the quantized value (q->q[data])
//       byte 1  row of the NEXT column's image for (this cluster, context qv); 0 = no such context
//       byte 2  state | hi<<7      the output symbol (q->output_alphabet->indexes[qv], quantizer choice)
//       byte 3  qratio - 1 of the NEXT column's context qv: the draw comparison of the next symbol is a signed
//               compare of (draw << 24) with the whole variant, no field extraction at all.
//   Address of the next lookup = base + 4*raw data byte + 4A*row + H*hi, H = offset of the hi plane = 255*c:
//       d = previous variant - (draw << 24)      sign bit = hi (draw >= qratio), low 24 bits = the variant's
//       r = PRMT(row word, d)                     bytes (data, row, hi ? 0xFF : 0, 0)      [sign-replicating selector]
//       address = dp4a(r, (4, 4A, c, 0), base)
//   i.e. four integer instructions and the load per symbol.
//   A context the reference would assert on (src/codebook.c:164) leads to row 0, whose entries lead to row 0 of
//   the next column: a line that ever got there ends there, and is reported once, after its last column.
//   The four output bytes / quantized values of a word are gathered with 2 + 2 byte permutes per word; the
//   distortion of -d M / -d A is VABSDIFF4 + IDP.4A per WORD (exact integers), other matrices add their terms
//   per symbol in column order (the reference's sequence of double additions).
// =====================================================================================================
#ifndef QB_THREADS
#define QB_THREADS 512                  // threads per CTA of the walk (512 x 8 lines: 128 registers per thread, no spills in the hot loop;
                                        // measured against 1024 x 4 at 64 registers: walk +4 % on cfg4, +10 % on cfg2 and cfg5)
#endif
#ifndef QB_CTAS
#define QB_CTAS 1                       // resident walk CTAs per SM
#endif
#ifndef QB_LPT
#define QB_LPT 8                        // independent lines per thread
#endif
#define QB_LINES (QB_THREADS * QB_LPT)
static_assert(QVZ_RUN_ALIGN % QB_LINES == 0, "a step (T slots) must be a whole number of walk batches");
#define QB_MAX_BUF 8                    // buffers of the table-image ring
#define QB_POISON_ENTRY 0x007F0000u     // qv 0, next row 0, state byte 0x7F (no quantizer has 127 states), ratio byte 0
#define QB_MAX_ROWS 255u                // the row index travels in one byte
#define QB_MAX_A 62u                    // 4*A must fit the dp4a coefficient byte

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

// largest quantized value any present context can emit for a data value <= smax  -> *vmax
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

// ---- rows of the column images -----------------------------------------------------------------------------
// Forward reachability, one CTA per cluster: reach[col][v] = 1 iff some line of this cluster can arrive at column
// `col` with previous quantized value v.  Column 0: v = 0 (src/qv_compressor.c:89).  Column col+1: every value either
// quantizer of a reachable, present context of column col emits for a data value that occurs there -- `support`
// (bit x of word x>>5 of support[(k*C + col)*3 ..]) when the counting stage recorded it, else every x <= smax.
__global__ void __launch_bounds__(256)
qvz_quantize_reach_kernel(uint32_t C, uint32_t smax, const uint32_t *__restrict__ W, const uint8_t *__restrict__ R,
                          const uint32_t *__restrict__ support, uint8_t *__restrict__ reach)
{
	__shared__ uint32_t cur[72], nxt[72];
	const uint32_t k = blockIdx.x, tid = threadIdx.x;
	if (tid < 72) cur[tid] = tid == 0;
	__syncthreads();
	for (uint32_t col = 0; col < C; ++col) {
		const uint64_t kc = (uint64_t) k * C + col;
		if (tid < 72) {
			reach[kc * 72 + tid] = (uint8_t) cur[tid];
			nxt[tid] = 0;
		}
		__syncthreads();
		if (col + 1 < C) {
			uint32_t sup[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
			if (support) {
				sup[0] = support[kc * 3 + 0];
				sup[1] = support[kc * 3 + 1];
				sup[2] = support[kc * 3 + 2];
			}
			for (uint32_t i = tid; i < 72u * (smax + 1); i += 256) {
				const uint32_t v = i / (smax + 1), x = i - v * (smax + 1);
				const uint32_t ratio = R[kc * 72 + v];
				if (cur[v] && ratio != 0xFF && ((sup[x >> 5] >> (x & 31)) & 1u)) {
					const uint32_t e = W[(kc * 72 + v) * 72 + x];
					if (ratio != 0) nxt[e & 0x7Fu] = 1;            // draw >= 0 always: the lo quantizer is never chosen when qratio == 0
					if (ratio != 128) nxt[(e >> 8) & 0x7Fu] = 1;   // draw <= 127: the hi quantizer is never chosen when qratio == 128
				}                                                  // (benign race: everybody writes 1)
			}
		}
		__syncthreads();
		if (tid < 72) cur[tid] = nxt[tid];
		__syncthreads();
	}
}

// One CTA per column: number the reachable, present (cluster, context) pairs 1, 2, ... -> rowmap[(k*C + col)*72 + v]
// (0 = no row).  Contexts that really mix their two quantizers (0 < qratio < 128) come first: only they need an entry in
// the hi plane -- a context with qratio 128 always takes lo, one with qratio 0 always takes hi (its hi quantizer is stored
// in the lo plane, see the compact kernel) -- so the hi plane is as long as the mixing rows only.
// The largest row count of any column -> *rows_max, the largest mixing-row count -> *mixed_max (atomicMax).
// compact = 0: row = 1 + k*A + v for every present context v < A (no reachability pass, full hi plane).
__global__ void __launch_bounds__(256)
qvz_quantize_rows_kernel(uint32_t K, uint32_t C, uint32_t A, int compact, const uint8_t *__restrict__ R,
                         const uint8_t *__restrict__ reach, uint8_t *__restrict__ rowmap, int *__restrict__ rows_max,
                         int *__restrict__ mixed_max)
{
	__shared__ uint32_t warp_tot[8];
	__shared__ uint32_t carry;
	const uint32_t col = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) carry = 0;
	__syncthreads();
	const uint32_t total = K * 72;
	if (!compact) {
		for (uint32_t i = tid; i < total; i += 256) {
			const uint32_t k = i / 72, v = i - k * 72;
			const uint64_t kc = (uint64_t) k * C + col;
			const bool on = v < A && R[kc * 72 + v] != 0xFF;
			const uint32_t row = on ? 1 + k * A + v : 0;
			rowmap[kc * 72 + v] = (uint8_t) (row > QB_MAX_ROWS ? 0 : row);
			if (on) atomicMax(rows_max, (int) row);
		}
		return;
	}
	for (int pass = 0; pass < 2; ++pass) {               // pass 0: mixing contexts, pass 1: the others
		for (uint32_t base = 0; base < total; base += 256) {
			const uint32_t i = base + tid, k = i / 72, v = i - k * 72;
			const uint64_t kc = (uint64_t) k * C + col;
			bool on = false;
			if (i < total) {
				const uint32_t ratio = R[kc * 72 + v];
				on = v < A && ratio != 0xFF && reach[kc * 72 + v] && ((ratio != 0 && ratio != 128) == (pass == 0));
			}
			const uint32_t m = __ballot_sync(0xFFFFFFFFu, on);
			if (lane == 0) warp_tot[warp] = __popc(m);
			__syncthreads();
			uint32_t before = carry;
			for (uint32_t w = 0; w < warp; ++w) before += warp_tot[w];
			const uint32_t row = 1 + before + __popc(m & ((1u << lane) - 1));
			if (on) rowmap[kc * 72 + v] = (uint8_t) (row > QB_MAX_ROWS ? 0 : row);
			else if (i < total && pass == 0) rowmap[kc * 72 + v] = 0;
			__syncthreads();
			if (tid == 0) {
				uint32_t t = carry;
				for (uint32_t w = 0; w < 8; ++w) t += warp_tot[w];
				carry = t;
			}
			__syncthreads();
		}
		if (pass == 0 && tid == 0) atomicMax(mixed_max, (int) carry);
	}
	if (tid == 0) atomicMax(rows_max, (int) carry);
}

__global__ void __launch_bounds__(256)
qvz_fill_u32_kernel(uint32_t *__restrict__ p, uint64_t n, uint32_t v)
{
	for (uint64_t i = (uint64_t) blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t) gridDim.x * 256) p[i] = v;
}

// full 72x72 tables -> one image per COLUMN (see the format above); G was filled with poison entries before.
// fold: a context with qratio 128 (always lo) or 0 (always hi) has ONE live quantizer; its variants go to the lo plane
// (with the hi bit of the output byte saying which quantizer it was) and everybody who leads there carries the ratio
// byte 127 -- "always lo" -- so the hi plane only holds the rows of contexts that mix (numbered first by the rows kernel).
// start[k] = the variant a line of cluster k "comes from" at column 0: row of (k, context 0), qratio of that context.
__global__ void __launch_bounds__(256)
qvz_quantize_compact_kernel(uint32_t K, uint32_t C, uint32_t A, uint32_t col_words, uint32_t hi_words, int fold,
                            const uint32_t *__restrict__ W, const uint8_t *__restrict__ R, const uint8_t *__restrict__ rowmap,
                            uint32_t *__restrict__ G, uint32_t *__restrict__ start)
{
	const uint64_t idx = (uint64_t) blockIdx.x * 256 + threadIdx.x;
	if (idx < K) {
		const uint64_t kc = idx * C;
		uint32_t r0 = R[kc * 72];
		if (fold && (r0 == 0 || r0 == 128)) r0 = 128;
		start[idx] = r0 == 0xFF ? 0u : ((uint32_t) rowmap[kc * 72] << 8) | (((r0 - 1u) & 0xFFu) << 24);
	}
	if (idx >= (uint64_t) K * C * A * A) return;
	const uint32_t x = idx % A, v = (idx / A) % A;
	const uint64_t kc = idx / ((uint64_t) A * A);
	const uint32_t k = kc / C, col = kc - (uint64_t) k * C;
	const uint32_t row = rowmap[kc * 72 + v];
	if (row == 0) return;                            // no such context, or nothing gets there
	const uint32_t ratio = R[kc * 72 + v];
	const uint32_t e = W[(kc * 72 + v) * 72 + x];
	uint32_t *g = G + (size_t) col * col_words + (size_t) row * A + x;
#pragma unroll
	for (uint32_t hi = 0; hi < 2; ++hi) {
		if (fold && ((ratio == 128 && hi == 1) || (ratio == 0 && hi == 0))) continue;      // never chosen
		const uint32_t qv = (e >> (8 * hi)) & 0xFFu, st = (e >> (16 + 8 * hi)) & 0xFFu;
		uint32_t nrow = 1, nr = 128;                 // last column: any non-zero row = "the line got through"
		if (col + 1 < C) {
			nrow = qv < 72 ? rowmap[(kc + 1) * 72 + qv] : 0;
			nr = nrow ? R[(kc + 1) * 72 + qv] : 128;
			if (fold && nr == 0) nr = 128;           // an always-hi context sits in the lo plane
		}
		nr = (nr - 1u) & 0xFFu;
		const uint32_t var = nrow ? (qv | (nrow << 8) | (st << 16) | (nr << 24)) : QB_POISON_ENTRY;
		const bool to_lo = hi == 0 || (fold && ratio == 0);
		g[to_lo ? 0 : hi_words] = var;
	}
}

// ---- TMA bulk copy + mbarrier (sm_90+/sm_100a): one elected thread moves a whole column-group image
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// PTX prmt with the full 4-bit selector nibbles: bit 3 of a nibble replicates the SIGN of the selected byte
// (__byte_perm keeps only 3 bits per nibble)
__device__ __forceinline__ uint32_t prmt_full(uint32_t a, uint32_t b, uint32_t sel) {
	uint32_t d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));      // sel is a constant after unrolling
	return d;
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

// Geometry of a column image: lo plane (rows*A words), hi plane (hrows*A words) at byte offset 255*c (c % 4 == 0, >= the lo plane's size),
// the whole image padded to a multiple of 16 bytes (TMA bulk copy granularity).
struct qb_geom {
	uint32_t c2;            // dp4a coefficient of the hi byte
	uint32_t hi_off;        // bytes
	uint32_t col_bytes;
};
static __host__ __device__ __forceinline__ qb_geom qb_geometry(uint32_t rows, uint32_t hrows, uint32_t A) {
	qb_geom g;
	const uint32_t plane = rows * A * 4;             // lo plane: every row; hi plane: the first hrows rows (the poison row and the mixing contexts)
	g.c2 = ((plane + 254) / 255 + 3) & ~3u;
	g.hi_off = 255 * g.c2;
	g.col_bytes = (g.hi_off + hrows * A * 4 + 15) & ~15u;
	return g;
}

// DM: 0 arbitrary 72x72 matrix (global loads), 1 doubles indexed by |x-y| (-d L), 2 integers indexed by |x-y|,
//     3 (x-y)^2 (-d M), 4 |x-y| (-d A).      S = columns staged per barrier (4, 2 or 1).
//     DRAWS = false: no reachable context mixes its two quantizers (every qratio is 0 or 128, e.g. what -f 1.0 designs), so no
//     draw can change a symbol: the draw stream is neither generated nor read, and the freed registers carry the row words
//     three word columns ahead (the walk is bound by the latency of its global loads, not by their bandwidth).
template <int DM, bool WANT_QV, int S, bool DRAWS>
__global__ void __launch_bounds__(QB_THREADS, QB_CTAS)
qvz_quantize_batched_kernel(qvz_layout L, const uint32_t *__restrict__ Xw, const uint32_t *__restrict__ Dw,
                            const uint8_t *__restrict__ cl, const uint8_t *__restrict__ G,
                            const uint32_t *__restrict__ start, const double *__restrict__ D, uint32_t rows,
                            uint32_t hrows, uint32_t A, uint32_t *__restrict__ Yw, uint32_t *__restrict__ Qw,
                            double *__restrict__ Ep, int *__restrict__ flags, uint32_t nbuf)
{
	extern __shared__ __align__(16) uint32_t smem[];
	// [dd: 72 doubles (or 72 words)][2 x 8 mbarriers][buffer 0] .. [buffer nbuf-1]; a buffer = S consecutive column images of G,
	// the buffers form a ring: the images of the column groups are requested nbuf-1 groups ahead of their use (the sequence
	// of column groups simply repeats batch after batch, so the ring runs across batch boundaries).
	// full[b]: the image has landed (TMA transaction count); empty[b]: every warp is done reading it (one arrival per warp).
	// There is no CTA-wide barrier in the walk: a warp only waits for the image it needs, and the one thread that issues
	// the copies waits for the buffer it is about to overwrite.
	constexpr uint32_t DD_WORDS = 2 * (QVZ_ALPHABET + 2);
	double *dd = (double *) smem;
	uint32_t *di = smem;
	uint64_t *full = (uint64_t *) (smem + DD_WORDS);
	uint64_t *empty = full + QB_MAX_BUF;
	const uint32_t A4 = A * 4;
	const qb_geom geo = qb_geometry(rows, hrows, A);
	const uint32_t col_bytes = geo.col_bytes;
	const uint32_t buf_bytes = S * col_bytes;
	volatile uint32_t *ps = smem + DD_WORDS + 4 * QB_MAX_BUF;     // producer state (thread 0 only; kept out of the registers of the walk)
	const uint32_t buf0 = smem_u32(smem + DD_WORDS + 4 * QB_MAX_BUF + 4);
	const uint32_t dd_addr = smem_u32(smem);
	const uint32_t tid = threadIdx.x;
	if (DM == 1 && tid < QVZ_ALPHABET) dd[tid] = D[tid];
	if (DM == 2 && tid < QVZ_ALPHABET) di[tid] = (uint32_t) D[tid];
	if (tid == 0) {
		for (uint32_t b = 0; b < nbuf; ++b) {
			mbar_init(&full[b], 1);
			mbar_init(&empty[b], QB_THREADS / 32);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	const uint32_t C = L.C, C4 = L.C4;
	const uint32_t coef = 4u | (A4 << 8) | (geo.c2 << 16);   // dp4a: 4 * data byte + 4A * row byte + c * (hi ? 255 : 0)
	// consumer side (uniform): buffer and phase parity of the column group being walked
	uint32_t cb = 0, cph = 0;
	// producer side (thread 0): ps[0] next buffer to fill, ps[1] how often it has been filled, ps[2] next column group, ps[3] groups left
	auto stage_next = [&]() {                        // thread 0: the next column group of the sequence -> the next buffer of the ring
		const uint32_t pleft = ps[3];
		if (!pleft) return;
		uint32_t pb = ps[0], pf = ps[1];
		const uint32_t pcol = ps[2];
		const uint32_t ncol = (C - pcol < (uint32_t) S) ? C - pcol : (uint32_t) S;
		const uint32_t bytes = ncol * col_bytes;
		if (pf) mbar_wait(&empty[pb], (pf - 1) & 1);                   // every warp has finished the buffer's previous image
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // ... and those generic reads are ordered before the async write
		mbar_expect_tx(&full[pb], bytes);
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		             ::"r"(buf0 + pb * buf_bytes), "l"(G + (uint64_t) pcol * col_bytes), "r"(bytes), "r"(smem_u32(&full[pb])) : "memory");
		ps[2] = pcol + S < C ? pcol + S : 0;
		if (++pb == nbuf) {
			pb = 0;
			pf += 1;
		}
		ps[0] = pb;
		ps[1] = pf;
		ps[3] = pleft - 1;
	};

	bool missing = false;
	// A batch = QB_LINES consecutive slots of ONE step i (slot = i*T + run; T % QB_LINES == 0, abi.cu): all its lines start
	// at the same draw i*C of their runs, so the position in the draw stream is uniform across the CTA.
	const uint32_t bps = L.T / QB_LINES;                 // batches per step
	const uint64_t nbatch = (uint64_t) bps * L.Lr;
	const uint32_t tailmask = (C & 3) ? (0xFFFFFFFFu >> (8 * (4 - (C & 3)))) : 0xFFFFFFFFu;     // live bytes of the last word
	if (tid == 0 && blockIdx.x < nbatch) {
		ps[0] = ps[1] = ps[2] = 0;
		ps[3] = (uint32_t) ((nbatch - blockIdx.x + gridDim.x - 1) / gridDim.x) * ((C + S - 1) / S);
		for (uint32_t b = 0; b + 1 < nbuf; ++b) stage_next();
	}
	for (uint64_t batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
		const uint32_t step = (uint32_t) (batch / bps), boff = (uint32_t) (batch - (uint64_t) step * bps) * QB_LINES;
		const uint64_t pbase = (uint64_t) step * L.T + boff + tid;
		const uint32_t d0 = step * C;                    // first draw of these lines within their runs
		const uint32_t dsh = 8 * (d0 & 3);               // ... = byte d0 & 3 of word d0 >> 2
		uint32_t vprev[QB_LPT], erri[QB_LPT];
		double errd[QB_LPT];
		bool valid[QB_LPT];
#pragma unroll
		for (int j = 0; j < QB_LPT; ++j) {
			const uint32_t kraw = cl[pbase + j * QB_THREADS];
			valid[j] = kraw != QVZ_NO_LINE;
			// column 0: previous value 0 (src/qv_compressor.c:89).  A slot without a line walks cluster 0's tables on
			// symbol 0; wherever that leads (the poison row included) stays inside the image and is not kept.
			vprev[j] = __ldg(start + (valid[j] ? kraw : 0u));
			erri[j] = 0;
			errd[j] = 0.0;                               // 0.0 + d == d exactly: same bits as "error = d" at column 0
		}
		uint32_t xn[QB_LPT], da[QB_LPT], db[QB_LPT];     // draw words w, w+1 of the current step: da is the lower one in even steps, db in odd steps
		                                                 // (!DRAWS: xn, da, db = the row words of word columns w, w+1, w+2 modulo 3)
		const uint32_t *xr = Xw + pbase;                 // running pointer: one word column (P slots) per step
		const uint32_t *drw = Dw + (uint64_t) (d0 >> 2) * L.T + (boff + tid);      // one draw word (T runs) per step
		uint32_t *yr = Yw + pbase, *qr = WANT_QV ? Qw + pbase : nullptr;
#pragma unroll
		for (int j = 0; j < QB_LPT; ++j) {
			xn[j] = ld_stream_u32(xr + j * QB_THREADS);
			if (DRAWS) {
				da[j] = ld_stream_u32(drw + j * QB_THREADS);
				db[j] = ld_stream_u32(drw + L.T + j * QB_THREADS);      // Dw has spare word rows at the end (abi.cu)
			} else {
				da[j] = C4 > 1 ? ld_stream_u32(xr + L.P + j * QB_THREADS) : 0u;
				db[j] = C4 > 2 ? ld_stream_u32(xr + 2 * L.P + j * QB_THREADS) : 0u;
			}
		}
		drw += 2 * (uint64_t) L.T;
		if (!DRAWS) xr += 2 * L.P;                       // xr = the last word column that has been requested
		// one data word (4 columns) of the QB_LPT lines of this thread; TAIL = the last, partial word
		// ODD alternates from word to word: the draw word that was the lower one is dead after the realignment and
		// receives the load for the next step, so no loaded value is ever copied (a copy would wait for the load)
		auto word = [&](uint32_t c4, auto tail_tag, auto odd_tag) {
			constexpr bool TAIL = decltype(tail_tag)::value;
			constexpr int RING = decltype(odd_tag)::value;       // DRAWS: 0 / 1 = even / odd word; !DRAWS: word index modulo 3
			constexpr bool ODD = RING == 1;
			uint32_t (&dlo)[QB_LPT] = ODD ? db : da;
			uint32_t (&dhi)[QB_LPT] = ODD ? da : db;
			uint32_t (&xcur)[QB_LPT] = DRAWS ? xn : (RING == 0 ? xn : RING == 1 ? da : db);
			uint32_t x[QB_LPT], dr[QB_LPT], t01[QB_LPT], t23[QB_LPT];
#pragma unroll
			for (int j = 0; j < QB_LPT; ++j) {
				// raw ASCII bytes index the tables directly: the -33 is folded into the table base below.  A slot without
				// a line (zero words) walks symbol 0 so that every lookup stays inside the image; nothing of it is kept.
				x[j] = valid[j] ? xcur[j] : 0x21212121u;
				if (TAIL) x[j] = (x[j] & tailmask) | (0x21212121u & ~tailmask);       // columns past C count as symbol 0 (never walked)
				// draws d0 + 4*c4 .. +3 of the run.  volatile: stays AHEAD of the loads issued below -- ptxas otherwise hoists those
				// above it, and the shift then waits on a scoreboard shared with loads that have only just been issued
				if (DRAWS) asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(dr[j]) : "r"(dlo[j]), "r"(dhi[j]), "r"(dsh));
				t01[j] = 0;
				t23[j] = 0;
			}
			if (!TAIL && DRAWS) {                        // next word's rows and draws: in flight during this word
				xr += L.P;
#pragma unroll
				for (int j = 0; j < QB_LPT; ++j) {
					xn[j] = ld_stream_u32(xr + j * QB_THREADS);
					dlo[j] = ld_stream_u32(drw + j * QB_THREADS);
				}
				drw += L.T;
			}
			if (!TAIL && !DRAWS && c4 + 3 < C4) {        // the rows of word column c4 + 3 take the place of this word's
				xr += L.P;
#pragma unroll
				for (int j = 0; j < QB_LPT; ++j) xcur[j] = ld_stream_u32(xr + j * QB_THREADS);
			}
#pragma unroll
			for (int g = 0; g < 4 / S; ++g) {            // the column groups (= staged images) inside this word
				const uint32_t col0 = 4 * c4 + g * S;
				if (!TAIL || col0 < C) {
					if (tid == 0) stage_next();          // the group nbuf-1 ahead of this one
					mbar_wait(&full[cb], cph);           // this group's image has landed
					const uint32_t tabg = buf0 + cb * buf_bytes;
#pragma unroll
					for (int sidx = 0; sidx < S; ++sidx) {
						const int b = g * S + sidx;
						if (!TAIL || col0 + sidx < C) {
							const uint32_t tab = tabg + sidx * col_bytes - 33u * 4u;       // indexed by the raw byte ('!' + value)
#pragma unroll
							for (int j = 0; j < QB_LPT; ++j) {
								// draw >= qratio  <=>  (int)((qratio-1) << 24 | low bits of the previous variant) - (int)(draw << 24) < 0
								// (exact in 32 bits: the variant is >= -2^24, the draw term <= 127 << 24), and the low 24 bits of the
								// difference are still the variant's
								// (!DRAWS: every ratio byte is 127 = "always lo", the sign of the variant itself is the answer)
								const uint32_t d = DRAWS ? vprev[j] - __byte_perm(dr[j], 0, 0x0444 + (b << 12)) : vprev[j];
								// byte 0 = raw data byte b of the row word, byte 1 = row byte of the previous variant,
								// byte 2 = 0xFF if hi (sign of d, replicated), byte 3 = 0 (sign of an ASCII byte)
								const uint32_t r = prmt_full(x[j], d, (uint32_t) (b | (5 << 4) | (0xF << 8) | ((8 | b) << 12)));
								const uint32_t v = lds_u32(__dp4a(r, coef, tab));
								// gather (qv, state) of the symbol pairs (0,1) and (2,3): bytes [qv_even, qv_odd, st_even, st_odd]
								// (the even symbol's variant is still in vprev when the odd one arrives)
								if (b == 1) t01[j] = __byte_perm(vprev[j], v, 0x6240);
								else if (b == 3) t23[j] = __byte_perm(vprev[j], v, 0x6240);
								vprev[j] = v;
							}
						}
					}
					__syncwarp();                        // this warp is done with the group's image
					if ((tid & 31) == 0) mbar_arrive(&empty[cb]);
					if (++cb == nbuf) {
						cb = 0;
						cph ^= 1;
					}
				}
			}
			uint32_t qvw[QB_LPT];
#pragma unroll
			for (int j = 0; j < QB_LPT; ++j) {
				if (TAIL) {                              // an unpaired last symbol (still in vprev): its partner is "nothing"
					if ((C & 3) == 1) t01[j] = __byte_perm(vprev[j], 0, 0x6240);
					if ((C & 3) == 3) t23[j] = __byte_perm(vprev[j], 0, 0x6240);
				}
				qvw[j] = __byte_perm(t01[j], t23[j], 0x5410);
				st_stream_u32(yr + j * QB_THREADS, __byte_perm(t01[j], t23[j], 0x7632));
			}
			yr += L.P;
			if (WANT_QV) {
#pragma unroll
				for (int j = 0; j < QB_LPT; ++j) st_stream_u32(qr + j * QB_THREADS, qvw[j] + 0x21212121u);
				qr += L.P;
			}
			// distortion of this word's symbols (src/qv_compressor.c:97,118): data and quantized values side by side
#pragma unroll
			for (int j = 0; j < QB_LPT; ++j) {
				const uint32_t xq = x[j] - 0x21212121u;      // every byte >= 33: no borrow
				if (DM == 3 || DM == 4) {
					const uint32_t d4 = __vabsdiffu4(xq, qvw[j]);
					erri[j] = __dp4a(d4, DM == 3 ? d4 : 0x01010101u, erri[j]);
				} else {
					const uint32_t d4 = DM == 0 ? 0u : __vabsdiffu4(xq, qvw[j]);
#pragma unroll
					for (int b = 0; b < 4; ++b) {
						if (!TAIL || 4 * c4 + b < C) {
							if (DM == 2) erri[j] += lds_u32(dd_addr + __byte_perm(d4, 0, 0x4440 + b) * 4);
							else if (DM == 1) errd[j] += lds_f64(dd_addr + __byte_perm(d4, 0, 0x4440 + b) * 8);
							else if (valid[j]) errd[j] += __ldg(&D[__byte_perm(xq, 0, 0x4440 + b) + 72u * __byte_perm(qvw[j], 0, 0x4440 + b)]);
						}
					}
				}
			}
		};
		uint32_t c4 = 0;
		using R0 = std::integral_constant<int, 0>;
		using R1 = std::integral_constant<int, 1>;
		using R2 = std::integral_constant<int, 2>;
		if (DRAWS) {
			for (; c4 + 2 < C4; c4 += 2) {
				word(c4, std::false_type{}, R0{});
				word(c4 + 1, std::false_type{}, R1{});
			}
			if (c4 + 1 < C4) {
				word(c4, std::false_type{}, R0{});
				word(c4 + 1, std::true_type{}, R1{});
			} else word(c4, std::true_type{}, R0{});
		} else {
			for (; c4 + 3 < C4; c4 += 3) {
				word(c4, std::false_type{}, R0{});
				word(c4 + 1, std::false_type{}, R1{});
				word(c4 + 2, std::false_type{}, R2{});
			}
			if (c4 + 2 < C4) {
				word(c4, std::false_type{}, R0{});
				word(c4 + 1, std::false_type{}, R1{});
				word(c4 + 2, std::true_type{}, R2{});
			} else if (c4 + 1 < C4) {
				word(c4, std::false_type{}, R0{});
				word(c4 + 1, std::true_type{}, R1{});
			} else word(c4, std::true_type{}, R0{});
		}
#pragma unroll
		for (int j = 0; j < QB_LPT; ++j) {
			// a line that met a context without a quantizer sits in the poison row (row byte 0) after its last column
			missing |= valid[j] && (vprev[j] & 0x0000FF00u) == 0u;
			// (an FMA-based correctly rounded quotient -- Markstein's step, RN(q + (a - q*C) * RN(1/C)) -- was checked exact on
			// 5*10^8 values and measured 7 % SLOWER for the whole walk: register allocation of the hot loop, not the division, decides)
			Ep[pbase + j * QB_THREADS] = ((DM >= 2) ? (double) erri[j] : errd[j]) / (double) C;
		}
	}
	if (missing) atomicOr(&flags[2], 1);
}

static size_t batched_smem(uint32_t rows, uint32_t hrows, uint32_t A, uint32_t S, uint32_t nbuf = 2) {
	return 2 * (QVZ_ALPHABET + 2) * sizeof(uint32_t) + 16 * QB_MAX_BUF + 16 + (size_t) nbuf * S * qb_geometry(rows, hrows, A).col_bytes;
}

// buffers of the ring: 2 (measured on cfg2, where more fit: 3 buffers 2.38 TB/s, 2 buffers 2.42; on cfg4 four buffers of 2
// columns 2.98 TB/s against 3.12 for two of 4 columns: the size of a group matters, the depth of the ring does not);
// QVZ_WALK_NBUF = up to as many as fit
static uint32_t batched_nbuf(uint32_t rows, uint32_t hrows, uint32_t A, uint32_t S) {
	uint32_t fit = 2, n = 2;
	while (fit < QB_MAX_BUF && batched_smem(rows, hrows, A, S, fit + 1) <= (QB_CTAS == 1 ? 226 : 112) * 1024) ++fit;
	if (const char *e = getenv("QVZ_WALK_NBUF")) n = (uint32_t) atoi(e) >= 2 && (uint32_t) atoi(e) <= fit ? (uint32_t) atoi(e) : n;
	return n;
}

size_t qvz_quantize_image_bytes(uint32_t C, uint32_t rows, uint32_t hrows, uint32_t A) { return (size_t) C * qb_geometry(rows, hrows, A).col_bytes; }

// columns staged per barrier: the largest of 4, 2, 1 whose double buffer fits; 0 = batched path unusable
uint32_t qvz_quantize_batched_group(uint32_t rows, uint32_t hrows, uint32_t A) {
	if (A > QB_MAX_A || rows > QB_MAX_ROWS + 1) return 0;
	uint32_t smax = 4;
	if (const char *e = getenv("QVZ_WALK_S")) smax = atoi(e) == 1 ? 1 : atoi(e) == 2 ? 2 : 4;      // tuning knob
	for (uint32_t S = smax; S >= 1; S >>= 1)
		if (batched_smem(rows, hrows, A, S) <= (QB_CTAS == 1 ? 226 : 112) * 1024) return S;       // 227 KB per CTA on sm_100
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

// rowmap (h->rowmap, K*C*72 bytes), the row count of the fullest column (-> flags[7]) and its mixing-row count (-> flags[4]); support may be nullptr
int qvz_quantize_rows(qvz_gpu *h, uint32_t K, uint32_t C, uint32_t A, int compact, const uint32_t *support) {
	if (compact) {
		qvz_quantize_reach_kernel<<<K, 256, 0, h->stream>>>(C, h->smax > 71 ? 71 : h->smax, h->W, h->R, support, h->reach);
		QVZ_LAUNCHED(h);
		QVZ_CUDA(h, cudaGetLastError());
	}
	qvz_quantize_rows_kernel<<<C, 256, 0, h->stream>>>(K, C, A, compact, h->R, h->reach, h->rowmap, h->flags + 7, h->flags + 4);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}

// rows = rows per plane of every column image (poison row included)
int qvz_quantize_compact(qvz_gpu *h, uint32_t K, uint32_t C, uint32_t A, uint32_t rows, uint32_t hrows, int fold) {
	const qb_geom geo = qb_geometry(rows, hrows, A);
	const uint64_t words = (uint64_t) C * (geo.col_bytes / 4);
	qvz_fill_u32_kernel<<<h->sm_count * 4, 256, 0, h->stream>>>((uint32_t *) h->G, words, QB_POISON_ENTRY);
	QVZ_LAUNCHED(h);
	const uint64_t total = (uint64_t) K * C * A * A;
	qvz_quantize_compact_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, h->stream>>>(K, C, A, geo.col_bytes / 4, geo.hi_off / 4, fold, h->W, h->R, h->rowmap, (uint32_t *) h->G, h->start);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}

template <int DM, bool WANT_QV, int S, bool DRAWS>
static void launch_batched_d(qvz_gpu *h, uint32_t rows, uint32_t hrows, uint32_t A) {
	auto kern = qvz_quantize_batched_kernel<DM, WANT_QV, S, DRAWS>;
	const uint32_t nbuf = batched_nbuf(rows, hrows, A, S);
	const size_t smem = batched_smem(rows, hrows, A, S, nbuf);
	cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	const uint64_t nbatch = (uint64_t) (h->L.T / QB_LINES) * h->L.Lr;     // T % QB_LINES == 0 (QVZ_RUN_ALIGN)
	const uint64_t resident = (uint64_t) h->sm_count * QB_CTAS;
	const unsigned grid = (unsigned) (nbatch < resident ? nbatch : resident);
	kern<<<grid, QB_THREADS, smem, h->stream>>>(h->L, h->Xw, h->Dw, h->cl, h->G, h->start, h->D, rows, hrows, A, h->Yw,
	                                            WANT_QV ? h->Qw : nullptr, h->Ep, h->flags, nbuf);
}

template <int DM, bool WANT_QV, int S>
static void launch_batched(qvz_gpu *h, uint32_t rows, uint32_t hrows, uint32_t A) {
	if (h->tab_nodraw) launch_batched_d<DM, WANT_QV, S, false>(h, rows, hrows, A);
	else launch_batched_d<DM, WANT_QV, S, true>(h, rows, hrows, A);
}

template <int DM, bool WANT_QV>
static void launch_batched_s(qvz_gpu *h, uint32_t rows, uint32_t hrows, uint32_t A, uint32_t S) {
	if (S == 4) launch_batched<DM, WANT_QV, 4>(h, rows, hrows, A);
	else if (S == 2) launch_batched<DM, WANT_QV, 2>(h, rows, hrows, A);
	else launch_batched<DM, WANT_QV, 1>(h, rows, hrows, A);
}

template <bool WANT_QV>
static void launch_batched_dm(qvz_gpu *h, uint32_t rows, uint32_t hrows, uint32_t A, uint32_t S, int dm) {
	switch (dm) {
	case 4: launch_batched_s<4, WANT_QV>(h, rows, hrows, A, S); break;
	case 3: launch_batched_s<3, WANT_QV>(h, rows, hrows, A, S); break;
	case 2: launch_batched_s<2, WANT_QV>(h, rows, hrows, A, S); break;
	case 1: launch_batched_s<1, WANT_QV>(h, rows, hrows, A, S); break;
	default: launch_batched_s<0, WANT_QV>(h, rows, hrows, A, S); break;
	}
}

// dm: see the kernel
int qvz_quantize_launch_batched(qvz_gpu *h, uint32_t rows, uint32_t hrows, uint32_t A, int want_qv, int dm) {
	const uint32_t S = qvz_quantize_batched_group(rows, hrows, A);
	if (getenv("QVZ_DEBUG_WALK")) fprintf(stderr, "[walk] rows %u hi rows %u A %u S %u ring %u smem %zu dm %d draws %d\n", rows, hrows, A, S, batched_nbuf(rows, hrows, A, S), batched_smem(rows, hrows, A, S, batched_nbuf(rows, hrows, A, S)), dm, !h->tab_nodraw);
	if (want_qv) launch_batched_dm<true>(h, rows, hrows, A, S, dm);
	else launch_batched_dm<false>(h, rows, hrows, A, S, dm);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}
