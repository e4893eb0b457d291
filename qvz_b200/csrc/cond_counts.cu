// cond_counts.cu -- first-order conditional counts per (cluster, column, previous RAW value, value).
//
// Reference: the counting loop of calculate_statistics (src/codebook.c:193-205):
//     pmf_increment(get_cond_pmf(list, 0, 0), x[0]-33)
//     pmf_increment(get_cond_pmf(list, c, x[c-1]-33), x[c]-33)          for c >= 1
// with get_cond_pmf -> pmfs[1 + (c-1)*72 + prev] (src/codebook.c:116-120) and pmf_increment ->
// counts[idx]++, total++ (src/pmf.c:211-214).  Output layout = that pmfs[] order, 72 counters per row.
//
// Mapping: one CTA owns one packed word column c4 (= 4 table columns) for a chunk of slots and a group of
// clusters; its 72x72 slices live in shared memory; lanes <-> slots so global reads are coalesced; one
// shared-memory atomic per symbol; non-zero counters are flushed with one global atomic each.
//
// Shared atomics: `red.shared.add 1` compiles to ATOMS.POPC.INC, which merges the lanes of a warp that hit the
// same address (measured on B200: 1 cycle per warp instruction when the distinct addresses fall in distinct
// banks, ~3.6 for random addresses; tools/ubench/atoms.cu).  Neighbouring reads agree on (prev, cur) very
// often, so all lanes count the same column at the same step (few distinct addresses per instruction).
// The slices are A x A 32-bit counters, A-1 = the largest symbol present in the rows (found at ingest): for
// Phred+33 data with Q <= 41 a slice is 7 KB instead of 20 KB, so up to 7 clusters are counted in one pass.
#include "qvz_internal.cuh"

#define CC_THREADS 1024
#define CC_PAD 8u                               // words between slices

__device__ __forceinline__ uint4 cc_ldg128(const uint32_t *p) {
	uint4 v;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}
__device__ __forceinline__ void cc_red_shared(uint32_t addr, uint32_t v) {
	asm volatile("red.shared::cta.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");     // ::cta: no cluster-window address arithmetic
}

// PTX prmt with the full 4-bit selector nibbles (bit 3 = replicate the sign of the selected byte; __byte_perm keeps 3 bits)
__device__ __forceinline__ uint32_t cc_prmt(uint32_t a, uint32_t b, uint32_t sel) {
	uint32_t d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));      // sel is a constant after unrolling
	return d;
}

// the 4 slots of one 16-byte load: w4 = this word column, q4 = the previous one, k4 = the 4 cluster ids
template <bool TAIL>
__device__ __forceinline__ void cc_count4(const uint4 &w4, const uint4 &q4, uint32_t k4, uint32_t kbase, uint32_t G,
                                          uint32_t tab_biased, uint32_t live, uint32_t coef, uint32_t slice_bytes, bool single)
{
	const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w}, qs[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
	for (int u = 0; u < 4; ++u) {
		const uint32_t g = ((k4 >> (8 * u)) & 0xFFu) - kbase;          // wraps to a huge value for k < kbase and for 0xFF
		if (ws[u] != 0u && g < G) {                      // a slot without a line holds zero words (real bytes are >= 33)
			const uint32_t pv = __funnelshift_r(qs[u], ws[u], 24);         // raw bytes: the previous column's value of bytes 0, 1, 2, 3
			const uint32_t b0 = single ? tab_biased : tab_biased + g * (4 * slice_bytes);
#pragma unroll
			for (uint32_t j = 0; j < 4; ++j) {
				if (!TAIL || j < live) {
					// (value, previous, previous, 0) . (4, c1, c2, 0) = 4*value + 4A*previous: the counter's byte offset in its slice
					// (the '!' offsets of the raw bytes are folded into tab_biased)
					const uint32_t r = cc_prmt(ws[u], pv, j | ((4 + j) << 4) | ((4 + j) << 8) | ((8 | j) << 12));
					cc_red_shared(__dp4a(r, coef, b0 + j * slice_bytes), 1u);
				}
			}
		}
	}
}

// A thread takes 4 consecutive slots per iteration (16-byte loads of the word column, of the previous word
// column and one 4-byte load of the cluster ids): 16 symbols per 3 loads; the loads of the NEXT iteration are issued
// before the 16 atomics of the current one.  TAIL = this word column holds the last, partial word of the lines
// (columns past C must not be counted).
template <bool TAIL>
__device__ __forceinline__ void cc_count_chunk(const uint32_t *__restrict__ xc, const uint32_t *__restrict__ xp,
                                               const uint8_t *__restrict__ clp, uint32_t n, bool first, bool single,
                                               uint32_t kbase, uint32_t G, uint32_t tab_addr, uint32_t live,
                                               uint32_t A, uint32_t slice_bytes)
{
	const uint32_t A4 = A * 4;                           // <= 288: split over two coefficient bytes
	const uint32_t c1 = A4 > 252 ? 252 : A4, c2 = A4 - c1;
	const uint32_t coef = 4u | (c1 << 8) | (c2 << 16);
	const uint32_t tab_biased = tab_addr - 33u * (4u + A4);
	const uint4 bang = make_uint4(0x21212121u, 0x21212121u, 0x21212121u, 0x21212121u);
	uint32_t i = 4 * threadIdx.x;
	if (i >= n) return;
	uint4 w4 = cc_ldg128(xc + i), q4 = first ? bang : cc_ldg128(xp + i);
	uint32_t k4 = single ? 0u : __ldg((const uint32_t *) (clp + i));
	for (;;) {
		const uint32_t inext = i + 4 * CC_THREADS;
		const bool more = inext < n;
		uint4 w4n = bang, q4n = bang;
		uint32_t k4n = 0;
		if (more) {
			w4n = cc_ldg128(xc + inext);
			if (!first) q4n = cc_ldg128(xp + inext);
			if (!single) k4n = __ldg((const uint32_t *) (clp + inext));
		}
		cc_count4<TAIL>(w4, q4, k4, kbase, G, tab_biased, live, coef, slice_bytes, single);
		if (!more) break;
		w4 = w4n;
		q4 = q4n;
		k4 = k4n;
		i = inext;
	}
}

__global__ void __launch_bounds__(CC_THREADS, 1)
qvz_cond_counts_kernel(qvz_layout L, const uint32_t *__restrict__ Xw, const uint8_t *__restrict__ cl,
                       uint32_t K, uint32_t G, uint32_t A, uint32_t chunk, uint32_t *__restrict__ counts)
{
	extern __shared__ uint32_t tab[];            // [G][4][slice]
	const uint32_t slice = A * A + CC_PAD;
	const uint32_t c4 = blockIdx.x;
	const uint32_t kbase = blockIdx.z * G;
	const uint64_t p0 = (uint64_t) blockIdx.y * chunk;
	const uint64_t p1 = (p0 + chunk < L.P) ? p0 + chunk : L.P;      // chunk % 4096 == 0 and P % 4096 == 0
	const uint32_t words = G * 4 * slice;

	for (uint32_t i = threadIdx.x; i < words; i += CC_THREADS) tab[i] = 0;
	__syncthreads();

	const uint32_t tab_addr = (uint32_t) __cvta_generic_to_shared(tab);
	const uint32_t live = L.C - 4 * c4;          // columns of this word that exist (>= 4 except in the last word)
	const uint32_t *xc = Xw + (uint64_t) c4 * L.P + p0;
	const uint32_t *xp = xc - L.P;               // only dereferenced for c4 > 0
	const uint32_t n = (uint32_t) (p1 - p0);
	if (4 * c4 + 3 < L.C) cc_count_chunk<false>(xc, xp, cl + p0, n, c4 == 0, K == 1, kbase, G, tab_addr, live, A, slice * 4);
	else cc_count_chunk<true>(xc, xp, cl + p0, n, c4 == 0, K == 1, kbase, G, tab_addr, live, A, slice * 4);
	__syncthreads();

	const uint64_t per_cluster = (uint64_t) (1 + 72 * (L.C - 1)) * 72;
	for (uint32_t i = threadIdx.x; i < words; i += CC_THREADS) {
		const uint32_t v = tab[i];
		if (!v) continue;
		const uint32_t s = i / slice, cell = i - s * slice;
		const uint32_t g = s >> 2, j = s & 3, col = 4 * c4 + j;
		if (kbase + g >= K || cell >= A * A) continue;
		const uint32_t prev = cell / A, cur = cell - prev * A;
		// column 0 only ever sees prev == 0 => pmfs[0]; column c >= 1, previous value prev => pmfs[1 + (c-1)*72 + prev]
		uint32_t *dst = counts + (uint64_t) (kbase + g) * per_cluster + (col ? (uint64_t) (1 + (col - 1) * 72 + prev) * 72 : 0);
		atomicAdd(dst + cur, v);
	}
}

// ---- one cluster (the reference's default, -c 1): lane-private counters over byte planes
//
// With K == 1 a slice is A x A counters, and for Phred+33 data with Q <= 41 (A = 42) thirty-two copies of it fit the
// 227 KB of shared memory: copy l lives entirely in bank l and is only ever touched by lane l, so every warp-wide
// atomic is conflict-free (one wavefront instead of ~3.3 for 32 random addresses; ATOMS across warps stay atomic).
// That needs one column per CTA at a time, so this kernel reads the rows as byte planes Xb[c][p] (kept next to Xw at
// ingest): 16 slots of column c and of column c-1 per thread and iteration (two 16-byte loads for 16 symbols; the
// plane of column c is read again as `prev` by the CTA working one column ahead at the same slots, an L2 hit).
// Work is cut into equal contiguous spans of the (column, slot) sequence, one per SM: a CTA crosses at most two
// column boundaries, where it reduces its 32 copies (one global atomic per non-zero counter) and clears them.
#define CP_THREADS 1024
#define CP_SLOTS 16u                            // slots per thread and iteration
#define CP_BLOCK (CP_THREADS * CP_SLOTS)        // slots per CTA iteration

__device__ __forceinline__ uint4 cp_ldg128(const uint8_t *p) {
	uint4 v;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}

// four slots of one column: byte b of cw = value, byte b of pw = previous column's value (raw ASCII).
// CHECK = some of the 16 slots of this thread hold no line (zero bytes): test every byte.
template <bool CHECK>
__device__ __forceinline__ void cp_count_word(uint32_t cw, uint32_t pw, uint32_t coef, int bias, uint32_t lane_addr)
{
#pragma unroll
	for (uint32_t b = 0; b < 4; ++b) {
		if (!CHECK || ((cw >> (8 * b)) & 0xFFu)) {
			const uint32_t r = __byte_perm(cw, pw, b | ((4 + b) << 4) | 0x8800u);       // value | prev << 8 (bytes < 128: upper bytes 0)
			const uint32_t idx = (uint32_t) __dp4a((int) r, (int) coef, bias);          // (prev-33)*A + (value-33)
			cc_red_shared(lane_addr + idx * 128u, 1u);
		}
	}
}

__device__ __forceinline__ uint32_t cp_zero_bytes(uint32_t w) { return (w - 0x01010101u) & ~w & 0x80808080u; }

// 16 slots of one column (cw) and of the column before it (pw)
__device__ __forceinline__ void cp_count16(const uint4 &cw, const uint4 &pw, uint32_t coef, int bias, uint32_t lane_addr)
{
	if ((cp_zero_bytes(cw.x) | cp_zero_bytes(cw.y) | cp_zero_bytes(cw.z) | cp_zero_bytes(cw.w)) == 0u) {   // 16 real lines
		cp_count_word<false>(cw.x, pw.x, coef, bias, lane_addr);
		cp_count_word<false>(cw.y, pw.y, coef, bias, lane_addr);
		cp_count_word<false>(cw.z, pw.z, coef, bias, lane_addr);
		cp_count_word<false>(cw.w, pw.w, coef, bias, lane_addr);
	} else {
		cp_count_word<true>(cw.x, pw.x, coef, bias, lane_addr);
		cp_count_word<true>(cw.y, pw.y, coef, bias, lane_addr);
		cp_count_word<true>(cw.z, pw.z, coef, bias, lane_addr);
		cp_count_word<true>(cw.w, pw.w, coef, bias, lane_addr);
	}
}

// n blocks of one column for this thread, two in flight (the loads of the next block are issued before the
// atomics of the current one).  FIRST = column 0: the previous value is 0 for every line (-> pmfs[0]).
template <bool FIRST>
__device__ __forceinline__ void cp_segment(const uint8_t *pc, uint64_t P, uint32_t n, uint32_t coef, int bias, uint32_t lane_addr)
{
	const uint4 bang = make_uint4(0x21212121u, 0x21212121u, 0x21212121u, 0x21212121u);
	if (n == 0) return;
	const uint8_t *pp = pc - P;
	uint4 a = cp_ldg128(pc), ap = bang, b, bp = bang;
	if (!FIRST) ap = cp_ldg128(pp);
	uint32_t k = 0;
	for (; k + 2 <= n; k += 2) {
		b = cp_ldg128(pc + CP_BLOCK);
		if (!FIRST) bp = cp_ldg128(pp + CP_BLOCK);
		pc += 2 * CP_BLOCK;
		pp += 2 * CP_BLOCK;
		cp_count16(a, ap, coef, bias, lane_addr);
		if (k + 2 < n) {
			a = cp_ldg128(pc);
			if (!FIRST) ap = cp_ldg128(pp);
		}
		cp_count16(b, bp, coef, bias, lane_addr);
	}
	if (k < n) cp_count16(a, ap, coef, bias, lane_addr);
}

__global__ void __launch_bounds__(CP_THREADS, 1)
qvz_cond_counts_planes_kernel(qvz_layout L, const uint8_t *__restrict__ Xb, uint32_t A, uint32_t S, uint32_t *__restrict__ counts)
{
	extern __shared__ uint32_t hist[];           // [A*A][32]: counter (prev, value) of lane l at word (prev*A + value)*32 + l
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t cells = A * A;
	for (uint32_t i = tid; i < cells * 32; i += CP_THREADS) hist[i] = 0;
	__syncthreads();

	const uint64_t nblk = (L.P + CP_BLOCK - 1) / CP_BLOCK;              // CTA iterations per column
	const uint32_t lane_addr = (uint32_t) __cvta_generic_to_shared(hist) + lane * 4u;
	const uint32_t coef = 1u | (A << 8);
	const int bias = -33 * (int) (A + 1);

	auto flush = [&](uint32_t col) {             // 32 copies -> pmfs[] rows of this column, then clear
		__syncthreads();
		for (uint32_t cell = warp; cell < cells; cell += CP_THREADS / 32) {
			const uint32_t v = hist[cell * 32 + lane];
			if (__any_sync(0xFFFFFFFFu, v != 0u)) {
				hist[cell * 32 + lane] = 0;
				const uint32_t s = __reduce_add_sync(0xFFFFFFFFu, v);
				if (lane == 0) {
					const uint32_t prev = cell / A, cur = cell - prev * A;
					atomicAdd(counts + (col ? (uint64_t) (1 + (col - 1) * 72 + prev) * 72 : 0) + cur, s);
				}
			}
		}
		__syncthreads();
	};

	const uint64_t toff = (uint64_t) tid * CP_SLOTS;
	auto job = [&](uint32_t col, uint64_t blk, uint64_t bend) {          // blocks [blk, bend) of one column, then flush
		const uint8_t *pc = Xb + (uint64_t) col * L.P + blk * CP_BLOCK + toff;
		// P % 4096 == 0: a thread's 16 slots are all inside the plane or all outside (last block of a column only)
		uint32_t n = (uint32_t) (bend - blk);
		if (n && bend == nblk && (nblk - 1) * CP_BLOCK + toff >= L.P) n -= 1;
		if (col) cp_segment<false>(pc, L.P, n, coef, bias, lane_addr);
		else cp_segment<true>(pc, L.P, n, coef, bias, lane_addr);
		flush(col);
	};
	// Neighbouring CTAs work on neighbouring columns at the same slots at the same time, so the plane a CTA reads as
	// `prev` is the one its neighbour is reading as `value` (one HBM read, one L2 hit).  First whole columns, one per
	// CTA and round; then the columns that are left, cut into S pieces each and dealt round-robin, column fastest.
	const uint32_t G = gridDim.x, rounds = L.C / G, rem = L.C - rounds * G;
	for (uint32_t r = 0; r < rounds; ++r) job(r * G + blockIdx.x, 0, nblk);
	for (uint64_t id = blockIdx.x; id < (uint64_t) S * rem; id += G) {
		const uint64_t sp = id / rem;
		const uint32_t j = (uint32_t) (id - sp * rem);
		job(rounds * G + j, sp * nblk / S, (sp + 1) * nblk / S);
	}
}

// Which data values occur per (cluster, column): bit x of support[(k*C + col)*3 + (x >> 5)].  Read off the count table
// (column marginals over the previous value); the quantize stage uses it to stage only the table rows that the
// resident rows can reach (quantize.cu).  grid (C, K), 96 threads: thread x, 72 coalesced row reads.
__global__ void __launch_bounds__(96)
qvz_cond_counts_support_kernel(uint32_t C, const uint32_t *__restrict__ counts, uint32_t *__restrict__ support)
{
	const uint32_t col = blockIdx.x, k = blockIdx.y, x = threadIdx.x;
	const uint64_t per_cluster = (uint64_t) (1 + 72 * (C - 1)) * 72;
	const uint32_t *base = counts + k * per_cluster;
	uint32_t any = 0;
	if (x < 72) {
		if (col == 0) any = base[x];
		else
			for (uint32_t prev = 0; prev < 72; ++prev) any |= base[(uint64_t) (1 + (col - 1) * 72 + prev) * 72 + x];
	}
	const uint32_t m = __ballot_sync(0xFFFFFFFFu, any != 0u);
	if ((x & 31) == 0) support[((uint64_t) k * C + col) * 3 + (x >> 5)] = m;
}

int qvz_cond_counts_support(qvz_gpu *h, const uint32_t *counts_dev) {
	const uint32_t K = h->K, C = h->L.C;
	const size_t bytes = (size_t) K * C * 3 * sizeof(uint32_t);
	if (h->support_cap < bytes) {
		if (h->support) cudaFree(h->support);
		h->support = nullptr;
		h->support_cap = 0;
		QVZ_CUDA(h, cudaMalloc(&h->support, bytes));
		h->support_cap = bytes;
	}
	qvz_cond_counts_support_kernel<<<dim3(C, K), 96, 0, h->stream>>>(C, counts_dev, h->support);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	h->support_valid = 1;
	h->support_K = K;
	return QVZ_OK;
}

// Xw[c4][p] -> Xb[c][p]: a thread turns the words of 4 consecutive slots into one word per byte plane (4 x 4 byte
// transpose with byte permutes), so both sides are coalesced.
__global__ void __launch_bounds__(256)
qvz_planes_build_kernel(qvz_layout L, const uint32_t *__restrict__ Xw, uint8_t *__restrict__ Xb)
{
	const uint64_t quads = L.P / 4;                      // P % 4096 == 0
	const uint32_t c4 = blockIdx.y;
	for (uint64_t q = (uint64_t) blockIdx.x * 256 + threadIdx.x; q < quads; q += (uint64_t) gridDim.x * 256) {
		const uint4 w = cc_ldg128(Xw + (uint64_t) c4 * L.P + 4 * q);
		const uint32_t t0 = __byte_perm(w.x, w.y, 0x5140), t1 = __byte_perm(w.z, w.w, 0x5140);
		const uint32_t t2 = __byte_perm(w.x, w.y, 0x7362), t3 = __byte_perm(w.z, w.w, 0x7362);
		const uint32_t o[4] = {__byte_perm(t0, t1, 0x5410), __byte_perm(t0, t1, 0x7632), __byte_perm(t2, t3, 0x5410), __byte_perm(t2, t3, 0x7632)};
#pragma unroll
		for (uint32_t j = 0; j < 4; ++j)
			if (4 * c4 + j < L.C) *(uint32_t *) (Xb + (uint64_t) (4 * c4 + j) * L.P + 4 * q) = o[j];
	}
}

// The lane-private kernel wants the rows as byte planes: a second copy of the rows, made here the first time a
// one-cluster count of the resident rows is asked for (and dropped again if the memory is needed: abi.cu).
// 42 = largest alphabet box whose 32 lane-private copies fit shared memory (Phred+33 with Q <= 41).
static bool planes_ready(qvz_gpu *h, uint32_t A) {
	if (h->K != 1 || (size_t) A * A * 128 > 226 * 1024 || getenv("QVZ_COUNTS_WORDS") || getenv("QVZ_NO_PLANES")) return false;
	const size_t bytes = (size_t) h->L.C * h->L.P;
	if (!h->Xb || h->Xb_cap < bytes) {
		if (h->Xb) cudaFree(h->Xb);
		h->Xb = nullptr;
		h->Xb_cap = 0;
		h->Xb_valid = 0;
		if (cudaMalloc(&h->Xb, bytes) != cudaSuccess) {
			cudaGetLastError();                          // no room for a second copy: the word-column kernel counts
			h->Xb = nullptr;
			return false;
		}
		h->Xb_cap = bytes;
	}
	if (!h->Xb_valid) {
		qvz_planes_build_kernel<<<dim3(h->sm_count * 2, h->L.C4), 256, 0, h->stream>>>(h->L, h->Xw, h->Xb);
		QVZ_LAUNCHED(h);
		if (cudaGetLastError() != cudaSuccess) return false;
		h->Xb_valid = 1;
	}
	return true;
}

int qvz_cond_counts_launch(qvz_gpu *h, uint32_t *counts_dev) {
	const uint32_t K = h->K;
	const uint32_t A = h->smax + 1 > 72 ? 72 : h->smax + 1;
	if (planes_ready(h, A)) {
		const size_t smem = (size_t) A * A * 128;
		QVZ_CUDA(h, cudaMemsetAsync(counts_dev, 0, qvz_gpu_cond_counts_len(1, h->L.C) * sizeof(uint32_t), h->stream));
		QVZ_CUDA(h, cudaFuncSetAttribute(qvz_cond_counts_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		// pieces per left-over column: fewest waves of (piece + flush), the flush costing about 3 blocks of counting
		const uint64_t G = (uint64_t) h->sm_count, rem = h->L.C % G, nblk = (h->L.P + CP_BLOCK - 1) / CP_BLOCK;
		uint32_t S = 1;
		double best = 0.0;
		for (uint32_t c = 1; rem && c <= 256 && c <= nblk; ++c) {
			const double cost = (double) ((c * rem + G - 1) / G) * ((double) nblk / c + 3.0);
			if (c == 1 || cost < best) {
				best = cost;
				S = c;
			}
		}
		qvz_cond_counts_planes_kernel<<<h->sm_count, CP_THREADS, smem, h->stream>>>(h->L, h->Xb, A, S, counts_dev);
		QVZ_LAUNCHED(h);
		QVZ_CUDA(h, cudaGetLastError());
		return QVZ_OK;
	}
	const size_t slice_bytes = ((size_t) A * A + CC_PAD) * sizeof(uint32_t);
	uint32_t G = (uint32_t) ((200 * 1024) / (4 * slice_bytes));       // clusters per pass that fit shared memory (2 at A = 72, 7 at A = 42)
	if (G > K) G = K;
	const size_t smem = (size_t) G * 4 * slice_bytes;
	// chunk of slots per CTA: sized for ~16 CTAs per SM over the whole grid so that zeroing/flushing the slices stays negligible
	const uint64_t want = (uint64_t) h->sm_count * 16 / h->L.C4 + 1;
	uint64_t c = (h->L.P + want - 1) / want;
	c = (c + 4095) / 4096 * 4096;
	const uint32_t chunk = (uint32_t) (c < 61440 ? 61440 : c);
	QVZ_CUDA(h, cudaMemsetAsync(counts_dev, 0, qvz_gpu_cond_counts_len(K, h->L.C) * sizeof(uint32_t), h->stream));
	dim3 grid(h->L.C4, (unsigned) ((h->L.P + chunk - 1) / chunk), (K + G - 1) / G);
	// the last cluster group may be partial: out-of-range ids are skipped by g >= G / kbase + g >= K
	QVZ_CUDA(h, cudaFuncSetAttribute(qvz_cond_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	qvz_cond_counts_kernel<<<grid, CC_THREADS, smem, h->stream>>>(h->L, h->Xw, h->cl, K, G, A, chunk, counts_dev);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}
