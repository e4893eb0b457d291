// kmeans.cu -- one k-means iteration as one fused pass over the resident rows.
//
// Reference (src/cluster.c): cluster_lines (:65-74) -> do_cluster_assignment (:136-144) ->
// find_distance (:176-187) + assign_cluster (:149-171), then recalculate_means (:80-131), which walks
// all rows a second time.  Here both happen in one kernel launch per iteration:
//
//   distance:  find_distance sums uint32 squares into a double; every partial sum is an integer
//              < 2^53 so integer arithmetic gives the same values and the same '<' outcomes.
//              argmin_k sum (x-m_k)^2  =  argmin_k ( sum m_k^2  -  2 * sum x*m_k )   (sum x^2 is common),
//              and sum x*m_k over 4 columns is one dp4a on the packed words.  Ties: strict '<' in
//              cluster order => lowest id wins, as in assign_cluster.
//   sums:      accumulator[cluster][col] += byte  (uint64 in the reference) -> the tile's rows are counting-
//              sorted by cluster in shared memory and summed column-word-wise as packed 16-bit halves,
//              CTA-level uint32 partials in shared memory, one 64-bit global atomic per (cluster, column,
//              CTA).  Integer sums are order independent.
//   recenter:  mean = (uint8)(sum / count), moved_k = sum (new-old)^2 -- tiny second kernel, which also evaluates
//              do_kmeans_clustering's loop condition so that the host never waits for it.
// Two implementations of the pass: up to 8 clusters (every BASELINE configuration) qvz_kmeans_assign_mma_kernel below --
// register-blocked distances, column sums as small integer matrix products on the tensor cores, and from the second
// iteration of a run on only the rows that changed cluster are added to / taken from the running sums; 9..16 clusters the
// counting-sort kernel that follows (more: kmeans_wide.cu).
#include <stdlib.h>

#include <type_traits>

#include "qvz_internal.cuh"

// Shared-memory plan of one CTA (R = blockDim.x rows per tile):
//   bar  tile[C4][R+4]  stile[C4][R+1]  mean4[C4][KP]  acc[K][C4*4]  msq[KP]  cnt[K]  off[(NW+2)*K]
// Per tile of R slots:
//   load     the C4 rows of the tile (R*4 contiguous bytes of each word column) arrive by TMA bulk copies
//            (cp.async.bulk + mbarrier); the copy of tile n+1 is issued as soon as tile n has been scattered, so
//            it flies under phase 2 and HBM never idles on a barrier;
//   phase 1  thread <-> row: its packed words come from shared memory (lane-consecutive: conflict free), the K
//            centroid words of a column are one or two 16-byte broadcast loads, K dp4a per word, argmin;
//   sort     counting sort of the tile's rows by cluster (ballot/popc ranks + a tiny scan) gives every row its
//            sorted position; the row's words are scattered there (stile: rows of one cluster are contiguous);
//   phase 2  thread <-> (column word, part): walks the contiguous row range of each cluster, one dp4a per
//            (word, column) against a one-hot byte selector (exact 32-bit sums, no unpacking), then one shared
//            atomic per (cluster, column).  Pitch R+1 makes the 32 column words of a warp fall in 32 banks.
//            (A warp-REDUX formulation over sorted rows was measured 2x slower: REDUX.SUM is a slow path.)
__device__ __forceinline__ uint32_t km_smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

// The packed centroids [C4][KP] in the CONSTANT bank, one slot per handle: every lane of a warp wants the same centroid
// word at the same time, which the constant cache serves as a uniform operand of the dp4a -- no shared-memory wavefronts
// (as 16-byte broadcast loads they were two thirds of the kernel's LSU traffic).  The update kernel leaves the packed
// copy in global memory (means_t); the launch code copies it here device-to-device, in stream order.
#define KM_CM_SLOTS 8
#define KM_CM_WORDS 1024
__constant__ uint32_t c_means[KM_CM_SLOTS * KM_CM_WORDS];

// INCR = a later iteration of a run: sums[] already holds the column sums of the previous assignment, so only the rows
// whose cluster CHANGED move their bytes (+ new cluster, - old cluster, signed 32-bit partials): no sort, no scatter, no
// second tile.  Integer sums: the totals are exactly those of a full recount.
template <int KT, bool INCR, bool CM>
__global__ void __launch_bounds__(QVZ_THREADS)
qvz_kmeans_assign_kernel(qvz_layout L, const uint32_t *__restrict__ Xw, uint8_t *__restrict__ cl,
                         const uint32_t *__restrict__ means_w, const uint32_t *__restrict__ means_sq,
                         uint32_t Krt, unsigned long long *__restrict__ sums, const uint32_t *__restrict__ ctl, uint32_t cm_slot)
{
	if (ctl[QVZ_CTL_DONE]) return;                       // the run has converged: this launch was enqueued speculatively (abi.cu)
	constexpr int KMAX = KT > 0 ? KT : QVZ_MAX_K;
	constexpr uint32_t KP = (KMAX + 3) & ~3;             // clusters padded to whole uint4s
	const uint32_t K = KT > 0 ? (uint32_t) KT : Krt;
	const uint32_t C4 = L.C4, R = blockDim.x, NW = R >> 5;
	const uint32_t pitch = R + 4, spitch = R + 1;
	extern __shared__ __align__(16) uint32_t sm[];
	uint64_t *bar = (uint64_t *) sm;
	uint32_t *tile = sm + 4;                             // [C4][pitch]   (TMA destination: 16-byte aligned rows)
	uint32_t *stile = tile + C4 * pitch;                 // [C4][spitch]  rows sorted by cluster
	uint32_t *mean4 = INCR ? stile : stile + ((C4 * spitch + 3) & ~3u);   // [C4][KP]   (INCR has no sorted tile)
	uint32_t *acc = mean4 + C4 * KP;                     // [K][C4*4]
	uint32_t *msq = acc + K * C4 * 4;                    // [KP]
	uint32_t *cnt = msq + KP;                            // [K]
	uint32_t *off = cnt + K;                             // [(NW+1)][K] exclusive offsets, then [K] totals

	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (!CM)
		for (uint32_t i = tid; i < C4 * KP; i += R) {
			const uint32_t c4 = i / KP, k = i - c4 * KP;
			mean4[i] = k < K ? means_w[k * C4 + c4] : 0u;
		}
	for (uint32_t i = tid; i < K * C4 * 4; i += R) acc[i] = 0;
	if (tid < KP) msq[tid] = tid < K ? means_sq[tid] : 0u;
	if (tid < K) cnt[tid] = 0;
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(km_smem_u32(bar)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	const uint32_t nparts = (R / C4) ? (R / C4) : 1;
	const uint64_t tiles = L.P / R;                      // P % 4096 == 0 and R | 256
	const uint32_t tile_bytes = C4 * R * 4;
	auto fetch = [&](uint64_t t) {                       // word column c4 of tile t: R*4 contiguous bytes
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the tile's earlier generic reads are done (barrier before)
		if (tid == 0)
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(km_smem_u32(bar)), "r"(tile_bytes) : "memory");
		for (uint32_t c4 = tid; c4 < C4; c4 += R)
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			             ::"r"(km_smem_u32(tile + c4 * pitch)), "l"(Xw + (uint64_t) c4 * L.P + t * R), "r"(R * 4),
			               "r"(km_smem_u32(bar)) : "memory");
	};
	if (blockIdx.x < tiles) fetch(blockIdx.x);

	uint32_t it = 0;
	for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
		{                                                // wait for this tile's bytes
			const uint32_t parity = it & 1, addr = km_smem_u32(bar);
			asm volatile(
			    "{\n"
			    ".reg .pred p;\n"
			    "KW_%=:\n"
			    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
			    "@p bra KD_%=;\n"
			    "bra KW_%=;\n"
			    "KD_%=:\n"
			    "}" ::"r"(addr), "r"(parity) : "memory");
		}
		const uint64_t p = t * R + tid;
		const uint32_t old = cl[p];
		const bool valid = old != QVZ_NO_LINE;
		uint32_t best = 0;
		{
			uint32_t D[KP];
#pragma unroll
			for (uint32_t k = 0; k < KP; ++k) D[k] = 0;
			const uint32_t *tp = tile + tid;
#pragma unroll 4
			for (uint32_t c4 = 0; c4 < C4; ++c4) {
				const uint32_t w = tp[c4 * pitch];
				if (CM) {                                    // centroid words straight from the constant bank (uniform index)
					const uint32_t *m = c_means + cm_slot * KM_CM_WORDS + c4 * KP;
#pragma unroll
					for (uint32_t k = 0; k < (uint32_t) KMAX; ++k)
						if (KT > 0 || k < K) D[k] = __dp4a(w, m[k], D[k]);
				} else {
					const uint4 *m = (const uint4 *) (mean4 + c4 * KP);
#pragma unroll
					for (uint32_t g = 0; g < KP / 4; ++g) {
						if (KT > 0 || 4 * g < K) {
							const uint4 mm = m[g];
							D[4 * g + 0] = __dp4a(w, mm.x, D[4 * g + 0]);
							D[4 * g + 1] = __dp4a(w, mm.y, D[4 * g + 1]);
							D[4 * g + 2] = __dp4a(w, mm.z, D[4 * g + 2]);
							D[4 * g + 3] = __dp4a(w, mm.w, D[4 * g + 3]);
						}
					}
				}
			}
			int bestv = (int) msq[0] - 2 * (int) D[0];
#pragma unroll
			for (uint32_t k = 1; k < (uint32_t) KMAX; ++k) {
				if (KT > 0 || k < K) {
					const int v = (int) msq[k] - 2 * (int) D[k];
					if (v < bestv) {                 // strict '<': lowest cluster id wins ties (assign_cluster)
						bestv = v;
						best = k;
					}
				}
			}
		}
		if (INCR) {
			const bool changed = valid && best != old;
			if (changed) cl[p] = (uint8_t) best;
			uint32_t mask = __ballot_sync(0xFFFFFFFFu, changed);
			while (mask) {                               // one changed row at a time, the warp's lanes over its column words
				const uint32_t src = __ffs(mask) - 1;
				mask &= mask - 1;
				const uint32_t nk = __shfl_sync(0xFFFFFFFFu, best, src), ok = __shfl_sync(0xFFFFFFFFu, old, src);
				const uint32_t *tr = tile + (warp * 32 + src);
				for (uint32_t c4 = lane; c4 < C4; c4 += 32) {
					const uint32_t w = tr[c4 * pitch];
					uint32_t *an = acc + (nk * C4 + c4) * 4, *ao = acc + (ok * C4 + c4) * 4;
#pragma unroll
					for (uint32_t j = 0; j < 4; ++j) {
						const uint32_t b = (w >> (8 * j)) & 0xFFu;
						if (b) {
							atomicAdd(an + j, b);
							atomicSub(ao + j, b);
						}
					}
				}
				if (lane == 0) {
					atomicAdd(&cnt[nk], 1u);
					atomicSub(&cnt[ok], 1u);
				}
			}
			__syncthreads();                             // the tile has been consumed
			if (t + gridDim.x < tiles) fetch(t + gridDim.x);
			continue;
		}
		if (valid) cl[p] = (uint8_t) best;
		const uint32_t key = valid ? best : 0;           // empty slots hold zero words: harmless in any list

		// counting sort by cluster
		uint32_t rank = 0;
		for (uint32_t k = 0; k < K; ++k) {
			const uint32_t m = __ballot_sync(0xFFFFFFFFu, key == k);
			if (key == k) rank = __popc(m & ((1u << lane) - 1));
			if (lane == 0) off[(warp + 1) * K + k] = __popc(m);
			const uint32_t nv = __popc(__ballot_sync(0xFFFFFFFFu, valid && best == k));
			if (lane == 0 && nv) atomicAdd(&cnt[k], nv);
		}
		__syncthreads();
		if (tid < K) {                                   // column k: running sum over warps, totals in the last row
			uint32_t run = 0;
			for (uint32_t w = 0; w < NW; ++w) {
				const uint32_t c = off[(w + 1) * K + tid];
				off[w * K + tid] = run;
				run += c;
			}
			off[NW * K + tid] = run;
		}
		__syncthreads();
		{                                                // scatter this row to its sorted position
			uint32_t spos = off[warp * K + key] + rank;
			for (uint32_t k = 0; k < key; ++k) spos += off[NW * K + k];
			uint32_t src = km_smem_u32(tile + tid), dst = km_smem_u32(stile + spos);
			const uint32_t sstep = pitch * 4, dstep = spitch * 4;
#pragma unroll 4
			for (uint32_t c4 = 0; c4 < C4; ++c4) {           // explicit shared addresses: 4 instructions per word
				uint32_t w;
				asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(src));
				asm volatile("st.shared.u32 [%0], %1;" ::"r"(dst), "r"(w) : "memory");
				src += sstep;
				dst += dstep;
			}
		}
		__syncthreads();                                 // the tile has been consumed: refill it under phase 2
		if (t + gridDim.x < tiles) fetch(t + gridDim.x);

		// column sums over the cluster-contiguous row ranges
		for (uint32_t item = tid; item < C4 * nparts; item += R) {
			const uint32_t c4 = item % C4, part = item / C4;
			const uint32_t *tc = stile + c4 * spitch;
			uint32_t seg = 0;
			for (uint32_t k = 0; k < K; ++k) {
				const uint32_t end = seg + off[NW * K + k];
				uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
				uint32_t j = seg + part;
				for (; j + 3 * nparts < end; j += 4 * nparts) {
					const uint32_t w0 = tc[j], w1 = tc[j + nparts], w2 = tc[j + 2 * nparts], w3 = tc[j + 3 * nparts];
					a0 = __dp4a(w0, 0x00000001u, a0); a1 = __dp4a(w0, 0x00000100u, a1); a2 = __dp4a(w0, 0x00010000u, a2); a3 = __dp4a(w0, 0x01000000u, a3);
					a0 = __dp4a(w1, 0x00000001u, a0); a1 = __dp4a(w1, 0x00000100u, a1); a2 = __dp4a(w1, 0x00010000u, a2); a3 = __dp4a(w1, 0x01000000u, a3);
					a0 = __dp4a(w2, 0x00000001u, a0); a1 = __dp4a(w2, 0x00000100u, a1); a2 = __dp4a(w2, 0x00010000u, a2); a3 = __dp4a(w2, 0x01000000u, a3);
					a0 = __dp4a(w3, 0x00000001u, a0); a1 = __dp4a(w3, 0x00000100u, a1); a2 = __dp4a(w3, 0x00010000u, a2); a3 = __dp4a(w3, 0x01000000u, a3);
				}
				for (; j < end; j += nparts) {
					const uint32_t w0 = tc[j];
					a0 = __dp4a(w0, 0x00000001u, a0); a1 = __dp4a(w0, 0x00000100u, a1); a2 = __dp4a(w0, 0x00010000u, a2); a3 = __dp4a(w0, 0x01000000u, a3);
				}
				if (a0 | a1 | a2 | a3) {
					uint32_t *a = acc + (k * C4 + c4) * 4;
					atomicAdd(a + 0, a0);
					atomicAdd(a + 1, a1);
					atomicAdd(a + 2, a2);
					atomicAdd(a + 3, a3);
				}
				seg = end;
			}
		}
		__syncthreads();                                 // stile and off are reused by the next tile
	}
	__syncthreads();
	for (uint32_t i = tid; i < K * C4 * 4; i += R) {     // INCR partials are signed: sign-extend (two's complement add)
		const uint32_t k = i / (C4 * 4), c = i - k * C4 * 4;
		if (c < L.C && acc[i]) atomicAdd(&sums[(uint64_t) k * L.C + c], INCR ? (unsigned long long) (long long) (int) acc[i] : (unsigned long long) acc[i]);
	}
	if (tid < K && cnt[tid]) atomicAdd(&sums[(uint64_t) K * L.C + tid], INCR ? (unsigned long long) (long long) (int) cnt[tid] : (unsigned long long) cnt[tid]);
}

// ---- register-blocked iteration with tensor-core column sums (K <= 8) -------------------------------------------
// The tile of 4*NT slots arrives by TMA bulk copies as above, but
//   distances   a thread owns 4 CONSECUTIVE slots: one 16-byte shared-memory load per word column brings 4 rows, the K
//               centroid words of the column (one or two 16-byte broadcasts) serve all 4, and the 4 x K dot products
//               stay in registers: 1 + KP/4 LDS.128 and 4K dp4a per 16 symbols (~1.6 instructions per symbol);
//   sums        recalculate_means' accumulator[cluster][col] += byte (src/cluster.c:96-104) is the product of a
//               (cluster x row) coefficient matrix with the (row x column) byte matrix, and a small integer one:
//               mma.sync.m16n8k32 (s8 x u8 -> s32, exact) does 32 rows x 8 columns per instruction.  A = +1 where the
//               row's NEW cluster is m, and in a later iteration of a run additionally -1 where its OLD cluster is m: rows
//               that stay put have an all-zero column and the running sums of the run only receive the rows that moved
//               (quads of rows without a change are not even read).  B comes from the same tile: a lane reads the 4
//               consecutive rows of one word column with one LDS.128 and turns them into four column words with 8 byte
//               permutes (4 x 4 byte transpose), which feed 4 MMAs.  One more B column of ones counts the lines
//               (cluster_t.count).  32-bit partials per CTA in shared memory, 64-bit global atomics at the end.
// The counting-sort kernel above remains for K > 8.
__device__ __forceinline__ void km_mma_s8u8(int (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
	asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
	             : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
	             : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}

// rows (r0..r3) x bytes (columns j)  ->  T[j] = byte j of r0, r1, r2, r3
__device__ __forceinline__ void km_transpose4(const uint4 &w, uint32_t (&T)[4]) {
	const uint32_t t0 = __byte_perm(w.x, w.y, 0x5140), t1 = __byte_perm(w.z, w.w, 0x5140);
	const uint32_t t2 = __byte_perm(w.x, w.y, 0x7362), t3 = __byte_perm(w.z, w.w, 0x7362);
	T[0] = __byte_perm(t0, t1, 0x5410);
	T[1] = __byte_perm(t0, t1, 0x7632);
	T[2] = __byte_perm(t2, t3, 0x5410);
	T[3] = __byte_perm(t2, t3, 0x7632);
}

static __host__ __device__ __forceinline__ uint32_t km_groups(uint32_t C4) { return (C4 + 8) / 8; }      // groups of 8 word columns incl. >= 1 padding column (the line counter)

template <int KT, int NT, int RPT, bool FULL, int NBUF>  // NT threads walk tiles of RPT*NT slots, RPT (2 or 4) consecutive slots per thread;
__global__ void __launch_bounds__(NT)                    // FULL = first iteration of a run (every row counts); NBUF tile buffers (1 or 2)
qvz_kmeans_assign_mma_kernel(qvz_layout L, const uint32_t *__restrict__ Xw, uint8_t *__restrict__ cl,
                             const uint32_t *__restrict__ means_t, const uint32_t *__restrict__ means_sq,
                             unsigned long long *__restrict__ sums, const uint32_t *__restrict__ ctl)
{
	if (ctl[QVZ_CTL_DONE]) return;                       // the run has converged: this launch was enqueued speculatively (abi.cu)
	constexpr uint32_t K = KT, KP = (KT + 3) & ~3;
	constexpr uint32_t ROWS = RPT * NT, PITCH = ROWS + 4;    // pitch in words: tile rows stay 16-byte aligned (TMA destination)
	constexpr uint32_t QPW = 8 * RPT, NB = RPT;          // quads of rows per warp, batches of 8 quads per warp
	typedef typename std::conditional<RPT == 4, uint4, typename std::conditional<RPT == 2, uint2, uint32_t>::type>::type xvec;
	typedef typename std::conditional<RPT == 4, uint32_t, typename std::conditional<RPT == 2, uint16_t, uint8_t>::type>::type idvec;
	extern __shared__ __align__(16) uint32_t sm[];
	const uint32_t C4 = L.C4, G = km_groups(C4), ACCW = G * 32;     // ACCW = accumulator words per cluster (column 4*C4 = line count)
	uint64_t *bar = (uint64_t *) sm;                     // [NBUF]
	uint32_t *tile0 = sm + 4;                            // [NBUF][C4][PITCH]: with two buffers the next tile lands while this one is worked on
	uint32_t *mean4 = tile0 + NBUF * C4 * PITCH;         // [C4][KP]
	uint32_t *acc = mean4 + C4 * KP;                     // [K][ACCW] signed partials
	uint32_t *msq = acc + K * ACCW;                      // [KP]
	uint16_t *idn = (uint16_t *) (msq + KP);             // [ROWS/4] new ids of a quad of rows, one NIBBLE per row (0xF = no line)
	uint16_t *ido = (uint16_t *) (msq + KP + ROWS / 4);  // [ROWS/4] old ids
	uint32_t *qlist = msq + KP + 2 * (ROWS / 4);          // [ROWS/4] per warp: its quads that hold a changed row, compacted
	const uint32_t tid = threadIdx.x, lane = tid & 31, wq = (tid >> 5) * QPW;
	for (uint32_t i = tid; i < C4 * KP; i += NT) mean4[i] = means_t[i];      // means_t is [C4][KP], padding centroids zero
	for (uint32_t i = tid; i < K * ACCW; i += NT) acc[i] = 0;
	if (tid < K) msq[tid] = means_sq[tid];
	if (tid == 0) {
		for (uint32_t b = 0; b < NBUF; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(km_smem_u32(bar + b)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	const uint64_t tiles = L.P / ROWS;                   // P % 4096 == 0
	const uint32_t tile_bytes = C4 * ROWS * 4;
	auto fetch = [&](uint64_t t, uint32_t b) {           // word column c4 of tile t: ROWS*4 contiguous bytes -> buffer b
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the buffer's earlier generic reads are done (barrier before)
		if (tid == 0) {
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(km_smem_u32(bar + b)), "r"(tile_bytes) : "memory");
			const uint32_t *src = Xw + t * ROWS;
			uint32_t dst = km_smem_u32(tile0 + b * C4 * PITCH);
			for (uint32_t c4 = 0; c4 < C4; ++c4, src += L.P, dst += PITCH * 4)
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				             ::"r"(dst), "l"(src), "r"(ROWS * 4), "r"(km_smem_u32(bar + b)) : "memory");
		}
	};
	for (uint32_t b = 0; b < NBUF; ++b)
		if (blockIdx.x + (uint64_t) b * gridDim.x < tiles) fetch(blockIdx.x + (uint64_t) b * gridDim.x, b);

	// this lane's row of the coefficient matrix = cluster m = lane/4: byte j of the pair (oh_lo, oh_hi) is 1 iff j == m, so one
	// byte permute with the 4 id nibbles of a quad as selector gives its 4 coefficients (nibble 0xF replicates the sign of
	// byte 7, i.e. 0: a slot without a line, or the "old" side of the first iteration, belongs to no cluster)
	const uint32_t oh_lo = (lane >> 2) < 4 ? 1u << (8 * (lane >> 2)) : 0u, oh_hi = (lane >> 2) >= 4 ? 1u << (8 * ((lane >> 2) - 4)) : 0u;
	auto coef4 = [&](uint32_t quad) {                    // (+1 new cluster, -1 old cluster) per row of the quad, as 4 signed bytes
		uint32_t en, eo;
		asm("prmt.b32 %0, %1, %2, %3;" : "=r"(en) : "r"(oh_lo), "r"(oh_hi), "r"((uint32_t) idn[quad]));
		asm("prmt.b32 %0, %1, %2, %3;" : "=r"(eo) : "r"(oh_lo), "r"(oh_hi), "r"((uint32_t) ido[quad]));
		return ((en | 0x80808080u) - eo) ^ 0x80808080u;  // per-byte en - eo: no borrow can cross a byte
	};
	const uint32_t q = lane & 3;
	uint32_t it = 0;
	for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
		const uint64_t p = t * ROWS + RPT * tid;         // this thread's slots
		const uint32_t oldw = __ldg((const idvec *) (cl + p));
		const uint32_t buf = NBUF == 2 ? (it & 1) : 0;
		const uint32_t *tile = tile0 + buf * C4 * PITCH;
		{                                                // wait for this tile's bytes
			const uint32_t parity = (NBUF == 2 ? (it >> 1) : it) & 1, addr = km_smem_u32(bar + buf);
			asm volatile(
			    "{\n"
			    ".reg .pred p;\n"
			    "KI_%=:\n"
			    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
			    "@p bra KJ_%=;\n"
			    "bra KI_%=;\n"
			    "KJ_%=:\n"
			    "}" ::"r"(addr), "r"(parity) : "memory");
		}
		uint32_t neww = oldw;
		{
			uint32_t D[RPT][K];
#pragma unroll
			for (uint32_t r = 0; r < RPT; ++r)
#pragma unroll
				for (uint32_t k = 0; k < K; ++k) D[r][k] = 0;
			const xvec *tp = (const xvec *) (tile + RPT * tid);
#pragma unroll 4
			for (uint32_t c4 = 0; c4 < C4; ++c4) {
				const xvec xv = tp[c4 * (PITCH / RPT)];
				uint32_t x[RPT];
				if constexpr (RPT == 4) { x[0] = xv.x; x[1] = xv.y; x[2] = xv.z; x[3] = xv.w; }
				else if constexpr (RPT == 2) { x[0] = xv.x; x[1] = xv.y; }
				else x[0] = xv;
				const uint4 *m = (const uint4 *) (mean4 + c4 * KP);
#pragma unroll
				for (uint32_t g = 0; g < KP / 4; ++g) {
					const uint4 mm = m[g];
					const uint32_t mk[4] = {mm.x, mm.y, mm.z, mm.w};
#pragma unroll
					for (uint32_t j = 0; j < 4; ++j)
						if (4 * g + j < K)
#pragma unroll
							for (uint32_t r = 0; r < RPT; ++r) D[r][4 * g + j] = __dp4a(x[r], mk[j], D[r][4 * g + j]);
				}
			}
#pragma unroll
			for (uint32_t r = 0; r < RPT; ++r) {
				uint32_t best = 0;
				int bestv = (int) msq[0] - 2 * (int) D[r][0];
#pragma unroll
				for (uint32_t k = 1; k < K; ++k) {
					const int v = (int) msq[k] - 2 * (int) D[r][k];
					if (v < bestv) {                     // strict '<': lowest cluster id wins ties (assign_cluster)
						bestv = v;
						best = k;
					}
				}
				if (((oldw >> (8 * r)) & 0xFFu) != QVZ_NO_LINE) neww = (neww & ~(0xFFu << (8 * r))) | (best << (8 * r));
			}
		}
		if (neww != oldw) *(idvec *) (cl + p) = (idvec) neww;
		// ids -> nibbles: byte r of neww / oldw -> nibble r of this thread's part of the quad
		if constexpr (RPT == 4) {
			idn[tid] = (uint16_t) ((neww & 0xFu) | ((neww >> 4) & 0xF0u) | ((neww >> 8) & 0xF00u) | ((neww >> 12) & 0xF000u));
			ido[tid] = FULL ? (uint16_t) 0xFFFFu : (uint16_t) ((oldw & 0xFu) | ((oldw >> 4) & 0xF0u) | ((oldw >> 8) & 0xF00u) | ((oldw >> 12) & 0xF000u));
		} else if constexpr (RPT == 2) {
			((uint8_t *) idn)[tid] = (uint8_t) ((neww & 0xFu) | ((neww >> 4) & 0xF0u));
			((uint8_t *) ido)[tid] = FULL ? (uint8_t) 0xFFu : (uint8_t) ((oldw & 0xFu) | ((oldw >> 4) & 0xF0u));     // first iteration: nothing to take back
		} else {                                         // one row per thread: the 4 lanes of a quad put their nibbles together
			uint32_t vn = (neww & 0xFu) << (4 * (lane & 3)), vo = (FULL ? 0xFu : (oldw & 0xFu)) << (4 * (lane & 3));
			vn |= __shfl_xor_sync(0xFFFFFFFFu, vn, 1);
			vo |= __shfl_xor_sync(0xFFFFFFFFu, vo, 1);
			vn |= __shfl_xor_sync(0xFFFFFFFFu, vn, 2);
			vo |= __shfl_xor_sync(0xFFFFFFFFu, vo, 2);
			if ((lane & 3) == 0) {
				idn[tid >> 2] = (uint16_t) vn;
				ido[tid >> 2] = (uint16_t) vo;
			}
		}
		__syncwarp();
		// lanes < QPW stand for this warp's quads of rows: a quad contributes to the sums if one of its rows changed cluster
		// (measured alternatives: pooling the quads of all warps of the CTA so that late iterations run fewer, fuller batches
		// gains 0.4-0.8 ms in the iterations where < 1 % of the rows move and loses 0.5-2.4 ms in all others; so does a
		// scalar path -- one shared atomic per byte -- for warps in which at most 1..4 rows moved: the mere presence of that
		// code costs the other iterations 0.4-0.7 ms each)
		const bool mine = lane < QPW && (FULL || idn[wq + lane] != ido[wq + lane]);
		const uint32_t qmask = __ballot_sync(0xFFFFFFFFu, mine);
		if (mine) qlist[wq + __popc(qmask & ((1u << lane) - 1))] = wq + lane;
		__syncwarp();
		const uint32_t nq = __popc(qmask), nb = (nq + 7) >> 3;        // batches of 8 quads = 32 rows = one MMA k-extent
		if (nb) {
			// coefficient fragments and tile offsets of the batches of this warp
			uint32_t a0[NB], a2[NB], off0[NB], off2[NB];
#pragma unroll
			for (uint32_t b = 0; b < NB; ++b) {
				const uint32_t s0 = 8 * b + q, s2 = s0 + 4;
				a0[b] = a2[b] = 0;
				off0[b] = off2[b] = 4 * wq;
				if (s0 < nq) {
					const uint32_t src = qlist[wq + s0];
					a0[b] = coef4(src);
					off0[b] = 4 * src;
				}
				if (s2 < nq) {
					const uint32_t src = qlist[wq + s2];
					a2[b] = coef4(src);
					off2[b] = 4 * src;
				}
			}
#pragma unroll 2
			for (uint32_t g = 0; g < G; ++g) {           // (two groups in flight: the chain LDS -> PRMT -> IMMA -> ATOMS of one group is latency, not work)
				const uint32_t c4 = 8 * g + (lane >> 2);
				int c[4][4];
#pragma unroll
				for (uint32_t j = 0; j < 4; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0;
				const uint32_t *tc = tile + c4 * PITCH;
#pragma unroll
				for (uint32_t b = 0; b < NB; ++b) {
					if (b < nb) {
						uint32_t T0[4], T2[4];
						if (c4 < C4) {
							km_transpose4(*(const uint4 *) (tc + off0[b]), T0);
							km_transpose4(*(const uint4 *) (tc + off2[b]), T2);
						} else {                         // padding columns: the first one counts lines
							T0[0] = T2[0] = c4 == C4 ? 0x01010101u : 0u;
							T0[1] = T0[2] = T0[3] = T2[1] = T2[2] = T2[3] = 0u;
						}
#pragma unroll
						for (uint32_t j = 0; j < 4; ++j) km_mma_s8u8(c[j], a0[b], a2[b], T0[j], T2[j]);
					}
				}
				if ((lane >> 2) < K) {                   // rows of C = clusters; columns n = 2*(lane%4), +1 <-> word columns 8g + n
					uint32_t *ap = acc + (lane >> 2) * ACCW + 4 * (8 * g + 2 * q);
#pragma unroll
					for (uint32_t j = 0; j < 4; ++j) {
						if (c[j][0]) atomicAdd(ap + j, (uint32_t) c[j][0]);
						if (c[j][1]) atomicAdd(ap + 4 + j, (uint32_t) c[j][1]);
					}
				}
			}
		}
		__syncthreads();                                 // the tile has been consumed
		if (t + (uint64_t) NBUF * gridDim.x < tiles) fetch(t + (uint64_t) NBUF * gridDim.x, buf);
	}
	__syncthreads();
	for (uint32_t i = tid; i < K * ACCW; i += NT) {      // signed partials: sign-extend (two's complement add)
		const uint32_t k = i / ACCW, c = i - k * ACCW;
		if (!acc[i]) continue;
		if (c < L.C) atomicAdd(&sums[(uint64_t) k * L.C + c], (unsigned long long) (long long) (int) acc[i]);
		else if (c == 4 * C4) atomicAdd(&sums[(uint64_t) K * L.C + k], (unsigned long long) (long long) (int) acc[i]);
	}
}

static size_t mma_smem(uint32_t K, uint32_t C4, uint32_t rows, uint32_t NT, uint32_t nbuf = 1) {
	const uint32_t KP = (K + 3) & ~3u;
	(void) NT;
	return (4 + (size_t) nbuf * C4 * (rows + 4) + (size_t) C4 * KP + (size_t) K * km_groups(C4) * 32 + KP + 3 * (rows / 4) + 8) * sizeof(uint32_t);
}

template <int KT, int NT, int RPT, bool FULL, int NBUF>
static void launch_mma_nb(qvz_gpu *h, int64_t *target) {
	auto kern = qvz_kmeans_assign_mma_kernel<KT, NT, RPT, FULL, NBUF>;
	const size_t smem = mma_smem(KT, h->L.C4, RPT * NT, NT, NBUF);
	cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	int per_sm = 0;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
	if (per_sm < 1) per_sm = 1;
	if (const char *e = getenv("QVZ_KM_CTAS")) per_sm = atoi(e) > 0 ? atoi(e) : per_sm;      // tuning knob
	const uint64_t blocks = h->L.P / (RPT * NT);
	const uint64_t cap = (uint64_t) h->sm_count * per_sm;
	kern<<<(unsigned) (blocks < cap ? blocks : cap), NT, smem, h->stream>>>(h->L, h->Xw, h->cl, h->means_t, h->means_sq,
	                                                                       (unsigned long long *) target, h->km_ctl);
}

template <int KT, int NT, int RPT, bool FULL>
static void launch_mma_nt(qvz_gpu *h, int64_t *target) {
	const char *e = getenv("QVZ_KM_NBUF");           // 2: double-buffered tiles (half as many CTAs per SM)
	if (e && atoi(e) == 2 && mma_smem(KT, h->L.C4, RPT * NT, NT, 2) <= 220 * 1024) launch_mma_nb<KT, NT, RPT, FULL, 2>(h, target);
	else launch_mma_nb<KT, NT, RPT, FULL, 1>(h, target);
}

// shape of a tile: QVZ_KM_SHAPE = <rows per thread><threads>, e.g. 2128 = 2 rows x 128 threads (256-slot tiles, 4 warps)
static uint32_t mma_shape(uint32_t K, uint32_t C4) {
	if (const char *e = getenv("QVZ_KM_SHAPE")) return (uint32_t) atoi(e);
	return mma_smem(K, C4, 256, 128) <= 56 * 1024 ? 2128 : 264;       // long rows: 128-slot tiles
}

template <int KT, bool FULL>
static void launch_mma_shape(qvz_gpu *h, int64_t *target) {
	switch (mma_shape(KT, h->L.C4)) {
	case 4128: launch_mma_nt<KT, 128, 4, FULL>(h, target); break;
	case 464: launch_mma_nt<KT, 64, 4, FULL>(h, target); break;
	case 432: launch_mma_nt<KT, 32, 4, FULL>(h, target); break;
	case 2256: launch_mma_nt<KT, 256, 2, FULL>(h, target); break;
	case 2128: launch_mma_nt<KT, 128, 2, FULL>(h, target); break;
	case 1256: launch_mma_nt<KT, 256, 1, FULL>(h, target); break;
	default: launch_mma_nt<KT, 64, 2, FULL>(h, target); break;
	}
}

template <int KT>
static void launch_mma(qvz_gpu *h, int64_t *target, bool full) {
	if (full) launch_mma_shape<KT, true>(h, target);
	else launch_mma_shape<KT, false>(h, target);
}

// K == 1: assign_cluster has nothing to compare -- every line lands in cluster 0 -- so one iteration is just
// recalculate_means' column sums (src/cluster.c:96-104).  Those sums are marginals of the conditional-count table
// that calculate_statistics needs right after (sum over prev and value of (value + '!') * count), so the rows are
// not read for k-means at all: the counting pass runs here, the exact integer sums are folded out of its table,
// and qvz_gpu_cond_counts hands out the same table (abi.cu: counts_cached).  One pass over the rows serves both
// stages; the sums are integers, so nothing about the centroids or the `moved` values changes.
__global__ void __launch_bounds__(256)
qvz_kmeans_k1_ids_kernel(uint64_t P, uint8_t *__restrict__ cl)
{
	// line->cluster = 0 for every real line (0xFF marks an empty slot); ids are < 16, so bit 7 is set only in 0xFF
	uint32_t *c32 = (uint32_t *) cl;
	for (uint64_t p = ((uint64_t) blockIdx.x * 256 + threadIdx.x) * 4; p < P; p += (uint64_t) gridDim.x * 1024) {
		const uint32_t v = c32[p >> 2];
		const uint32_t nv = ((v >> 7) & 0x01010101u) * 0xFFu;
		if (nv != v) c32[p >> 2] = nv;
	}
}

// one CTA per column: rows 1+(c-1)*72 .. +71 of the count table (row 0 for column 0), 72 counters each
__global__ void __launch_bounds__(256)
qvz_kmeans_k1_sums_kernel(uint32_t C, uint64_t n_lines, const uint32_t *__restrict__ counts,
                          unsigned long long *__restrict__ sums)
{
	__shared__ unsigned long long red[8];
	const uint32_t c = blockIdx.x, tid = threadIdx.x;
	const uint32_t *rows = counts + (c ? (uint64_t) (1 + (c - 1) * 72) * 72 : 0);
	const uint32_t cells = c ? 72u * 72u : 72u;
	unsigned long long s = 0;
	for (uint32_t i = tid; i < cells; i += 256) s += (unsigned long long) rows[i] * (i % 72u + 33u);
	for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
	if ((tid & 31) == 0) red[tid >> 5] = s;
	__syncthreads();
	if (tid == 0) {
		for (int w = 1; w < 8; ++w) s += red[w];
		sums[c] = s;
		if (c == 0) sums[C] = n_lines;               // cluster_t.count of the only cluster
	}
}

// recalculate_means (src/cluster.c:106-128) on the reduced sums; also repacks the centroids for dp4a.
// first = 1: only pack the initial centroids (initialize_kmeans_clustering copied them from rows).
__global__ void __launch_bounds__(QVZ_THREADS)
qvz_kmeans_update_kernel(uint32_t K, uint32_t C, uint32_t C4, const unsigned long long *__restrict__ sums,
                         uint8_t *__restrict__ means_b, uint32_t *__restrict__ means_w,
                         uint32_t *__restrict__ means_sq, double *__restrict__ moved,
                         int *__restrict__ flags, int first, uint32_t *__restrict__ ctl, double threshold,
                         uint32_t max_iter, double *__restrict__ moved_log, unsigned long long *__restrict__ last_counts,
                         uint32_t *__restrict__ means_t, uint32_t KP)
{
	__shared__ unsigned long long red_moved[QVZ_THREADS / 32];
	__shared__ unsigned long long red_sq[QVZ_THREADS / 32];
	if (!first && ctl[QVZ_CTL_DONE]) return;             // speculative launch after convergence
	const uint32_t iter = first ? 0u : ctl[QVZ_CTL_ITER];
	double move_max = 0.0;                               // thread 0 only
	bool empty = false;
	for (uint32_t k = 0; k < K; ++k) {
		unsigned long long count = first ? 1ull : sums[(uint64_t) K * C + k];
		if (count == 0) {
			if (threadIdx.x == 0) atomicOr(&flags[1], 1);
			count = 1;
			empty = true;
		}
		unsigned long long mv = 0, sq = 0;
		for (uint32_t c = threadIdx.x; c < C; c += QVZ_THREADS) {
			const uint32_t old = means_b[k * C + c];
			uint32_t nm = old;
			if (!first) {
				nm = (uint32_t) (sums[(uint64_t) k * C + c] / count) & 0xFFu;   // (uint8_t)(acc/count)
				const int d = (int) nm - (int) old;
				mv += (unsigned long long) (d * d);
				means_b[k * C + c] = (uint8_t) nm;
			}
			sq += (unsigned long long) nm * nm;
		}
		for (int o = 16; o > 0; o >>= 1) {
			mv += __shfl_down_sync(0xFFFFFFFFu, mv, o);
			sq += __shfl_down_sync(0xFFFFFFFFu, sq, o);
		}
		if ((threadIdx.x & 31) == 0) {
			red_moved[threadIdx.x >> 5] = mv;
			red_sq[threadIdx.x >> 5] = sq;
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			unsigned long long a = 0, b = 0;
			for (int w = 0; w < QVZ_THREADS / 32; ++w) {
				a += red_moved[w];
				b += red_sq[w];
			}
			if (!first) {
				moved[k] = (double) a;               // exact: an integer < 2^53, like the reference's double sum
				if (iter < QVZ_MAX_KMEANS_ITER) moved_log[(size_t) iter * K + k] = (double) a;
				last_counts[k] = sums[(uint64_t) K * C + k];
				if ((double) a > move_max) move_max = (double) a;
			}
			means_sq[k] = (uint32_t) b;
		}
		__syncthreads();
		for (uint32_t c4 = threadIdx.x; c4 < C4; c4 += QVZ_THREADS) {
			uint32_t w = 0;
			for (uint32_t j = 0; j < 4; ++j)
				if (4 * c4 + j < C) w |= (uint32_t) means_b[k * C + 4 * c4 + j] << (8 * j);
			means_w[k * C4 + c4] = w;
			if (means_t) means_t[c4 * KP + k] = w;       // [C4][KP]: the layout the assign kernel reads (constant bank copy)
		}
		__syncthreads();
	}
	// do_kmeans_clustering's loop condition, evaluated here so that the host never waits for it:
	//     loop = moved > cluster_threshold;  while (iter_count < MAX && loop)        (src/cluster.c:221-234)
	if (!first && threadIdx.x == 0) {
		ctl[QVZ_CTL_ITER] = iter + 1;
		if (!(move_max > threshold) || iter + 1 >= max_iter || empty) ctl[QVZ_CTL_DONE] = 1;
	}
}

static size_t assign_smem(uint32_t K, uint32_t C4, uint32_t R, bool incr = false) {
	const uint32_t NW = R / 32, KP = ((K <= 8 ? K : QVZ_MAX_K) + 3) & ~3u;
	size_t words = 4 + (size_t) C4 * (R + 4) + (incr ? 0 : (((size_t) C4 * (R + 1) + 3) & ~(size_t) 3)) + (size_t) C4 * KP + (size_t) K * C4 * 4 + KP + K + (NW + 2) * K;
	words += 4;
	return words * sizeof(uint32_t);
}

template <int KT, bool INCR, bool CM>
static void launch_assign_v(qvz_gpu *h, int64_t *sums_dev, unsigned grid, unsigned R, size_t smem) {
	auto kern = qvz_kmeans_assign_kernel<KT, INCR, CM>;
	cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	kern<<<grid, R, smem, h->stream>>>(h->L, h->Xw, h->cl, h->means_w, h->means_sq, h->km_K, (unsigned long long *) sums_dev, h->km_ctl,
	                                   (uint32_t) (h->cm_slot < 0 ? 0 : h->cm_slot));
}

template <int KT>
static void launch_assign(qvz_gpu *h, int64_t *sums_dev, unsigned grid, unsigned R, size_t smem, bool incr, bool cm) {
	if (incr) {
		if (cm) launch_assign_v<KT, true, true>(h, sums_dev, grid, R, smem);
		else launch_assign_v<KT, true, false>(h, sums_dev, grid, R, smem);
	} else {
		if (cm) launch_assign_v<KT, false, true>(h, sums_dev, grid, R, smem);
		else launch_assign_v<KT, false, false>(h, sums_dev, grid, R, smem);
	}
}

// centroid slots of the constant bank: one per live handle of this process (more handles than slots: shared-memory path)
#include <mutex>
static std::mutex cm_mutex;
static bool cm_used[KM_CM_SLOTS];
int qvz_kmeans_slot_acquire() {
	std::lock_guard<std::mutex> g(cm_mutex);
	for (int i = 0; i < KM_CM_SLOTS; ++i)
		if (!cm_used[i]) {
			cm_used[i] = true;
			return i;
		}
	return -1;
}
void qvz_kmeans_slot_release(int slot) {
	std::lock_guard<std::mutex> g(cm_mutex);
	if (slot >= 0 && slot < KM_CM_SLOTS) cm_used[slot] = false;
}
uint32_t qvz_kmeans_kp(uint32_t K) { return ((K <= 8 ? K : QVZ_MAX_K) + 3) & ~3u; }      // centroids per column word, padded (kernel: KP)

int qvz_kmeans_launch_assign(qvz_gpu *h, int64_t *sums_dev) {
	const uint32_t K = h->km_K, C4 = h->L.C4;
	const size_t sum_bytes = ((size_t) K * h->L.C + K) * sizeof(int64_t);
	if (K == 1) {
		// With one cluster the assignment cannot change, so every iteration of this k-means run has the same
		// column sums as the first one (the reference recomputes them; recalculate_means then finds moved == 0
		// and stops, src/cluster.c:231-233).  The first iteration takes them from the count table (see above);
		// later iterations copy the sums.
		if (h->k1_valid) {
			QVZ_CUDA(h, cudaMemcpyAsync(sums_dev, h->k1_sums, sum_bytes, cudaMemcpyDeviceToDevice, h->stream));
			return QVZ_OK;
		}
		qvz_kmeans_k1_ids_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(h->L.P, h->cl);
		QVZ_LAUNCHED(h);
		QVZ_CUDA(h, cudaGetLastError());
		int rc = qvz_cond_counts_launch(h, h->counts_dev);           // abi.cu sized counts_dev in kmeans_begin
		if (rc) return rc;
		h->counts_cached = 1;
		qvz_kmeans_k1_sums_kernel<<<h->L.C, 256, 0, h->stream>>>(h->L.C, h->L.n_lines, h->counts_dev, (unsigned long long *) sums_dev);
		QVZ_LAUNCHED(h);
		QVZ_CUDA(h, cudaGetLastError());
		QVZ_CUDA(h, cudaMemcpyAsync(h->k1_sums, sums_dev, sum_bytes, cudaMemcpyDeviceToDevice, h->stream));
		h->k1_valid = 1;
		return QVZ_OK;
	}
	// K >= 2: the first iteration of a run counts everything; later ones only move the rows that changed cluster.
	// k1_sums doubles as the run's local running sums (the caller may all-reduce sums_dev in place).
	const bool incr = h->k1_valid && !getenv("QVZ_KM_FULL");
	if (K <= 8 && h->means_t && mma_smem(K, C4, 128, 64) <= 220 * 1024 && !getenv("QVZ_KM_SORTED")) {      // register-blocked, tensor-core sums (see above)
		{
			int64_t *target = incr ? h->k1_sums : sums_dev;                 // later iterations add signed deltas to the running sums
			if (!incr) QVZ_CUDA(h, cudaMemsetAsync(sums_dev, 0, sum_bytes, h->stream));
			switch (K) {
			case 2: launch_mma<2>(h, target, !incr); break;
			case 3: launch_mma<3>(h, target, !incr); break;
			case 4: launch_mma<4>(h, target, !incr); break;
			case 5: launch_mma<5>(h, target, !incr); break;
			case 6: launch_mma<6>(h, target, !incr); break;
			case 7: launch_mma<7>(h, target, !incr); break;
			default: launch_mma<8>(h, target, !incr); break;
			}
			QVZ_LAUNCHED(h);
			QVZ_CUDA(h, cudaGetLastError());
			if (incr) QVZ_CUDA(h, cudaMemcpyAsync(sums_dev, h->k1_sums, sum_bytes, cudaMemcpyDeviceToDevice, h->stream));
			else QVZ_CUDA(h, cudaMemcpyAsync(h->k1_sums, sums_dev, sum_bytes, cudaMemcpyDeviceToDevice, h->stream));
			h->k1_valid = 1;
			return QVZ_OK;
		}
	}
	unsigned R = 128;                                                  // measured: 128-row tiles (3-5 CTAs per SM) beat 256 and 64
	while (R > 64 && assign_smem(K, C4, R) > 110 * 1024) R >>= 1;       // keep >= 2 CTAs per SM when possible
	if (const char *e = getenv("QVZ_KM_R")) R = (unsigned) atoi(e);     // tuning knob: 64, 128 or 256
	const size_t smem = assign_smem(K, C4, R, incr);
	if (smem > 220 * 1024) QVZ_FAIL(h, QVZ_ERR_UNSUPPORTED, "k-means: K*columns too large for shared memory");
	const uint64_t blocks = h->L.P / R;
	const uint64_t per_sm = (220 * 1024) / smem < (2048 / R) ? (220 * 1024) / smem : (2048 / R);
	const uint64_t cap = (uint64_t) h->sm_count * (per_sm ? per_sm : 1);
	const unsigned grid = (unsigned) (blocks < cap ? blocks : cap);
	int64_t *target = incr ? h->k1_sums : sums_dev;                     // INCR adds signed deltas to the running sums
	if (!incr) QVZ_CUDA(h, cudaMemsetAsync(sums_dev, 0, sum_bytes, h->stream));
	// the packed centroids of this iteration -> this handle's slot of the constant bank (device to device, stream ordered)
	const uint32_t KP = qvz_kmeans_kp(K);
	// (opt-in: measured SLOWER than the shared-memory broadcast on B200 -- 2.80 vs 2.91 TB/s per iteration on cfg4 --
	// the LDC operand fetch costs more issue slots than the two LDS.128 it replaces; kept for the record, QVZ_KM_CONST_MEANS=1)
	const bool cm = getenv("QVZ_KM_CONST_MEANS") && h->cm_slot >= 0 && h->means_t && (size_t) C4 * KP <= KM_CM_WORDS;
	if (cm)
		QVZ_CUDA(h, cudaMemcpyToSymbolAsync(c_means, h->means_t, (size_t) C4 * KP * sizeof(uint32_t),
		                                    (size_t) h->cm_slot * KM_CM_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->stream));
	switch (K) {
	case 2: launch_assign<2>(h, target, grid, R, smem, incr, cm); break;
	case 3: launch_assign<3>(h, target, grid, R, smem, incr, cm); break;
	case 4: launch_assign<4>(h, target, grid, R, smem, incr, cm); break;
	case 5: launch_assign<5>(h, target, grid, R, smem, incr, cm); break;
	case 6: launch_assign<6>(h, target, grid, R, smem, incr, cm); break;
	case 7: launch_assign<7>(h, target, grid, R, smem, incr, cm); break;
	case 8: launch_assign<8>(h, target, grid, R, smem, incr, cm); break;
	default: launch_assign<0>(h, target, grid, R, smem, incr, cm); break;
	}
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	if (incr) QVZ_CUDA(h, cudaMemcpyAsync(sums_dev, h->k1_sums, sum_bytes, cudaMemcpyDeviceToDevice, h->stream));
	else QVZ_CUDA(h, cudaMemcpyAsync(h->k1_sums, sums_dev, sum_bytes, cudaMemcpyDeviceToDevice, h->stream));
	h->k1_valid = 1;
	return QVZ_OK;
}

int qvz_kmeans_launch_update(qvz_gpu *h, const int64_t *sums_dev, double threshold, uint32_t max_iter) {
	qvz_kmeans_update_kernel<<<1, QVZ_THREADS, 0, h->stream>>>(
	    h->km_K, h->L.C, h->L.C4, (const unsigned long long *) sums_dev, h->means_b, h->means_w,
	    h->means_sq, h->moved, h->flags, sums_dev == nullptr, h->km_ctl, threshold, max_iter, h->moved_log,
	    (unsigned long long *) h->last_counts, h->km_K <= QVZ_MAX_K ? h->means_t : nullptr, qvz_kmeans_kp(h->km_K));
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}
