// well.cu -- WELL1024a jump-ahead so that every run of lines starts at its reference draw index.
//
// Reference: well_1024a (src/well.c:8-24) is F2-linear in its 32x32-bit state.  In the rotated frame
// u[i] = s[(n+i)&31] one step is the constant map  u' = [newV0, newV1, u[1], ..., u[30]]  and the
// returned word is u'[0]; the seed written by initialize_arithStream (src/qv_stream.c:76-93, n = 0) is u.
// The reference consumes exactly one 7-bit draw per (line, column) in line-major order
// (src/codebook.c:162-171 via src/well.c:33-46: 4 draws per word, top 4 bits dropped), so a run that
// starts at line L0 (L0 % 4 == 0) starts at word L0*C/4 of the stream.
//
// Matrices over F2 are kept in COLUMN form: Mc[j] = M * e_j (1024 bits = 32 words), bit j of a vector
// = (u[j>>5] >> (j&31)) & 1.  M*v = XOR of the columns selected by v; with a byte table
// tab[pos][b] = XOR_{t in b} Mc[8*pos+t] that is 128 coalesced 128-byte row reads per product
// ("four Russians").  A matrix product is the same kernel applied to the 1024 columns of the right factor.
//
//   P_j   = A^(2^j)                     cached per handle (squarings)
//   M_0   = A^(words per run),  M_b = M_{b-1}^2          cached per (words per run)
//   state[0] = A^(first word) * seed;   state[2^b + i] = M_b * state[i]      (doubling, one launch per level)
#include <stdlib.h>

#include <vector>

#include "qvz_internal.cuh"

#define WELL_BITS 1024
#define WELL_WORDS 32
#define WELL_TAB_WORDS (128u * 256u * 32u)

struct qvz_well_cache {
	std::vector<uint32_t *> pow2;     // column forms of A^(2^j)
	std::vector<uint32_t *> pow2_tabs; // their byte tables, built on first use by a vector jump (nullptr = not yet)
	uint32_t *tab;                    // scratch byte table, 4 MiB
	uint32_t *tmp_a, *tmp_b;          // scratch column forms
	uint32_t *vec_a, *vec_b;          // scratch single vectors
	uint64_t lw;                      // words per run the levels below were built for
	std::vector<uint32_t *> levels;   // column forms of M_b
	std::vector<uint32_t *> level_tabs; // byte tables of M_b (4 MiB each), kept so that a quantize call only applies them
};

// host: one step in the rotated frame (restates src/well.c:8-24 with n folded away)
static void well_step_frame(uint32_t u[WELL_WORDS]) {
	const uint32_t oldest = u[31], a = u[3], b = u[24], c = u[10];
	const uint32_t z1 = u[0] ^ (a ^ (a >> 8));
	const uint32_t z2 = (b ^ (b << 19)) ^ (c ^ (c << 14));
	const uint32_t v0 = (oldest ^ (oldest << 11)) ^ (z1 ^ (z1 << 7)) ^ (z2 ^ (z2 << 13));
	for (int i = 31; i >= 2; --i) u[i] = u[i - 1];
	u[1] = z1 ^ z2;
	u[0] = v0;
}

// tab[pos][b][lane] for one column-form matrix: one 32-thread CTA per byte position
__global__ void __launch_bounds__(32)
qvz_f2_build_table_kernel(const uint32_t *__restrict__ Mc, uint32_t *__restrict__ tab)
{
	__shared__ uint32_t t[256 * 32];
	const uint32_t pos = blockIdx.x, lane = threadIdx.x;
	t[lane] = 0;
	for (uint32_t b = 1; b < 256; ++b) {
		const uint32_t low = __ffs(b) - 1;
		t[b * 32 + lane] = t[(b & (b - 1)) * 32 + lane] ^ Mc[(8 * pos + low) * 32 + lane];
	}
	for (uint32_t b = 0; b < 256; ++b) tab[(pos * 256 + b) * 32 + lane] = t[b * 32 + lane];
}

// out[v] = M * in[v] for `count` vectors; one warp per vector, lane = output word
__global__ void __launch_bounds__(QVZ_THREADS)
qvz_f2_apply_kernel(const uint32_t *__restrict__ tab, const uint32_t *__restrict__ in,
                    uint32_t *__restrict__ out, uint32_t count)
{
	const uint32_t v = (blockIdx.x * QVZ_THREADS + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (v >= count) return;
	const uint32_t mine = in[(uint64_t) v * 32 + lane];
	uint32_t acc = 0;
#pragma unroll 32
	for (uint32_t pos = 0; pos < 128; ++pos) {          // the 128 row loads are independent: keep 32 in flight
		const uint32_t w = __shfl_sync(0xFFFFFFFFu, mine, pos >> 2);
		const uint32_t b = (w >> (8 * (pos & 3))) & 0xFFu;
		acc ^= __ldg(&tab[(pos * 256 + b) * 32 + lane]);
	}
	out[(uint64_t) v * 32 + lane] = acc;
}

static int f2_table(qvz_gpu *h, const uint32_t *Mc, uint32_t *tab = nullptr) {
	qvz_f2_build_table_kernel<<<128, 32, 0, h->stream>>>(Mc, tab ? tab : h->well->tab);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}

static int f2_apply(qvz_gpu *h, const uint32_t *in, uint32_t *out, uint32_t count, const uint32_t *tab = nullptr) {
	if (!count) return QVZ_OK;
	const unsigned grid = (count * 32 + QVZ_THREADS - 1) / QVZ_THREADS;
	qvz_f2_apply_kernel<<<grid, QVZ_THREADS, 0, h->stream>>>(tab ? tab : h->well->tab, in, out, count);
	QVZ_LAUNCHED(h);
	QVZ_CUDA(h, cudaGetLastError());
	return QVZ_OK;
}

int qvz_well_init(qvz_gpu *h) {
	qvz_well_cache *w = new qvz_well_cache();
	h->well = w;
	w->lw = 0;
	const size_t msz = (size_t) WELL_BITS * WELL_WORDS * sizeof(uint32_t);
	QVZ_CUDA(h, cudaMalloc(&w->tab, (size_t) WELL_TAB_WORDS * sizeof(uint32_t)));
	QVZ_CUDA(h, cudaMalloc(&w->tmp_a, msz));
	QVZ_CUDA(h, cudaMalloc(&w->tmp_b, msz));
	QVZ_CUDA(h, cudaMalloc(&w->vec_a, WELL_WORDS * sizeof(uint32_t)));
	QVZ_CUDA(h, cudaMalloc(&w->vec_b, WELL_WORDS * sizeof(uint32_t)));

	// column j of A = one step applied to the unit vector e_j
	std::vector<uint32_t> A((size_t) WELL_BITS * WELL_WORDS);
	for (int j = 0; j < WELL_BITS; ++j) {
		uint32_t u[WELL_WORDS] = {0};
		u[j >> 5] = 1u << (j & 31);
		well_step_frame(u);
		for (int k = 0; k < WELL_WORDS; ++k) A[(size_t) j * WELL_WORDS + k] = u[k];
	}
	uint32_t *p0 = nullptr;
	QVZ_CUDA(h, cudaMalloc(&p0, msz));
	QVZ_CUDA(h, cudaMemcpyAsync(p0, A.data(), msz, cudaMemcpyHostToDevice, h->stream));
	QVZ_CUDA(h, cudaStreamSynchronize(h->stream));
	w->pow2.push_back(p0);
	return QVZ_OK;
}

void qvz_well_free(qvz_gpu *h) {
	qvz_well_cache *w = h->well;
	if (!w) return;
	for (uint32_t *p : w->pow2) cudaFree(p);
	for (uint32_t *p : w->pow2_tabs) cudaFree(p);
	for (uint32_t *p : w->levels) cudaFree(p);
	for (uint32_t *p : w->level_tabs) cudaFree(p);
	cudaFree(w->tab);
	cudaFree(w->tmp_a);
	cudaFree(w->tmp_b);
	cudaFree(w->vec_a);
	cudaFree(w->vec_b);
	delete w;
	h->well = nullptr;
}

// make sure A^(2^j) exists
static int ensure_pow2(qvz_gpu *h, uint32_t j) {
	qvz_well_cache *w = h->well;
	const size_t msz = (size_t) WELL_BITS * WELL_WORDS * sizeof(uint32_t);
	while (w->pow2.size() <= j) {
		uint32_t *next = nullptr;
		QVZ_CUDA(h, cudaMalloc(&next, msz));
		const uint32_t *prev = w->pow2.back();
		int rc = f2_table(h, prev);
		if (rc) return rc;
		rc = f2_apply(h, prev, next, WELL_BITS);      // P_{j+1} = P_j * P_j, column by column
		if (rc) return rc;
		w->pow2.push_back(next);
	}
	return QVZ_OK;
}

// vec (device, 32 words) <- A^e * vec, using vec_a/vec_b as ping-pong
static int jump_vector(qvz_gpu *h, uint64_t e, uint32_t *vec /* = w->vec_a */) {
	qvz_well_cache *w = h->well;
	uint32_t *cur = vec, *other = (vec == w->vec_a) ? w->vec_b : w->vec_a;
	for (uint32_t j = 0; e >> j; ++j) {
		if (!((e >> j) & 1)) continue;
		int rc = ensure_pow2(h, j);
		if (rc) return rc;
		if (w->pow2_tabs.size() <= j) w->pow2_tabs.resize(j + 1, nullptr);
		if (!w->pow2_tabs[j]) {
			QVZ_CUDA(h, cudaMalloc(&w->pow2_tabs[j], (size_t) WELL_TAB_WORDS * sizeof(uint32_t)));
			rc = f2_table(h, w->pow2[j], w->pow2_tabs[j]);
			if (rc) return rc;
		}
		rc = f2_apply(h, cur, other, 1, w->pow2_tabs[j]);
		if (rc) return rc;
		uint32_t *t = cur;
		cur = other;
		other = t;
	}
	if (cur != vec)
		QVZ_CUDA(h, cudaMemcpyAsync(vec, cur, WELL_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->stream));
	return QVZ_OK;
}

int qvz_well_jump_state(qvz_gpu *h, const uint32_t seed[32], uint64_t words, uint32_t *state_dev) {
	qvz_well_cache *w = h->well;
	QVZ_CUDA(h, cudaMemcpyAsync(w->vec_a, seed, WELL_WORDS * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
	int rc = jump_vector(h, words, w->vec_a);
	if (rc) return rc;
	QVZ_CUDA(h, cudaMemcpyAsync(state_dev, w->vec_a, WELL_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->stream));
	return QVZ_OK;
}

// (re)build M_b = A^(lw * 2^b) for b < nlevels
static int ensure_levels(qvz_gpu *h, uint64_t lw, uint32_t nlevels) {
	qvz_well_cache *w = h->well;
	const size_t msz = (size_t) WELL_BITS * WELL_WORDS * sizeof(uint32_t);
	if (w->lw != lw) {
		for (uint32_t *p : w->levels) cudaFree(p);
		for (uint32_t *p : w->level_tabs) cudaFree(p);
		w->levels.clear();
		w->level_tabs.clear();
		w->lw = lw;
	}
	if (w->levels.empty() && nlevels) {
		// M_0 = product of P_j over the set bits of lw; start from the identity in column form
		std::vector<uint32_t> I((size_t) WELL_BITS * WELL_WORDS, 0);
		for (int j = 0; j < WELL_BITS; ++j) I[(size_t) j * WELL_WORDS + (j >> 5)] = 1u << (j & 31);
		uint32_t *cur = w->tmp_a, *other = w->tmp_b;
		QVZ_CUDA(h, cudaMemcpyAsync(cur, I.data(), msz, cudaMemcpyHostToDevice, h->stream));
		QVZ_CUDA(h, cudaStreamSynchronize(h->stream));   // I is a stack-lifetime host buffer
		for (uint32_t j = 0; lw >> j; ++j) {
			if (!((lw >> j) & 1)) continue;
			int rc = ensure_pow2(h, j);
			if (rc) return rc;
			rc = f2_table(h, w->pow2[j]);
			if (rc) return rc;
			rc = f2_apply(h, cur, other, WELL_BITS);
			if (rc) return rc;
			uint32_t *t = cur;
			cur = other;
			other = t;
		}
		uint32_t *m0 = nullptr;
		QVZ_CUDA(h, cudaMalloc(&m0, msz));
		QVZ_CUDA(h, cudaMemcpyAsync(m0, cur, msz, cudaMemcpyDeviceToDevice, h->stream));
		w->levels.push_back(m0);
	}
	while (w->levels.size() < nlevels) {
		uint32_t *next = nullptr;
		QVZ_CUDA(h, cudaMalloc(&next, msz));
		const uint32_t *prev = w->levels.back();
		int rc = f2_table(h, prev);
		if (rc) return rc;
		rc = f2_apply(h, prev, next, WELL_BITS);
		if (rc) return rc;
		w->levels.push_back(next);
	}
	while (w->level_tabs.size() < nlevels) {
		uint32_t *tab = nullptr;
		QVZ_CUDA(h, cudaMalloc(&tab, (size_t) WELL_TAB_WORDS * sizeof(uint32_t)));
		w->level_tabs.push_back(tab);
		int rc = f2_table(h, w->levels[w->level_tabs.size() - 1], tab);
		if (rc) return rc;
	}
	return QVZ_OK;
}

// debugging aid (QVZ_DEBUG_WELL=1): xor checksums of the cached level tables and of the run states
void qvz_well_debug(qvz_gpu *h, const char *where) {
	if (!getenv("QVZ_DEBUG_WELL")) return;
	cudaStreamSynchronize(h->stream);
	qvz_well_cache *w = h->well;
	std::vector<uint32_t> buf(WELL_TAB_WORDS);
	fprintf(stderr, "[well %s] lw=%llu tabs:", where, (unsigned long long) w->lw);
	for (size_t b = 0; b < w->level_tabs.size(); ++b) {
		cudaMemcpy(buf.data(), w->level_tabs[b], buf.size() * 4, cudaMemcpyDeviceToHost);
		uint32_t x = 0;
		for (uint32_t v : buf) x = (x * 31u) ^ v;
		fprintf(stderr, " %08x", x);
	}
	std::vector<uint32_t> rs((size_t) h->L.T * 32);
	cudaMemcpy(rs.data(), h->run_states, rs.size() * 4, cudaMemcpyDeviceToHost);
	uint32_t lo = 0, hi = 0;
	for (size_t i = 0; i < rs.size() / 2; ++i) lo = (lo * 31u) ^ rs[i];
	for (size_t i = rs.size() / 2; i < rs.size(); ++i) hi = (hi * 31u) ^ rs[i];
	fprintf(stderr, " | states lo %08x hi %08x (ptr %p, tab7 %p)\n", lo, hi, (void *) h->run_states, w->level_tabs.empty() ? nullptr : (void *) w->level_tabs.back());
}

int qvz_well_run_states(qvz_gpu *h, const uint32_t seed[32]) {
	const qvz_layout &L = h->L;
	if ((L.first_line * (uint64_t) L.C) & 3)
		QVZ_FAIL(h, QVZ_ERR_ARG, "first_line*columns must be a multiple of 4 (WELL word boundary)");
	const uint64_t w0 = L.first_line * (uint64_t) L.C / 4;
	const uint64_t lw = (uint64_t) L.Lr * L.C / 4;
	uint32_t nlevels = 0;
	while ((1ull << nlevels) < L.T) ++nlevels;
	int rc = ensure_levels(h, lw, nlevels);
	if (rc) return rc;
	rc = qvz_well_jump_state(h, seed, w0, h->run_states);
	if (rc) return rc;
	for (uint32_t b = 0; b < nlevels; ++b) {
		const uint64_t have = 1ull << b;
		const uint64_t todo = (L.T - have < have) ? L.T - have : have;
		rc = f2_apply(h, h->run_states, h->run_states + have * WELL_WORDS, (uint32_t) todo, h->well->level_tabs[b]);
		if (rc) return rc;
	}
	return QVZ_OK;
}
