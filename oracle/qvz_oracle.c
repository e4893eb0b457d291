/*
 * qvz_oracle.c -- TEST INFRASTRUCTURE ONLY (see qvz_oracle.h).
 *
 * Plain-C restatement of the reference algorithm for the hot path.  Each function names the
 * reference lines it follows (paths relative to the reference tree).  It keeps the reference's
 * arithmetic types on purpose (uint32 squares summed in a double, uint64 accumulators, double
 * error sums in column order) so that it is a restatement and not a re-derivation.
 */
#include "qvz_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ WELL1024a (src/well.c) */

void oracle_well_seed(struct oracle_well *w, const uint32_t seed[32]) {
	/* initialize_arithStream: memset 0, copy the 32 words, n = 0 (src/qv_stream.c:70-93) */
	memset(w, 0, sizeof(*w));
	memcpy(w->state, seed, 32 * sizeof(uint32_t));
}

/* src/well.c:8-24 -- M1=3, M2=24, M3=10; index steps backwards through the ring */
uint32_t oracle_well_next(struct oracle_well *w) {
	uint32_t *s = w->state;
	uint32_t n = w->n;
	uint32_t oldest = s[(n + 31) & 31];
	uint32_t a = s[(n + 3) & 31];
	uint32_t b = s[(n + 24) & 31];
	uint32_t c = s[(n + 10) & 31];
	uint32_t z1 = s[n] ^ (a ^ (a >> 8));
	uint32_t z2 = (b ^ (b << 19)) ^ (c ^ (c << 14));
	uint32_t out;
	s[n] = z1 ^ z2;
	n = (n + 31) & 31;
	out = (oldest ^ (oldest << 11)) ^ (z1 ^ (z1 << 7)) ^ (z2 ^ (z2 << 13));
	s[n] = out;
	w->n = n;
	return out;
}

/* src/well.c:33-46 with bits == 7: refill when fewer than 7 bits remain, serve low bits first */
uint32_t oracle_well_bits7(struct oracle_well *w) {
	uint32_t r;
	if (w->bits_left < 7) {
		w->bit_output = oracle_well_next(w);
		w->bits_left = 32;
	}
	r = w->bit_output & 127u;
	w->bit_output >>= 7;
	w->bits_left -= 7;
	return r;
}

void oracle_well_words(const uint32_t seed[32], uint64_t skip, uint64_t count, uint32_t *out) {
	struct oracle_well w;
	uint64_t i;
	oracle_well_seed(&w, seed);
	for (i = 0; i < skip; ++i) (void) oracle_well_next(&w);
	for (i = 0; i < count; ++i) out[i] = oracle_well_next(&w);
}

void oracle_well_draws(const uint32_t seed[32], uint64_t count, uint8_t *out) {
	struct oracle_well w;
	uint64_t i;
	oracle_well_seed(&w, seed);
	for (i = 0; i < count; ++i) out[i] = (uint8_t) oracle_well_bits7(&w);
}

void oracle_well_state_after(const uint32_t seed[32], uint64_t words, uint32_t state_out[32]) {
	struct oracle_well w;
	uint64_t i;
	uint32_t k;
	oracle_well_seed(&w, seed);
	for (i = 0; i < words; ++i) (void) oracle_well_next(&w);
	for (k = 0; k < 32; ++k) state_out[k] = w.state[(w.n + k) & 31];
}

/* ------------------------------------------------------------------ k-means (src/cluster.c) */

int32_t oracle_kmeans(const uint8_t *rows, uint64_t n_lines, uint32_t columns, uint32_t row_stride,
                      uint32_t K, const uint8_t *init_means, double threshold, uint32_t max_iter,
                      uint8_t *cluster_ids, uint8_t *means_out, uint32_t *counts_out,
                      double *moved_log)
{
	uint8_t *mean = (uint8_t *) malloc((size_t) K * columns);
	uint64_t *acc = (uint64_t *) malloc((size_t) K * columns * sizeof(uint64_t));
	uint32_t *count = (uint32_t *) malloc((size_t) K * sizeof(uint32_t));
	double *dist = (double *) malloc((size_t) K * sizeof(double));
	uint32_t iter = 0, k, i;
	uint64_t n;
	int loop = 1;
	int32_t rc;

	/* initialize_kmeans_clustering copies the picked rows (src/cluster.c:201) */
	memcpy(mean, init_means, (size_t) K * columns);

	/* do_kmeans_clustering (src/cluster.c:221-239) */
	while (iter < max_iter && loop) {
		double move_max = 0.0;
		for (k = 0; k < K; ++k) count[k] = 0;

		for (n = 0; n < n_lines; ++n) {
			const uint8_t *x = rows + n * (uint64_t) row_stride;
			uint32_t best = 0;
			double d;
			/* find_distance (src/cluster.c:176-187): uint32 difference squared, double sum */
			for (k = 0; k < K; ++k) {
				double s = 0.0;
				for (i = 0; i < columns; ++i) {
					uint32_t data = x[i], m = mean[(size_t) k * columns + i];
					s += (data - m) * (data - m);
				}
				dist[k] = s;
			}
			/* assign_cluster (src/cluster.c:149-171): strict '<', lowest id wins ties */
			d = dist[0];
			for (k = 1; k < K; ++k) {
				if (dist[k] < d) {
					best = k;
					d = dist[k];
				}
			}
			cluster_ids[n] = (uint8_t) best;
			count[best] += 1;
		}

		/* recalculate_means (src/cluster.c:80-131) */
		memset(acc, 0, (size_t) K * columns * sizeof(uint64_t));
		for (n = 0; n < n_lines; ++n) {
			const uint8_t *x = rows + n * (uint64_t) row_stride;
			uint64_t *a = acc + (size_t) cluster_ids[n] * columns;
			for (i = 0; i < columns; ++i) a[i] += x[i];
		}
		for (k = 0; k < K; ++k) {
			double moved = 0.0;
			if (count[k] == 0) {
				rc = -1;            /* integer division by zero in the reference (:113) */
				goto done;
			}
			for (i = 0; i < columns; ++i) {
				uint8_t nm = (uint8_t) (acc[(size_t) k * columns + i] / count[k]);
				double dd = nm - mean[(size_t) k * columns + i];
				moved += dd * dd;
				mean[(size_t) k * columns + i] = nm;
			}
			if (moved > move_max) move_max = moved;
			if (moved_log) moved_log[(size_t) iter * K + k] = moved;
		}
		loop = move_max > threshold;
		iter += 1;
	}
	rc = (int32_t) iter;
done:
	if (means_out) memcpy(means_out, mean, (size_t) K * columns);
	if (counts_out) memcpy(counts_out, count, (size_t) K * sizeof(uint32_t));
	free(mean);
	free(acc);
	free(count);
	free(dist);
	return rc;
}

/* ------------------------------------- conditional counts (src/codebook.c:193-205, :116-120) */

void oracle_cond_counts(const uint8_t *rows, uint64_t n_lines, uint32_t columns, uint32_t row_stride,
                        uint32_t K, const uint8_t *cluster_ids, uint32_t *counts)
{
	uint64_t per_cluster = (uint64_t) (1 + QVZ_ALPHABET * (columns - 1)) * QVZ_ALPHABET;
	uint64_t n;
	uint32_t c;
	memset(counts, 0, (size_t) (per_cluster * K) * sizeof(uint32_t));
	for (n = 0; n < n_lines; ++n) {
		const uint8_t *x = rows + n * (uint64_t) row_stride;
		uint32_t *t = counts + per_cluster * cluster_ids[n];
		/* get_cond_pmf(list, 0, 0) -> pmfs[0]; pmf_increment(pmf, x0 - 33) */
		t[(uint32_t) (x[0] - 33)] += 1;
		/* get_cond_pmf(list, c, prev) -> pmfs[1 + (c-1)*72 + prev] with prev = RAW x[c-1]-33 */
		for (c = 1; c < columns; ++c) {
			uint32_t row = 1 + (c - 1) * QVZ_ALPHABET + (uint32_t) (x[c - 1] - 33);
			t[(uint64_t) row * QVZ_ALPHABET + (uint32_t) (x[c] - 33)] += 1;
		}
	}
}

/* ------------------ quantize walk (src/qv_compressor.c:76-135, src/codebook.c:162-171) */

double oracle_quantize(const uint8_t *rows, uint64_t n_lines, uint32_t columns, uint32_t row_stride,
                       uint64_t first_line, const uint8_t *cluster_ids,
                       const struct qvz_flat_tables *t, const uint32_t seed[32],
                       uint8_t *symbols, uint8_t *qv_image, double *line_err)
{
	struct oracle_well w;
	double distortion = 0.0;
	uint64_t n, skip;
	uint32_t s;

	/* one draw per (line, column) in line-major order => this shard starts at draw first_line*C;
	 * replay the bit server from the seed up to that draw. */
	oracle_well_seed(&w, seed);
	skip = first_line * (uint64_t) columns;
	for (n = 0; n < skip; ++n) (void) oracle_well_bits7(&w);

	for (n = 0; n < n_lines; ++n) {
		const uint8_t *x = rows + n * (uint64_t) row_stride;
		uint32_t k = cluster_ids[n];
		uint8_t prev = 0;
		double error = 0.0;
		for (s = 0; s < columns; ++s) {
			size_t kc = (size_t) k * columns + s;
			/* choose_quantizer: idx = indexes[prev]; draw >= qratio picks the hi quantizer */
			uint32_t ctx = t->ctx_of[kc * QVZ_ALPHABET + prev];
			uint32_t draw, hi;
			uint64_t q;
			uint8_t data, qv, st;
			if (ctx == QVZ_CTX_ABSENT) return -1.0;        /* assert at src/codebook.c:164 */
			draw = oracle_well_bits7(&w);
			hi = draw >= t->qratio[t->q_off[kc] / 2 + ctx];
			q = t->q_off[kc] + 2 * ctx + hi;
			data = (uint8_t) (x[s] - 33);
			qv = t->qmap[q * QVZ_ALPHABET + data];
			st = t->smap[q * QVZ_ALPHABET + qv];
			symbols[n * columns + s] = (uint8_t) (st | (hi << 7));
			if (qv_image) qv_image[n * (uint64_t) (columns + 1) + s] = (uint8_t) (qv + 33);
			/* error = d(x, qv) at s == 0, += afterwards (src/qv_compressor.c:97, :118) */
			if (s == 0) error = t->distortion[data + QVZ_ALPHABET * qv];
			else error += t->distortion[data + QVZ_ALPHABET * qv];
			prev = qv;
		}
		if (qv_image) qv_image[n * (uint64_t) (columns + 1) + columns] = '\n';
		if (line_err) line_err[n] = error / ((double) columns);
		distortion += error / ((double) columns);
	}
	return distortion / ((double) n_lines);
}
