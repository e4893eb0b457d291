/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin ctypes-friendly shim around the UNMODIFIED reference sources of
 * mikelhernaez/qvz, which oracle/Makefile compiles from where they lie under
 * /root/reference into oracle/_ref/libqvzref.so (flags: -O3 -DLINUX -DDEBUG, the
 * reference's own `make debug` seed semantics, src/Makefile:21, qv_stream.c:79-83).
 *
 * Every function here only builds the reference's own structures and calls the
 * reference's own functions:
 *   do_kmeans_clustering / cluster_lines / recalculate_means   (src/cluster.c:65-244)
 *   calculate_statistics                                        (src/codebook.c:185-220)
 *   generate_codebooks                                          (src/codebook.c:355-468)
 *   choose_quantizer / well_1024a_bits                          (src/codebook.c:162-171, src/well.c:33-46)
 *   encode()                                                    (src/main.c:18-127, via -Dmain=qvz_ref_main)
 * and copies results out into flat arrays so that tests can compare them with the
 * CUDA path and with oracle/qvz_oracle.c.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "codebook.h"
#include "cluster.h"
#include "qv_compressor.h"

/* defined in the reference's main.c (non-static) */
void encode(char *input_name, char *output_name, struct qv_options_t *opts);
void decode(char *input_file, char *output_file, struct qv_options_t *opts);

#define REF_ALPHABET 72

struct ref_ctx {
	struct quality_file_t info;
	struct qv_options_t opts;
	const uint8_t *rows;
	int have_stats;
	int have_codebooks;
};

/* Build quality_file_t over a caller-owned row buffer exactly as load_file does over its
 * mmap (src/lines.c:54-79): m_data = base + line*(columns+1), blocks of <= 1e6 lines. */
struct ref_ctx *ref_create(const uint8_t *rows, uint64_t lines, uint32_t columns,
                           uint32_t clusters, double threshold,
                           int mode, double ratio, int distortion)
{
	struct ref_ctx *h = (struct ref_ctx *) calloc(1, sizeof(*h));
	uint64_t n;
	uint32_t b, l;

	h->rows = rows;
	h->opts.verbose = 0;
	h->opts.stats = 0;
	h->opts.mode = (uint8_t) mode;
	h->opts.clusters = (uint8_t) clusters;
	h->opts.uncompressed = 0;
	h->opts.distortion = (uint8_t) distortion;
	h->opts.ratio = ratio;
	h->opts.cluster_threshold = threshold;

	h->info.alphabet = alloc_alphabet(REF_ALPHABET);
	h->info.dist = generate_distortion_matrix(REF_ALPHABET, distortion);
	h->info.cluster_count = (uint8_t) clusters;
	h->info.lines = lines;
	h->info.columns = columns;
	if (alloc_blocks(&h->info) != LF_ERROR_NONE) {
		free(h);
		return NULL;
	}
	n = 0;
	for (b = 0; b < h->info.block_count; ++b) {
		for (l = 0; l < h->info.blocks[b].count; ++l) {
			h->info.blocks[b].lines[l].m_data = rows + n * (uint64_t)(columns + 1);
			n += 1;
		}
	}
	h->info.clusters = alloc_cluster_list(&h->info);
	h->info.opts = &h->opts;
	return h;
}

static struct line_t *ref_line(struct ref_ctx *h, uint64_t n) {
	return &h->info.blocks[n / MAX_LINES_PER_BLOCK].lines[n % MAX_LINES_PER_BLOCK];
}

/* Unmodified reference driver: srand(1) restores glibc's start-up rand() stream so the picks
 * equal those of a fresh `qvz` process (cluster.c:199-200 uses unseeded rand()). */
void ref_kmeans_reference_driver(struct ref_ctx *h, uint8_t *ids_out) {
	uint64_t n;
	srand(1);
	do_kmeans_clustering(&h->info);
	for (n = 0; n < h->info.lines; ++n)
		ids_out[n] = ref_line(h, n)->cluster;
}

/* Same loop as do_kmeans_clustering (cluster.c:221-239) but with caller-chosen initial rows and
 * per-iteration logging, driving the reference's own cluster_lines / recalculate_means. */
uint32_t ref_kmeans(struct ref_ctx *h, const uint64_t *init_lines, uint32_t max_iter,
                    uint8_t *ids_out, uint8_t *means_out, uint32_t *counts_out,
                    double *moved_log /* max_iter*K or NULL */)
{
	struct cluster_list_t *cl = h->info.clusters;
	uint32_t K = h->info.cluster_count, C = h->info.columns;
	uint32_t iter = 0, j, c;
	uint64_t n;
	int loop = 1;
	uint8_t *old = (uint8_t *) malloc((size_t) K * C);

	for (j = 0; j < K; ++j)
		memcpy(cl->clusters[j].mean, ref_line(h, init_lines[j])->m_data, C);

	while (iter < max_iter && loop) {
		double moved;
		for (j = 0; j < K; ++j) {
			cl->clusters[j].count = 0;
			memcpy(old + (size_t) j * C, cl->clusters[j].mean, C);
		}
		for (j = 0; j < h->info.block_count; ++j)
			cluster_lines(&h->info.blocks[j], &h->info);
		moved = recalculate_means(&h->info);
		if (moved_log) {
			for (j = 0; j < K; ++j) {
				double m = 0.0;
				for (c = 0; c < C; ++c) {
					double d = (double) cl->clusters[j].mean[c] - (double) old[(size_t) j * C + c];
					m += d * d;
				}
				moved_log[(size_t) iter * K + j] = m;
			}
		}
		loop = moved > h->opts.cluster_threshold;
		iter += 1;
	}
	for (n = 0; n < h->info.lines; ++n)
		ids_out[n] = ref_line(h, n)->cluster;
	for (j = 0; j < K; ++j) {
		if (means_out) memcpy(means_out + (size_t) j * C, cl->clusters[j].mean, C);
		if (counts_out) counts_out[j] = cl->clusters[j].count;
	}
	free(old);
	return iter;
}

/* Install cluster ids computed elsewhere (lets stats/quantize be tested in isolation). */
void ref_set_clusters(struct ref_ctx *h, const uint8_t *ids) {
	uint64_t n;
	for (n = 0; n < h->info.lines; ++n)
		ref_line(h, n)->cluster = ids[n];
}

/* calculate_statistics (codebook.c:185), then copy every pmf's counts in get_cond_pmf order
 * (codebook.c:116-120): out[k][p][72], p = 0 | 1 + (col-1)*72 + prev. */
void ref_stats(struct ref_ctx *h, uint32_t *counts_out, uint32_t *totals_out) {
	uint32_t K = h->info.cluster_count, C = h->info.columns;
	uint32_t rows = 1 + REF_ALPHABET * (C - 1);
	uint32_t k, p;
	if (!h->have_stats) {
		calculate_statistics(&h->info);
		h->have_stats = 1;
	}
	for (k = 0; k < K; ++k) {
		struct cond_pmf_list_t *pl = h->info.clusters->clusters[k].training_stats;
		for (p = 0; p < rows; ++p) {
			if (counts_out)
				memcpy(counts_out + ((size_t) k * rows + p) * REF_ALPHABET, pl->pmfs[p]->counts,
				       REF_ALPHABET * sizeof(uint32_t));
			if (totals_out)
				totals_out[(size_t) k * rows + p] = pl->pmfs[p]->total;
		}
	}
}

void ref_codebooks(struct ref_ctx *h) {
	if (!h->have_stats) {
		calculate_statistics(&h->info);
		h->have_stats = 1;
	}
	if (!h->have_codebooks) {
		generate_codebooks(&h->info);
		h->have_codebooks = 1;
	}
}

/* Number of quantizers (lo+hi over all contexts/columns/clusters) after ref_codebooks. */
uint64_t ref_tables_count(struct ref_ctx *h) {
	uint32_t K = h->info.cluster_count, C = h->info.columns, k, c;
	uint64_t nq = 0;
	for (k = 0; k < K; ++k) {
		struct cond_quantizer_list_t *ql = h->info.clusters->clusters[k].qlist;
		for (c = 0; c < C; ++c)
			nq += 2 * (uint64_t) ql->input_alphabets[c]->size;
	}
	return nq;
}

/* Flatten cond_quantizer_list_t (codebook.h:61-69) into the pointer-free layout of
 * include/qvz_gpu.h `struct qvz_flat_tables`:
 *   nctx[k*C+c], ctx_of[(k*C+c)*72 + v], q_off[k*C+c] (first quantizer of the column),
 *   qratio[q_off/2 + ctx], qmap[(q_off + 2*ctx+hi)*72 + x], smap[(...)*72 + qv]. */
void ref_tables_export(struct ref_ctx *h, uint32_t *nctx, uint8_t *ctx_of, uint64_t *q_off,
                       uint8_t *qratio, uint8_t *qmap, uint8_t *smap, double *distortion)
{
	uint32_t K = h->info.cluster_count, C = h->info.columns, k, c, v, j;
	uint64_t nq = 0;
	for (k = 0; k < K; ++k) {
		struct cond_quantizer_list_t *ql = h->info.clusters->clusters[k].qlist;
		for (c = 0; c < C; ++c) {
			const struct alphabet_t *A = ql->input_alphabets[c];
			size_t kc = (size_t) k * C + c;
			nctx[kc] = A->size;
			q_off[kc] = nq;
			for (v = 0; v < REF_ALPHABET; ++v) {
				uint32_t idx = A->indexes[v];
				ctx_of[kc * REF_ALPHABET + v] = (idx == ALPHABET_SYMBOL_NOT_FOUND) ? 0xFF : (uint8_t) idx;
			}
			for (j = 0; j < A->size; ++j)
				qratio[nq / 2 + j] = ql->qratio[c][j];
			for (j = 0; j < 2 * A->size; ++j) {
				const struct quantizer_t *q = ql->q[c][j];
				for (v = 0; v < REF_ALPHABET; ++v) {
					uint32_t idx = q->output_alphabet->indexes[v];
					qmap[(nq + j) * REF_ALPHABET + v] = q->q[v];
					smap[(nq + j) * REF_ALPHABET + v] = (idx == ALPHABET_SYMBOL_NOT_FOUND) ? 0xFF : (uint8_t) idx;
				}
			}
			nq += 2 * (uint64_t) A->size;
		}
	}
	if (distortion)
		memcpy(distortion, h->info.dist->distortion, REF_ALPHABET * REF_ALPHABET * sizeof(double));
}

/* The quantize walk of start_qv_compression (qv_compressor.c:76-135) with the arithmetic coder
 * calls left out: same order of choose_quantizer (one WELL draw per symbol), same error sums. */
double ref_quantize(struct ref_ctx *h, const uint32_t seed[32], uint8_t *symbols /* N*C */,
                    uint8_t *qv_image /* N*(C+1) or NULL */, double *line_err /* N or NULL */)
{
	uint32_t C = h->info.columns, s, idx;
	uint64_t n;
	double distortion = 0.0;

	memset(&h->info.well, 0, sizeof(struct well_state_t));
	memcpy(h->info.well.state, seed, 32 * sizeof(uint32_t));
	h->info.well.n = 0;

	for (n = 0; n < h->info.lines; ++n) {
		struct line_t *line = ref_line(h, n);
		struct cond_quantizer_list_t *ql = h->info.clusters->clusters[line->cluster].qlist;
		uint8_t prev = 0;
		double error = 0.0;
		for (s = 0; s < C; ++s) {
			struct quantizer_t *q = choose_quantizer(ql, &h->info.well, s, prev, &idx);
			uint8_t data = line->m_data[s] - 33;
			uint8_t qv = q->q[data];
			uint32_t st = get_symbol_index(q->output_alphabet, qv);
			symbols[n * C + s] = (uint8_t) (st | ((idx & 1) << 7));
			if (qv_image) qv_image[n * (uint64_t)(C + 1) + s] = qv + 33;
			if (s == 0) error = get_distortion(h->info.dist, data, qv);
			else error += get_distortion(h->info.dist, data, qv);
			prev = qv;
		}
		if (qv_image) qv_image[n * (uint64_t)(C + 1) + C] = '\n';
		if (line_err) line_err[n] = error / ((double) C);
		distortion += error / ((double) C);
	}
	return distortion / ((double) h->info.lines);
}

/* First `count` raw WELL1024a words / 7-bit draws from a seed (well.c:8-46). */
void ref_well_words(const uint32_t seed[32], uint64_t skip, uint64_t count, uint32_t *out) {
	struct well_state_t w;
	uint64_t i;
	memset(&w, 0, sizeof(w));
	memcpy(w.state, seed, sizeof(w.state));
	for (i = 0; i < skip; ++i) (void) well_1024a(&w);
	for (i = 0; i < count; ++i) out[i] = well_1024a(&w);
}

void ref_well_draws(const uint32_t seed[32], uint64_t count, uint8_t *out) {
	struct well_state_t w;
	uint64_t i;
	memset(&w, 0, sizeof(w));
	memcpy(w.state, seed, sizeof(w.state));
	for (i = 0; i < count; ++i) out[i] = (uint8_t) well_1024a_bits(&w, 7);
}

/* Whole unmodified encode()/decode() on files (DEBUG WELL seed 0x55555555 in this build). */
void ref_encode_file(const char *in, const char *out, const char *ufile, uint32_t clusters,
                     double threshold, int mode, double ratio, int distortion)
{
	struct qv_options_t o;
	memset(&o, 0, sizeof(o));
	o.mode = (uint8_t) mode;
	o.clusters = (uint8_t) clusters;
	o.distortion = (uint8_t) distortion;
	o.ratio = ratio;
	o.cluster_threshold = threshold;
	if (ufile) {
		o.uncompressed = 1;
		o.uncompressed_name = (char *) ufile;
	}
	srand(1);
	encode((char *) in, (char *) out, &o);
}

void ref_decode_file(const char *in, const char *out) {
	struct qv_options_t o;
	memset(&o, 0, sizeof(o));
	decode((char *) in, (char *) out, &o);
}

/* glibc rand() stream from seed 1, as initialize_kmeans_clustering consumes it. */
void ref_rand_stream(uint32_t count, int32_t *out) {
	uint32_t i;
	srand(1);
	for (i = 0; i < count; ++i) out[i] = rand();
}
