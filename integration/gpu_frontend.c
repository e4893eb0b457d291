/*
 * gpu_frontend.c -- the reference-side binding of the B200 front end: the three functions that the
 * reference's unmodified encode() calls (src/main.c:54, :62, :91), with the reference's own signatures,
 * implemented on top of include/qvz_gpu.h.
 *
 *     void     do_kmeans_clustering(struct quality_file_t *)                              include/cluster.h:24
 *     void     calculate_statistics(struct quality_file_t *)                              include/codebook.h:90
 *     uint32_t start_qv_compression(struct quality_file_t *, FILE *, double *, FILE *)    include/qv_compressor.h:95
 *
 * integration/Makefile compiles the reference's unmodified sources from where they lie (main.c included) with
 * the three originals renamed on their own translation units (-Ddo_kmeans_clustering=ref_do_kmeans_clustering ...),
 * and links this file + libqvz_gpu.so in their place.  Everything else -- argument parsing, load_file, codebook
 * design, the codebook text, the arithmetic coder, every printf -- is the reference's own object code.
 * There is no CPU fallback: if the device cannot be opened the program exits 1 like the reference does on its
 * own errors (src/main.c:43-46).
 *
 * This file is compiled against the reference's headers (-I$(REF)/include); it holds none of its code.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cluster.h"
#include "codebook.h"
#include "qv_compressor.h"

#include "qvz_gpu.h"

/* the originals, renamed by the build recipe on their own translation units */
void ref_calculate_statistics(struct quality_file_t *info);

static qvz_gpu *g;

static void die(const char *what, int rc) {
	printf("GPU front end: %s failed (%d): %s\n", what, rc, g ? qvz_gpu_last_error(g) : "no device");
	exit(1);
}

/* The lines of the mmap'ed file are contiguous, columns+1 bytes apart (src/lines.c:62-79): line 0 is the image. */
static void attach(struct quality_file_t *info) {
	int rc;
	if (g) return;
	rc = qvz_gpu_open(&g, getenv("QVZ_DEVICE") ? atoi(getenv("QVZ_DEVICE")) : 0);
	if (rc) die("qvz_gpu_open", rc);
	rc = qvz_gpu_load_rows(g, info->blocks[0].lines[0].m_data, info->lines, info->columns, info->columns + 1, 0);
	if (rc) die("qvz_gpu_load_rows", rc);
}

static struct line_t *line_at(struct quality_file_t *info, uint64_t n) {
	return &info->blocks[n / MAX_LINES_PER_BLOCK].lines[n % MAX_LINES_PER_BLOCK];
}

/* ---- src/cluster.c:212-244 ------------------------------------------------------------------------ */
void do_kmeans_clustering(struct quality_file_t *info) {
	const uint32_t K = info->cluster_count, C = info->columns;
	uint8_t *init = (uint8_t *) malloc((size_t) K * C), *means = (uint8_t *) malloc((size_t) K * C);
	uint8_t *ids = (uint8_t *) malloc(info->lines);
	uint32_t *counts = (uint32_t *) malloc(K * sizeof(uint32_t));
	double *moved = (double *) malloc(sizeof(double) * MAX_KMEANS_ITERATIONS * K);
	uint32_t iters = 0, it, j;
	uint64_t n;
	int rc;

	attach(info);
	initialize_kmeans_clustering(info);              /* the reference's own: rand() picks and "Chose block" lines */
	for (j = 0; j < K; ++j) memcpy(init + (size_t) j * C, info->clusters->clusters[j].mean, C);

	rc = qvz_gpu_kmeans(g, K, init, info->opts->cluster_threshold, MAX_KMEANS_ITERATIONS, ids, means, counts, moved, &iters);
	if (rc) die("qvz_gpu_kmeans", rc);               /* an empty cluster: the reference dies of SIGFPE at src/cluster.c:113 */

	for (n = 0; n < info->lines; ++n) line_at(info, n)->cluster = ids[n];
	for (j = 0; j < K; ++j) {
		memcpy(info->clusters->clusters[j].mean, means + (size_t) j * C, C);
		info->clusters->clusters[j].count = counts[j];
	}
	if (info->opts->verbose) {                       /* src/cluster.c:126-127, :236-238, :241-243 */
		for (it = 0; it < iters; ++it) {
			for (j = 0; j < K; ++j) printf("Cluster %d moved %f.\n", j, moved[(size_t) it * K + j]);
			printf("\n");
		}
		printf("\nTotal number of iterations: %d.\n", iters);
	}
	free(init); free(means); free(ids); free(counts); free(moved);
}

/* ---- src/codebook.c:185-220 ----------------------------------------------------------------------- */
void calculate_statistics(struct quality_file_t *info) {
	const uint32_t K = info->cluster_count, C = info->columns;
	const uint64_t rows = 1 + (uint64_t) ALPHABET_INDEX_SIZE_HINT * (C - 1);
	uint32_t *counts = (uint32_t *) malloc((size_t) qvz_gpu_cond_counts_len(K, C) * sizeof(uint32_t));
	uint32_t k, saved_blocks, x;
	uint64_t p;
	int rc;

	attach(info);
	rc = qvz_gpu_cond_counts(g, counts);
	if (rc) die("qvz_gpu_cond_counts", rc);
	/* install: pmfs[] is in get_cond_pmf order (src/codebook.c:116-120), which is the order of counts[] */
	for (k = 0; k < K; ++k) {
		struct cond_pmf_list_t *list = info->clusters->clusters[k].training_stats;
		for (p = 0; p < rows; ++p) {
			struct pmf_t *pmf = list->pmfs[p];
			const uint32_t *src = counts + ((size_t) k * rows + p) * ALPHABET_INDEX_SIZE_HINT;
			uint32_t total = 0;
			for (x = 0; x < ALPHABET_INDEX_SIZE_HINT; ++x) {
				pmf->counts[x] = src[x];
				total += src[x];
			}
			pmf->total = total;                      /* pmf_increment keeps total = sum of the counters (src/pmf.c:211-214) */
			pmf->pmf_ready = 0;
		}
	}
	free(counts);
	/* the marginal PMFs (src/codebook.c:208-219) are the reference's own code: its calculate_statistics over ZERO
	 * blocks counts nothing and then derives them from the counters installed above */
	saved_blocks = info->block_count;
	info->block_count = 0;
	ref_calculate_statistics(info);
	info->block_count = saved_blocks;
}

/* cond_quantizer_list_t of every cluster (include/codebook.h:61-69) -> struct qvz_flat_tables */
struct flat {
	struct qvz_flat_tables t;
	uint32_t *nctx;
	uint8_t *ctx_of, *qratio, *qmap, *smap;
	uint64_t *q_off;
};

static void flatten(struct quality_file_t *info, struct flat *f) {
	const uint32_t K = info->cluster_count, C = info->columns, A = ALPHABET_INDEX_SIZE_HINT;
	uint64_t nq = 0, q;
	uint32_t k, c, ctx, hi, v;
	for (k = 0; k < K; ++k)
		for (c = 0; c < C; ++c) nq += 2ull * info->clusters->clusters[k].qlist->input_alphabets[c]->size;
	f->nctx = (uint32_t *) calloc((size_t) K * C, sizeof(uint32_t));
	f->ctx_of = (uint8_t *) malloc((size_t) K * C * A);
	f->q_off = (uint64_t *) calloc((size_t) K * C, sizeof(uint64_t));
	f->qratio = (uint8_t *) calloc(nq / 2 + 1, 1);
	f->qmap = (uint8_t *) calloc(nq * A + 1, 1);
	f->smap = (uint8_t *) malloc(nq * A + 1);
	memset(f->ctx_of, QVZ_CTX_ABSENT, (size_t) K * C * A);
	memset(f->smap, 0xFF, nq * A + 1);
	q = 0;
	for (k = 0; k < K; ++k) {
		struct cond_quantizer_list_t *ql = info->clusters->clusters[k].qlist;
		for (c = 0; c < C; ++c) {
			const struct alphabet_t *in = ql->input_alphabets[c];
			const size_t kc = (size_t) k * C + c;
			f->nctx[kc] = in->size;
			f->q_off[kc] = q;
			for (v = 0; v < A; ++v)
				if (in->indexes[v] != ALPHABET_SYMBOL_NOT_FOUND) f->ctx_of[kc * A + v] = (uint8_t) in->indexes[v];
			for (ctx = 0; ctx < in->size; ++ctx) {
				f->qratio[q / 2 + ctx] = ql->qratio[c][ctx];
				for (hi = 0; hi < 2; ++hi) {
					const struct quantizer_t *qz = ql->q[c][2 * ctx + hi];
					const uint64_t qi = q + 2 * ctx + hi;
					memcpy(f->qmap + qi * A, qz->q, A);
					for (v = 0; v < A; ++v)
						if (qz->output_alphabet->indexes[v] != ALPHABET_SYMBOL_NOT_FOUND)
							f->smap[qi * A + v] = (uint8_t) qz->output_alphabet->indexes[v];
				}
			}
			q += 2ull * in->size;
		}
	}
	f->t.clusters = K;
	f->t.columns = C;
	f->t.nctx = f->nctx;
	f->t.ctx_of = f->ctx_of;
	f->t.q_off = f->q_off;
	f->t.qratio = f->qratio;
	f->t.qmap = f->qmap;
	f->t.smap = f->smap;
	f->t.distortion = info->dist->distortion;        /* 72 x 72 doubles, index x + 72*y (src/distortion.c:151-153) */
}

/* ---- src/qv_compressor.c:48-143 ------------------------------------------------------------------- */
uint32_t start_qv_compression(struct quality_file_t *info, FILE *fout, double *dis, FILE *funcompressed) {
	const uint32_t C = info->columns;
	const uint64_t N = info->lines;
	uint8_t *sym = (uint8_t *) malloc((size_t) N * C);
	uint8_t *qvimg = funcompressed ? (uint8_t *) malloc((size_t) N * (C + 1)) : NULL;
	double *err = (double *) malloc((size_t) N * sizeof(double));
	double distortion = 0.0;
	struct flat f;
	qv_compressor qvc;
	uint32_t osSize, s;
	uint64_t n;
	int rc;

	attach(info);
	/* the reference's own: seeds info->well (time-seeded, or 0x55555555 under -DDEBUG), writes the 128-byte seed,
	 * builds the coder's adaptive tables (src/qv_stream.c:66-125) */
	qvc = initialize_qv_compressor(fout, COMPRESSION, info);
	flatten(info, &f);
	rc = qvz_gpu_quantize(g, &f.t, info->well.state, sym, qvimg, err);
	if (rc) die("qvz_gpu_quantize", rc);

	/* the coder consumes the symbol stream in the reference's order (src/qv_compressor.c:86,96,117): per line the cluster
	 * id, then one (state, quantizer index) per column; the quantizer index is 2*ctx + hi with ctx = index of the
	 * previous quantized value in the column's input alphabet (src/codebook.c:162-171) */
	for (n = 0; n < N; ++n) {
		const uint8_t cluster = line_at(info, n)->cluster;
		struct cond_quantizer_list_t *ql = info->clusters->clusters[cluster].qlist;
		const uint8_t *ls = sym + (size_t) n * C;
		uint8_t prev = 0;
		if (info->opts->verbose && n % MAX_LINES_PER_BLOCK == 0) printf("Line: %dM\n", (int) (n / MAX_LINES_PER_BLOCK));
		qv_write_cluster(qvc->Quals, cluster);
		for (s = 0; s < C; ++s) {
			const uint32_t idx = 2 * ql->input_alphabets[s]->indexes[prev] + (ls[s] >> 7);
			const uint32_t state = ls[s] & 0x7F;
			compress_qv(qvc->Quals, state, cluster, s, idx);
			prev = ql->q[s][idx]->output_alphabet->symbols[state];
		}
		distortion += err[n];                        /* error / columns, added in line order (src/qv_compressor.c:127) */
	}
	if (funcompressed) fwrite(qvimg, 1, (size_t) N * (C + 1), funcompressed);       /* qv+33 per symbol, '\n' per line (:100-125) */
	osSize = encoder_last_step(qvc->Quals->a, qvc->Quals->os);
	if (dis) *dis = distortion / ((double) info->lines);

	free(sym); free(qvimg); free(err);
	free(f.nctx); free(f.ctx_of); free(f.q_off); free(f.qratio); free(f.qmap); free(f.smap);
	qvz_gpu_close(g);
	g = NULL;
	return osSize;
}
