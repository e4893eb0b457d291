/*
 * qvz_gpu.h -- C ABI of the B200-native qvz compression front end.
 *
 * This is the drop-in boundary for the three data-parallel calls that the reference's
 * encode() makes (reference: src/main.c:54, :62, :91):
 *
 *     do_kmeans_clustering(&qv_info)        include/cluster.h:24        -> qvz_gpu_kmeans
 *     calculate_statistics(&qv_info)        include/codebook.h:90       -> qvz_gpu_cond_counts
 *     start_qv_compression(&qv_info, ...)   include/qv_compressor.h:95  -> qvz_gpu_quantize
 *
 * plus the ingest that replaces the mmap + per-line pointer table of load_file
 * (src/lines.c:27-82) with one resident copy in HBM              -> qvz_gpu_load_rows.
 *
 * Plain pointers and sizes only; no C++/torch types; every call returns an int status
 * (0 = ok) and never throws.  All *host* pointers may be pageable or pinned memory.
 * Functions ending in _dev take DEVICE pointers owned by the caller and are the
 * stepping interface used for the multi-GPU path (the caller all-reduces the integer
 * buffers between steps, e.g. with NCCL).
 *
 * Results are bit-exact with the reference on identical inputs: cluster ids, iteration
 * count, centroid bytes, every conditional counter, every emitted (state, hi) symbol,
 * the `-u` image and the per-line distortion doubles.
 */
#ifndef QVZ_GPU_H
#define QVZ_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QVZ_ALPHABET 72u            /* ALPHABET_SIZE, src/main.c:13; ALPHABET_INDEX_SIZE_HINT, include/pmf.h:11 */
#define QVZ_MAX_COLUMNS 1022u       /* MAX_READS_PER_LINE, include/lines.h:13 */
#define QVZ_MAX_KMEANS_ITER 1000u   /* MAX_KMEANS_ITERATIONS, include/cluster.h:9 */
#define QVZ_MAX_CLUSTERS 255u       /* line_t.cluster and qv_options_t.clusters are uint8_t (include/lines.h:26, include/codebook.h:31) */
#define QVZ_CTX_ABSENT 0xFFu        /* ALPHABET_SYMBOL_NOT_FOUND (include/pmf.h:9) narrowed to a byte */

/* status codes */
#define QVZ_OK 0
#define QVZ_ERR_CUDA 1              /* a CUDA call failed; see qvz_gpu_last_error */
#define QVZ_ERR_ARG 2               /* bad argument / call order */
#define QVZ_ERR_EMPTY_CLUSTER 3     /* a cluster lost all its lines: the reference divides by zero here (src/cluster.c:113) */
#define QVZ_ERR_SYMBOL_RANGE 4      /* a quality byte outside ['!', '!'+71]: the reference indexes out of bounds (src/pmf.c:372-381) */
#define QVZ_ERR_CONTEXT 5           /* quantize reached a context with no quantizer: the reference asserts (src/codebook.c:164) */
#define QVZ_ERR_UNSUPPORTED 6

typedef struct qvz_gpu qvz_gpu;     /* one handle = one device = one shard of lines */

/*
 * Pointer-free mirror of `struct cond_quantizer_list_t` (include/codebook.h:61-69) for all
 * clusters, as produced by generate_codebooks (src/codebook.c:355-468) or read_codebooks (:560-581).
 * kc = k*columns + c.  Quantizer (k, c, q_idx = 2*ctx + hi) has pool index q_off[kc] + q_idx.
 */
struct qvz_flat_tables {
	uint32_t clusters;
	uint32_t columns;
	const uint32_t *nctx;        /* [K*C]      input_alphabets[c]->size                       */
	const uint8_t  *ctx_of;      /* [K*C*72]   input_alphabets[c]->indexes[prev qv] or 0xFF   */
	const uint64_t *q_off;       /* [K*C]      pool index of quantizer (k, c, 0); always even */
	const uint8_t  *qratio;      /* [nq/2]     qratio[c][ctx] at q_off[kc]/2 + ctx, 0..128     */
	const uint8_t  *qmap;        /* [nq*72]    quantizer_t.q[x]: input symbol -> quantized    */
	const uint8_t  *smap;        /* [nq*72]    output_alphabet->indexes[qv] -> state or 0xFF  */
	const double   *distortion;  /* [72*72]    distortion_t.distortion, index x + 72*y        */
};

/* Device-side durations (CUDA events on the library's stream) of the last call of each stage, ms. */
struct qvz_gpu_timings {
	float load_h2d_ms;           /* host->device copy of the raw rows                           */
	float load_layout_ms;        /* on-device re-layout (strip '\n', pack, interleave)           */
	float kmeans_ms;             /* all iterations: assign+accumulate and recenter kernels       */
	float kmeans_assign_ms;      /* sum over iterations of the assign+accumulate kernel alone    */
	float cond_counts_ms;        /* conditional-count kernel(s) incl. table zeroing              */
	float quantize_setup_ms;     /* main stream before the walk: table upload/composition, waiting for the draws */
	float quantize_ms;           /* quantize_draws_ms + duration of the walk kernel              */
	float quantize_draws_ms;     /* WELL draw generator kernel (auxiliary stream, overlapped with the setup) */
	float quantize_d2h_ms;       /* output re-layout + device->host copies                       */
	uint32_t kmeans_iters;
	uint32_t kernel_launches;    /* kernels launched by this handle since the last reset         */
};

/* ---- life cycle ------------------------------------------------------------------------- */
int  qvz_gpu_open(qvz_gpu **out, int device);
void qvz_gpu_close(qvz_gpu *h);
const char *qvz_gpu_last_error(const qvz_gpu *h);
/* cudaStream_t of the handle as an opaque pointer (for ordering against caller-side collectives). */
void *qvz_gpu_stream(qvz_gpu *h);
int  qvz_gpu_get_timings(qvz_gpu *h, struct qvz_gpu_timings *out);
int  qvz_gpu_reset_launch_count(qvz_gpu *h);

/* ---- ingest: replaces load_file/alloc_blocks (src/lines.c:27-126) ----------------------- */
/* rows: n_lines rows of `columns` raw ASCII quality bytes ('!'+q, offset NOT removed), row
 * pitch `row_stride` >= columns (columns+1 for a '\n'-terminated file image).
 * first_line: global index of rows[0] in the whole file (0 on a single GPU; a multiple of 4 on
 * a shard) -- it positions this shard in the WELL draw stream (draw = line*columns + column). */
int qvz_gpu_load_rows(qvz_gpu *h, const uint8_t *rows, uint64_t n_lines, uint32_t columns,
                      uint32_t row_stride, uint64_t first_line);

/* ---- stage 1: k-means, replaces do_kmeans_clustering (src/cluster.c:212-244) ------------ */
/* init_means: K*columns raw bytes = the rows picked by initialize_kmeans_clustering
 * (src/cluster.c:192-206; the host keeps calling libc rand() for the picks).
 * Outputs (each may be NULL): cluster_ids_out[n_lines]; means_out[K*columns] and counts_out[K]
 * = cluster_t.mean/.count after the last recentering; moved_log_out[max_iter*K] = the
 * "Cluster %d moved %f." values (src/cluster.c:127); iters_out = iteration count (:242). */
int qvz_gpu_kmeans(qvz_gpu *h, uint32_t K, const uint8_t *init_means, double threshold,
                   uint32_t max_iter, uint8_t *cluster_ids_out, uint8_t *means_out,
                   uint32_t *counts_out, double *moved_log_out, uint32_t *iters_out);
/* Install cluster ids computed elsewhere (e.g. read back from a previous run). */
int qvz_gpu_set_clusters(qvz_gpu *h, uint32_t K, const uint8_t *cluster_ids);

/* stepping form (multi-GPU): sums_dev = int64[K*columns + K] (column sums then line counts).
 * begin -> { assign_dev -> caller all-reduces sums_dev -> update_dev } until converged -> end. */
int qvz_gpu_kmeans_begin(qvz_gpu *h, uint32_t K, const uint8_t *init_means);
int qvz_gpu_kmeans_assign_dev(qvz_gpu *h, int64_t *sums_dev);
int qvz_gpu_kmeans_update_dev(qvz_gpu *h, const int64_t *sums_dev, double *moved_out /* K */,
                              uint32_t *counts_out /* K or NULL */);
int qvz_gpu_kmeans_end(qvz_gpu *h, uint8_t *cluster_ids_out, uint8_t *means_out);
/* The same loop without a host wait per iteration: update_async recentres AND evaluates the loop condition of
 * do_kmeans_clustering (moved > threshold, iteration < max_iter; src/cluster.c:221-234) on the device.  The caller
 * enqueues iteration i+1 (assign_dev, all-reduce, update_async) before it polls the outcome of iteration i: kernels
 * enqueued after the run has ended return at once.  poll(idx) waits for the idx-th update_async of the run (at most 3
 * later ones may be in flight); result() hands out the iteration count, the "Cluster %d moved %f." values
 * (iters x K doubles, src/cluster.c:127) and cluster_t.count, and reports an empty cluster. */
int qvz_gpu_kmeans_update_async(qvz_gpu *h, const int64_t *sums_dev, double threshold, uint32_t max_iter);
int qvz_gpu_kmeans_poll(qvz_gpu *h, uint32_t idx, int *done, uint32_t *iters);
int qvz_gpu_kmeans_result(qvz_gpu *h, uint32_t *iters_out, double *moved_log_out, uint32_t *counts_out);
/* the same two steps with the sums crossing to HOST memory, for a single process that drives several devices
 * and adds the (<= 6 KB of) integer sums itself: assign_host copies this shard's sums out, update_host takes the
 * totals of all shards back in and recentres. */
int qvz_gpu_kmeans_assign_host(qvz_gpu *h, int64_t *sums_out);
int qvz_gpu_kmeans_update_host(qvz_gpu *h, const int64_t *sums_in, double *moved_out /* K */,
                               uint32_t *counts_out /* K or NULL */);

/* ---- stage 2: conditional counts, replaces the counting loop of calculate_statistics
 *      (src/codebook.c:193-205) + pmf_increment (src/pmf.c:211-214) ----------------------- */
/* counts_out: K * (1 + 72*(columns-1)) * 72 uint32 in get_cond_pmf order (src/codebook.c:116-120):
 * row 0 = column 0; row 1 + (col-1)*72 + prev = column col after raw previous value prev.
 * pmf_t.total of a row is the sum of its 72 counters. */
int qvz_gpu_cond_counts(qvz_gpu *h, uint32_t *counts_out);
int qvz_gpu_cond_counts_dev(qvz_gpu *h, uint32_t *counts_dev);
uint64_t qvz_gpu_cond_counts_len(uint32_t K, uint32_t columns);   /* number of uint32 */

/* ---- stage 3: quantize walk, replaces the per-line loop of start_qv_compression
 *      (src/qv_compressor.c:76-135) incl. choose_quantizer (src/codebook.c:162-171) and the
 *      WELL1024a bit server (src/well.c:8-46) ---------------------------------------------- */
/* well_seed: the 32 words written to the file by initialize_arithStream (src/qv_stream.c:76-90).
 * symbols_out[n_lines*columns]: q_state | hi<<7 per (line, column), line-major: exactly what
 *   compress_qv needs (state, and q_idx = 2*ctx + hi with ctx recomputable from the previous qv).
 * qv_out[n_lines*(columns+1)]  (NULL = skip): the `-u` image, qv+33 per symbol and '\n' per line.
 * line_err_out[n_lines]        (NULL = skip): error/columns per line (qv_compressor.c:127); the
 *   caller adds them in line order and divides by the line count to get *dis. */
int qvz_gpu_quantize(qvz_gpu *h, const struct qvz_flat_tables *t, const uint32_t well_seed[32],
                     uint8_t *symbols_out, uint8_t *qv_out, double *line_err_out);
/* The tables of `t` -> device memory, once; qvz_gpu_quantize(h, NULL, ...) then walks with them (they are an INPUT of
 * stage 3 -- qlist in the reference, src/qv_compressor.c:84 -- so a caller that quantizes the same rows again, or a
 * measurement of the walk with resident inputs, need not upload them per call).  Valid until the next load_rows. */
int qvz_gpu_upload_tables(qvz_gpu *h, const struct qvz_flat_tables *t);

/* Optional: start generating the WELL draws for `well_seed` now, on the handle's auxiliary stream.  The draws
 * depend only on the seed and on the resident rows' shape, so a caller that knows the seed early (the reference
 * draws it right before start_qv_compression, src/qv_stream.c:76-90, but nothing depends on that order) can overlap
 * their generation with k-means / counting / codebook design.  qvz_gpu_quantize with the same seed then uses them. */
int qvz_gpu_prefetch_draws(qvz_gpu *h, const uint32_t well_seed[32]);

/* ---- WELL1024a helpers (src/well.c:8-24), used by tests and by the decoder-side host ------ */
/* state_out = state after `words` calls of well_1024a starting from seed (n = 0 frame). */
int qvz_gpu_well_jump(qvz_gpu *h, const uint32_t seed[32], uint64_t words, uint32_t state_out[32]);

#ifdef __cplusplus
}
#endif
#endif /* QVZ_GPU_H */
