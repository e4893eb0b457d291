/*
 * qvz_host.h -- C ABI of the host side of the B200-native qvz encoder: everything between the three GPU
 * stage calls of include/qvz_gpu.h.
 *
 *   conditional counts (GPU)  ->  qvz_host_design    codebook design, replaces the marginals of calculate_statistics
 *                                                     (src/codebook.c:208-219) + generate_codebooks (:355-468)
 *                             ->  qvz_host_tables     flat mirror of cond_quantizer_list_t = input of qvz_gpu_quantize
 *   symbol stream (GPU)       ->  qvz_host_encode     container writer: write_codebooks (src/codebook.c:474-555), the
 *                                                     WELL seed (src/qv_stream.c:76-93) and the adaptive arithmetic
 *                                                     coder fed in line order (src/qv_compressor.c:8-22,76-137,
 *                                                     src/arith.c:5-116, src/qv_stream.c:9-61, src/os_stream.c)
 *
 * Results are bit-exact with the reference: the same doubles in the same order (built with -ffp-contract=off),
 * the same tables, the same .qvz bytes.  These functions are sequential small-alphabet host work and stay on
 * the CPU by design (BASELINE.json north_star).
 */
#ifndef QVZ_HOST_H
#define QVZ_HOST_H

#include <stdint.h>

#include "qvz_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#define QVZ_MODE_RATIO 0            /* MODE_RATIO, include/codebook.h:22 : target = ratio * entropy of the context pmf */
#define QVZ_MODE_FIXED 1            /* MODE_FIXED, include/codebook.h:23 : target = fixed bits per symbol */
#define QVZ_DIST_MANHATTAN 1        /* DISTORTION_MANHATTAN, include/distortion.h:7 */
#define QVZ_DIST_MSE 2              /* DISTORTION_MSE,       include/distortion.h:8 */
#define QVZ_DIST_LORENTZ 3          /* DISTORTION_LORENTZ,   include/distortion.h:9 */

typedef struct qvz_codebooks qvz_codebooks;

/* 72x72 distortion matrix, index x + 72*y (src/distortion.c:50-93).  Returns 0, or -1 for an unknown type. */
int qvz_host_distortion(int type, double out[QVZ_ALPHABET * QVZ_ALPHABET]);
/* -D FILE: gen_custom_distortion (src/distortion.c:100-145).  Returns 0, or -1 if the file cannot be read. */
int qvz_host_distortion_file(const char *path, double out[QVZ_ALPHABET * QVZ_ALPHABET]);

/* counts: clusters * (1 + 72*(columns-1)) * 72 uint32 in get_cond_pmf order, as produced by qvz_gpu_cond_counts.
 * mode/target: opts->mode / opts->ratio.  threads: total host threads (0 = all the hardware has): clusters are
 * independent and are designed in parallel; inside a cluster the columns are sequential but the contexts of a column
 * are not, and share the threads that are left (<= 8 per cluster).  The tables do not depend on the thread count.
 * Returns NULL on bad arguments. */
qvz_codebooks *qvz_host_design(const uint32_t *counts, uint32_t clusters, uint32_t columns, int mode, double target,
                               const double distortion[QVZ_ALPHABET * QVZ_ALPHABET], int threads);
void qvz_host_free(qvz_codebooks *cb);

/* Fills *out with pointers into memory owned by cb (valid until qvz_host_free). */
int qvz_host_tables(const qvz_codebooks *cb, struct qvz_flat_tables *out);

/* Size in bytes of header + codebooks as write_codebooks emits them, and the bytes themselves. */
uint64_t qvz_host_codebook_bytes(const qvz_codebooks *cb);
int qvz_host_write_codebooks(const qvz_codebooks *cb, uint64_t n_lines, uint8_t *out);

/* Writes the whole .qvz file: header + codebooks, the 32 seed words, the arithmetic-coded stream of
 * (cluster id, then one symbol per column) per line.  symbols[n_lines*columns] = q_state | hi<<7 as produced
 * by qvz_gpu_quantize; cluster_ids[n_lines].  *stream_bytes_out = what start_qv_compression returns
 * (bytes written by the coder).  Returns 0, -1 if the file cannot be written, -2 on a malformed symbol. */
int qvz_host_encode(const qvz_codebooks *cb, const char *path, uint64_t n_lines, const uint8_t *cluster_ids,
                    const uint8_t *symbols, const uint32_t well_seed[32], uint64_t *stream_bytes_out);

/* decode() (src/main.c:132-160): read_codebooks + start_qv_decompression: writes the quantized lines of a .qvz file
 * (what `-u` dumped at encode time).  Returns 0, -1 on I/O errors, -2 on a malformed file. */
int qvz_host_decode(const char *in_path, const char *out_path, uint64_t *lines_out);

#ifdef __cplusplus
}
#endif
#endif /* QVZ_HOST_H */
