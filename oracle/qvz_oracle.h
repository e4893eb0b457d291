/*
 * qvz_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's data-parallel front end (k-means, conditional counts,
 * WELL1024a, quantize walk).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product (qvz_b200/) never does.
 *
 * Parity is PINNED: tests/test_oracle.py checks every function below against the unmodified
 * reference compiled into oracle/_ref/libqvzref.so, against the survey's known-answer vectors
 *  and against the fixtures in tests/golden/ that were generated from the
 * reference by tests/golden/make_golden.py.
 */
#ifndef QVZ_ORACLE_H
#define QVZ_ORACLE_H

#include <stdint.h>
#include "../include/qvz_gpu.h"

struct oracle_well {
	uint32_t state[32];
	uint32_t n;
	uint32_t bit_output;
	uint32_t bits_left;
};

void     oracle_well_seed(struct oracle_well *w, const uint32_t seed[32]);
uint32_t oracle_well_next(struct oracle_well *w);
uint32_t oracle_well_bits7(struct oracle_well *w);
void     oracle_well_words(const uint32_t seed[32], uint64_t skip, uint64_t count, uint32_t *out);
void     oracle_well_draws(const uint32_t seed[32], uint64_t count, uint8_t *out);
/* rotated-frame state (u[i] = s[(n+i)&31]) after `words` steps */
void     oracle_well_state_after(const uint32_t seed[32], uint64_t words, uint32_t state_out[32]);

/* returns iteration count, or -1 if a cluster became empty (the reference would SIGFPE) */
int32_t oracle_kmeans(const uint8_t *rows, uint64_t n_lines, uint32_t columns, uint32_t row_stride,
                      uint32_t K, const uint8_t *init_means, double threshold, uint32_t max_iter,
                      uint8_t *cluster_ids, uint8_t *means_out, uint32_t *counts_out,
                      double *moved_log);

void oracle_cond_counts(const uint8_t *rows, uint64_t n_lines, uint32_t columns, uint32_t row_stride,
                        uint32_t K, const uint8_t *cluster_ids, uint32_t *counts);

/* returns mean distortion (*dis of start_qv_compression), or a negative value on a missing context */
double oracle_quantize(const uint8_t *rows, uint64_t n_lines, uint32_t columns, uint32_t row_stride,
                       uint64_t first_line, const uint8_t *cluster_ids,
                       const struct qvz_flat_tables *t, const uint32_t seed[32],
                       uint8_t *symbols, uint8_t *qv_image, double *line_err);

#endif
