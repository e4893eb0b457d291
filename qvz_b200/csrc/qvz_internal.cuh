// qvz_internal.cuh -- shared declarations of the B200 (sm_100a) qvz front end.
//
// Device-resident data layout:
//
//   slot p  <->  line n :   the shard's lines are cut into T "runs" of Lr consecutive lines
//                           (Lr % 4 == 0: every run starts on a WELL word boundary; T % 4096 == 0, so P % 4096 == 0);
//                           run r, step i  =  line r*Lr + i  =  slot p = i*T + r.
//                           Thread r of the quantize kernel walks run r sequentially, so its WELL1024a
//                           stream is one contiguous piece of the reference's draw stream, while
//                           adjacent threads touch adjacent slots => every access below is coalesced.
//   Xw[c4][p]  uint32       bytes 4*c4 .. 4*c4+3 of the line in slot p (raw ASCII, '\n' stripped,
//                           zero padded past the last column / past the last line)
//   cl[p]      uint8        cluster id of slot p (0xFF = slot holds no line)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <vector>

#include "../../include/qvz_gpu.h"

#define QVZ_THREADS 256
#define QVZ_RUN_ALIGN 4096    // runs per shard are a multiple of this: one batch of the walk = 4096 slots of ONE step (quantize.cu), P % 4096 == 0
#define QVZ_NFLAGS 8
#define QVZ_NO_LINE 0xFFu
#define QVZ_MAX_K 16            // register-resident distances in the fused k-means kernel (more clusters: kmeans_wide.cu)
#define QVZ_CTL_ITER 0          // km_ctl[]: iterations completed in this k-means run
#define QVZ_CTL_DONE 1          //           the run has left do_kmeans_clustering's loop (src/cluster.c:221)
#define QVZ_CTL_WORDS 4
#define QVZ_KM_RING 4           // host-visible copies of km_ctl in flight (abi.cu: speculative enqueue)

struct qvz_layout {
	uint64_t n_lines;    // lines in this shard
	uint64_t first_line; // global index of line 0 of the shard
	uint32_t C;          // columns
	uint32_t C4;         // ceil(C/4) words per line
	uint32_t Lr;         // lines per run (multiple of 4)
	uint32_t T;          // runs (multiple of QVZ_THREADS)
	uint64_t P;          // slots = T * Lr
};

static __host__ __device__ __forceinline__ uint64_t qvz_slot_line(const qvz_layout &L, uint64_t p) {
	uint64_t i = p / L.T, r = p - i * L.T;
	return r * L.Lr + i;
}

struct qvz_well_cache;   // well.cu

struct qvz_gpu {
	int device;
	cudaStream_t stream;
	char err[512];
	int sm_count;

	qvz_layout L;
	uint32_t *Xw;            // [C4][P]
	uint8_t *Xb;             // [C][P] the same rows as one byte plane per column, made on first use by the K == 1 counting pass (cond_counts.cu)
	int Xb_valid;            // Xb holds the resident rows
	uint8_t *cl;             // [P]
	uint32_t K;              // clusters currently installed in cl (0 = none)
	int *flags;              // device [8]: 0 symbol range, 1 empty cluster, 2 missing context, 3 malformed tables,
	                         //             4 scratch max quantized value, 5 largest symbol (byte-33) in the rows
	int *h_flags;            // pinned mirror

	// k-means state
	uint32_t km_K;
	uint8_t *means_b;        // [K][C] current centroids (raw ASCII bytes)
	uint32_t *means_w;       // [K][C4] same, packed for dp4a (zero padded)
	uint32_t *means_sq;      // [K] sum of squares of each centroid
	uint32_t *means_t;       // [C4][KP] the packed centroids transposed and padded: what the assign kernel reads (kmeans.cu: constant bank copy)
	size_t means_t_cap;
	int cm_slot;             // this handle's centroid slot of the constant bank, -1 = none
	int64_t *sums;           // [K*C + K] column sums, then line counts
	int64_t *k1_sums;        // this shard's running sums of the current k-means run (kmeans.cu): K == 1 reuses them, K >= 2 updates them
	size_t k1_cap;
	int k1_valid;
	double *moved;           // [K] device
	double *h_moved;         // pinned [256]
	int64_t *h_counts;       // pinned [256] line counts
	uint32_t *km_ctl;        // device [QVZ_CTL_WORDS]: iteration count and the loop decision, written by the update kernel
	double *moved_log;       // device [QVZ_MAX_KMEANS_ITER][K]: "Cluster %d moved %f." of every iteration (src/cluster.c:127)
	int64_t *last_counts;    // device [K]: cluster_t.count after the last recentering
	uint32_t *h_ctl;         // pinned [QVZ_KM_RING][QVZ_CTL_WORDS]
	cudaEvent_t ev_iter[QVZ_KM_RING];
	uint32_t km_enq;         // update launches enqueued in this run
	size_t moved_log_cap, last_counts_cap;
	std::vector<cudaEvent_t> *km_ev;   // event pairs around every assign launch of the run (timings are read lazily)
	uint32_t km_ev_used;
	uint32_t *counts_dev;    // conditional-count table of the host-pointer entry point
	int counts_cached;       // counts_dev holds the table of the resident rows and ids (left there by the K == 1 k-means pass)
	size_t means_b_cap, means_w_cap, means_sq_cap, sums_cap, moved_cap, counts_cap;

	// quantize state
	uint32_t *W;             // [K][C][72 prev][72 data] -> qv_lo | qv_hi<<8 | state_lo<<16 | state_hi<<24
	uint8_t *R;              // [K][C][72 prev] qratio, 0xFF = no such context
	double *D;               // [72*72]
	size_t W_cap, R_cap;
	uint8_t *flat;           // staging copy of the caller's flat tables
	size_t flat_cap;
	uint32_t *run_states;    // [T][32] WELL state (n = 0 frame) at the first draw of each run
	uint32_t *Yw, *Qw;       // [C4][P] packed outputs (state|hi<<7, qv+33)
	uint32_t *Dw;            // [Lr*C/4 (+2)][T] the 7-bit WELL draws of every run in sequence order, one byte per draw
	uint8_t *G;              // table images of the batched walk, one per column (quantize.cu)
	size_t G_cap;
	uint8_t *rowmap, *reach; // [K][C][72]: row of (cluster, context value) in its column's image; reachability scratch
	uint32_t *start;         // [K] the variant a line of cluster k starts from at column 0
	uint32_t *support;       // [K][C][3] bit x set: data value x occurs at (cluster, column) in the resident rows (cond_counts.cu)
	size_t rowmap_cap, reach_cap, start_cap, support_cap;
	int support_valid;
	uint32_t support_K;
	uint32_t smax;           // largest symbol value in the resident rows
	double *Ep;              // [P] per-slot error / C
	qvz_well_cache *well;

	// draw generation overlapped on its own stream (abi.cu:start_draws)
	cudaStream_t aux_stream;
	cudaEvent_t ev_draws_start, ev_jump_done, ev_draws, ev_walk_done;
	int draws_state, walk_recorded;
	uint32_t draws_seed[32];

	// host <-> device pipeline: two staging buffers, a copy stream, one event pair per buffer
	cudaStream_t copy_stream;
	uint8_t *stage[2];
	size_t stage_bytes;
	uint8_t *pin[2];         // pinned bounce buffers for pageable host memory (abi.cu)
	size_t pin_bytes;
	cudaEvent_t ev_copied[2], ev_consumed[2];
	size_t Xb_cap;
	size_t Xw_cap, cl_cap, Yw_cap, Qw_cap, Dw_cap, Ep_cap, rs_cap;

	// quantizer tables resident on the device (qvz_gpu_upload_tables)
	uint32_t tab_K, tab_C, tab_A, tab_rows, tab_hrows;    // tab_A = 0: the line-major walk (W / R only)
	uint32_t tab_box;                // alphabet box of the images: > every symbol of the rows and every value a quantizer can emit for one
	int tab_dmode, tab_valid, tab_toeplitz, tab_dm, tab_support_used;
	int tab_nodraw;                  // no reachable context mixes its two quantizers: the walk needs no draws (quantize.cu)

	// events / timings: recorded without synchronising, turned into milliseconds by qvz_gpu_get_timings
	cudaEvent_t ev[8];
	cudaEvent_t ev_km[2], ev_cc[2], ev_q[4];
	int tm_km_pending, tm_cc_pending, tm_q_pending;
	qvz_gpu_timings tm;
};

#define QVZ_CUDA(h, call)                                                                         \
	do {                                                                                          \
		cudaError_t e__ = (call);                                                                 \
		if (e__ != cudaSuccess) {                                                                 \
			snprintf((h)->err, sizeof((h)->err), "%s:%d: %s: %s", __FILE__, __LINE__, #call,      \
			         cudaGetErrorString(e__));                                                    \
			return QVZ_ERR_CUDA;                                                                  \
		}                                                                                         \
	} while (0)

#define QVZ_FAIL(h, code, ...)                                                                    \
	do {                                                                                          \
		snprintf((h)->err, sizeof((h)->err), __VA_ARGS__);                                        \
		return (code);                                                                            \
	} while (0)

#define QVZ_LAUNCHED(h) ((h)->tm.kernel_launches += 1)

// layout.cu -- every call handles the whole runs [r0, r0+nr) = lines [r0*Lr, (r0+nr)*Lr); stage_dev row 0 = line r0*Lr
int qvz_layout_ingest(qvz_gpu *h, uint32_t r0, uint32_t nr, const uint8_t *stage_dev, uint32_t row_stride);
int qvz_layout_ids_to_lines(qvz_gpu *h, uint32_t r0, uint32_t nr, uint8_t *stage_dev);
int qvz_layout_ids_from_lines(qvz_gpu *h, uint32_t r0, uint32_t nr, const uint8_t *stage_dev, uint32_t K);
int qvz_layout_words_to_lines(qvz_gpu *h, uint32_t r0, uint32_t nr, const uint32_t *Yw, uint8_t *stage_dev,
                              uint32_t out_stride, int add_newline);
int qvz_layout_doubles_to_lines(qvz_gpu *h, uint32_t r0, uint32_t nr, const double *Ep, double *stage_dev);

// kmeans.cu
int qvz_kmeans_launch_assign(qvz_gpu *h, int64_t *sums_dev);
int qvz_kmeans_slot_acquire();
void qvz_kmeans_slot_release(int slot);
uint32_t qvz_kmeans_kp(uint32_t K);
int qvz_kmeans_launch_assign_wide(qvz_gpu *h, int64_t *sums_dev);      // kmeans_wide.cu: more than QVZ_MAX_K clusters
int qvz_kmeans_launch_update(qvz_gpu *h, const int64_t *sums_dev, double threshold, uint32_t max_iter);

// cond_counts.cu
int qvz_cond_counts_launch(qvz_gpu *h, uint32_t *counts_dev);
int qvz_cond_counts_support(qvz_gpu *h, const uint32_t *counts_dev);   // fills h->support from a count table of the resident rows

// well.cu
int qvz_well_init(qvz_gpu *h);
void qvz_well_free(qvz_gpu *h);
int qvz_well_run_states(qvz_gpu *h, const uint32_t seed[32]);          // fills h->run_states
void qvz_well_debug(qvz_gpu *h, const char *where);
int qvz_well_jump_state(qvz_gpu *h, const uint32_t seed[32], uint64_t words, uint32_t *state_dev);

// quantize.cu
int qvz_quantize_launch(qvz_gpu *h, int want_qv, int want_err, int toeplitz);
int qvz_quantize_draws(qvz_gpu *h);
int qvz_quantize_vmax(qvz_gpu *h, uint32_t KC, uint32_t smax);
int qvz_quantize_rows(qvz_gpu *h, uint32_t K, uint32_t C, uint32_t A, int compact, const uint32_t *support);
int qvz_quantize_compact(qvz_gpu *h, uint32_t K, uint32_t C, uint32_t A, uint32_t rows, uint32_t hrows, int fold);
uint32_t qvz_quantize_batched_group(uint32_t rows, uint32_t hrows, uint32_t A);
size_t qvz_quantize_image_bytes(uint32_t C, uint32_t rows, uint32_t hrows, uint32_t A);
int qvz_quantize_launch_batched(qvz_gpu *h, uint32_t rows, uint32_t hrows, uint32_t A, int want_qv, int dm);
int qvz_quantize_compose(qvz_gpu *h, uint32_t KC, const uint32_t *nctx, const uint8_t *ctx_of, const uint64_t *q_off,
                         const uint8_t *qratio, const uint8_t *qmap, const uint8_t *smap);
