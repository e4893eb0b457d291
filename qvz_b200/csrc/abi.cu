// abi.cu -- extern "C" entry points of include/qvz_gpu.h.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

#include "qvz_internal.cuh"

#define QVZ_TARGET_RUNS (148u * 1024u)     // runs per shard (4 draw-generator CTAs per SM, a whole number of walk batches per step): enough threads for
                                           // the draw generator, few enough that the WELL jump-ahead (one F2 mat-vec per run) stays cheap

enum { EV_A = 0, EV_B, EV_C, EV_D, EV_E, EV_F };

static void free_dev(void *p) {
	if (p) cudaFree(p);
}

static void release_rows(qvz_gpu *h) {
	free_dev(h->Xw); h->Xw = nullptr; h->Xw_cap = 0;
	free_dev(h->Xb); h->Xb = nullptr; h->Xb_cap = 0;
	free_dev(h->cl); h->cl = nullptr; h->cl_cap = 0;
	free_dev(h->run_states); h->run_states = nullptr; h->rs_cap = 0;
	free_dev(h->Yw); h->Yw = nullptr; h->Yw_cap = 0;
	free_dev(h->Qw); h->Qw = nullptr; h->Qw_cap = 0;
	free_dev(h->Ep); h->Ep = nullptr; h->Ep_cap = 0;
	free_dev(h->Dw); h->Dw = nullptr; h->Dw_cap = 0;
	h->K = 0;
}

// grow-only device buffer: repeated calls on same-shaped inputs never touch the allocator
template <class T>
static int ensure_buf(qvz_gpu *h, T **p, size_t *cap, size_t bytes) {
	if (*p && *cap >= bytes) return QVZ_OK;
	free_dev(*p);
	*p = nullptr;
	*cap = 0;
	cudaError_t e = cudaMalloc(p, bytes);
	if (e == cudaErrorMemoryAllocation && h->Xb && (void *) p != (void *) &h->Xb) {
		cudaGetLastError();                          // the optional byte planes give way to a buffer that is needed
		free_dev(h->Xb);
		h->Xb = nullptr;
		h->Xb_cap = 0;
		e = cudaMalloc(p, bytes);
	}
	QVZ_CUDA(h, e);
	*cap = bytes;
	return QVZ_OK;
}

// ---- host <-> device pipeline ---------------------------------------------------------------------------
// The host side of every transfer is line-major, the device side is the packed slot layout, and a range of
// whole runs is a contiguous range of lines (layout.cu).  So a transfer is cut into pieces of whole runs that
// fit one of two staging buffers; piece k crosses PCIe on the copy stream while piece k-1 is re-laid out on
// the compute stream.  Events order the two streams per buffer; nothing blocks the host until the end.
#define QVZ_STAGE_BYTES ((size_t) 64 << 20)

static int ensure_stage(qvz_gpu *h, size_t min_bytes) {
	size_t want = QVZ_STAGE_BYTES > min_bytes ? QVZ_STAGE_BYTES : min_bytes;
	if (h->stage[0] && h->stage_bytes >= want) return QVZ_OK;
	QVZ_CUDA(h, cudaStreamSynchronize(h->stream));
	QVZ_CUDA(h, cudaStreamSynchronize(h->copy_stream));
	for (int b = 0; b < 2; ++b) {
		if (h->pin[b]) cudaFreeHost(h->pin[b]);
		free_dev(h->stage[b]);
		h->stage[b] = nullptr;
		QVZ_CUDA(h, cudaMalloc(&h->stage[b], want));
	}
	h->stage_bytes = want;
	return QVZ_OK;
}

struct piece {
	uint32_t r0, nr;         // runs
	uint64_t l0, l1;         // lines [l0, l1) that exist in this piece (l0 == l1: padding runs only)
	int buf;
};

// pieces of whole runs whose line-major image (row pitch `stride`) fits a staging buffer
static int plan_pieces(qvz_gpu *h, size_t stride, uint32_t *runs_per_piece) {
	const qvz_layout &L = h->L;
	const size_t per_run = (size_t) L.Lr * stride;
	int rc = ensure_stage(h, per_run);
	if (rc) return rc;
	uint64_t n = h->stage_bytes / per_run;
	if (n >= 64) n &= ~(uint64_t) 31;            // whole warps of runs: coalesced on the packed side
	if (n > L.T) n = L.T;
	*runs_per_piece = (uint32_t) n;
	return QVZ_OK;
}

static piece make_piece(const qvz_gpu *h, uint32_t r0, uint32_t per, int k) {
	const qvz_layout &L = h->L;
	piece pc;
	pc.r0 = r0;
	pc.nr = (L.T - r0 < per) ? L.T - r0 : per;
	pc.l0 = (uint64_t) r0 * L.Lr;
	pc.l1 = (uint64_t) (r0 + pc.nr) * L.Lr;
	if (pc.l0 > L.n_lines) pc.l0 = L.n_lines;
	if (pc.l1 > L.n_lines) pc.l1 = L.n_lines;
	pc.buf = k & 1;
	return pc;
}

// Host side of a piece.  Pinned (or registered) memory is handed to the copy engine as it is.  Pageable memory -- an
// mmap'ed file, a malloc'ed output array: what the reference's callers have (src/lines.c:62-79) -- goes through one of two
// PINNED bounce buffers, filled or drained by a few host threads while the other buffer's DMA is in flight: the driver's
// own path for pageable memory is a single-threaded staged copy.
static bool host_is_pinned(const void *p) {
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
		cudaGetLastError();
		return false;
	}
	return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

static void par_memcpy(void *dst, const void *src, size_t n) {
	unsigned nt = std::thread::hardware_concurrency() / 2;
	if (const char *e = getenv("QVZ_COPY_THREADS")) nt = (unsigned) atoi(e);
	if (nt > 8) nt = 8;
	if (nt < 2 || n < ((size_t) 8 << 20)) {
		memcpy(dst, src, n);
		return;
	}
	const size_t per = ((n + nt - 1) / nt + 4095) & ~(size_t) 4095;
	std::vector<std::thread> pool;
	for (unsigned t = 0; t < nt; ++t) {
		const size_t lo = (size_t) t * per, hi = lo + per < n ? lo + per : n;
		if (lo >= hi) break;
		pool.emplace_back([=]() { memcpy((char *) dst + lo, (const char *) src + lo, hi - lo); });
	}
	for (auto &t : pool) t.join();
}

static int ensure_pin(qvz_gpu *h) {
	if (h->pin[0] && h->pin_bytes >= h->stage_bytes) return QVZ_OK;
	for (int b = 0; b < 2; ++b) {
		if (h->pin[b]) cudaFreeHost(h->pin[b]);
		h->pin[b] = nullptr;
		QVZ_CUDA(h, cudaMallocHost(&h->pin[b], h->stage_bytes));
	}
	h->pin_bytes = h->stage_bytes;
	return QVZ_OK;
}

// host -> device.  consume(piece) launches the re-layout kernel on h->stream.
template <class F>
static int pipeline_h2d(qvz_gpu *h, const uint8_t *host, size_t stride, size_t row_bytes, F consume) {
	uint32_t per = 0;
	int rc = plan_pieces(h, stride, &per);
	if (rc) return rc;
	const bool bounce = !host_is_pinned(host) && !getenv("QVZ_NO_BOUNCE");
	if (bounce && (rc = ensure_pin(h))) return rc;
	bool used[2] = {false, false};
	int k = 0;
	for (uint32_t r0 = 0; r0 < h->L.T; r0 += per, ++k) {
		const piece pc = make_piece(h, r0, per, k);
		if (pc.l1 > pc.l0) {
			const size_t bytes = (size_t) (pc.l1 - pc.l0 - 1) * stride + row_bytes;
			const uint8_t *src = host + (size_t) pc.l0 * stride;
			if (bounce) {
				if (used[pc.buf]) QVZ_CUDA(h, cudaEventSynchronize(h->ev_copied[pc.buf]));   // the DMA out of this bounce buffer is done
				par_memcpy(h->pin[pc.buf], src, bytes);
				src = h->pin[pc.buf];
			}
			if (used[pc.buf]) QVZ_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_consumed[pc.buf], 0));
			QVZ_CUDA(h, cudaMemcpyAsync(h->stage[pc.buf], src, bytes, cudaMemcpyHostToDevice, h->copy_stream));
			QVZ_CUDA(h, cudaEventRecord(h->ev_copied[pc.buf], h->copy_stream));
			QVZ_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_copied[pc.buf], 0));
		}
		rc = consume(pc);
		if (rc) return rc;
		if (pc.l1 > pc.l0) {
			QVZ_CUDA(h, cudaEventRecord(h->ev_consumed[pc.buf], h->stream));
			used[pc.buf] = true;
		}
	}
	return QVZ_OK;
}

// device -> host.  produce(piece) launches the re-layout kernel on h->stream, writing stage[piece.buf].
template <class F>
static int pipeline_d2h(qvz_gpu *h, uint8_t *host, size_t stride, size_t row_bytes, F produce) {
	uint32_t per = 0;
	int rc = plan_pieces(h, stride, &per);
	if (rc) return rc;
	const bool bounce = !host_is_pinned(host) && !getenv("QVZ_NO_BOUNCE");
	if (bounce && (rc = ensure_pin(h))) return rc;
	bool used[2] = {false, false};
	uint8_t *pend_dst = nullptr;                     // bounce: the piece whose DMA has been enqueued but not yet drained to `host`
	size_t pend_bytes = 0;
	int pend_buf = 0;
	auto drain = [&]() -> int {
		if (!pend_dst) return QVZ_OK;
		QVZ_CUDA(h, cudaEventSynchronize(h->ev_consumed[pend_buf]));
		par_memcpy(pend_dst, h->pin[pend_buf], pend_bytes);
		pend_dst = nullptr;
		return QVZ_OK;
	};
	int k = 0;
	for (uint32_t r0 = 0; r0 < h->L.T; r0 += per, ++k) {
		const piece pc = make_piece(h, r0, per, k);
		if (pc.l1 == pc.l0) break;               // only padding runs from here on
		if (used[pc.buf]) QVZ_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_consumed[pc.buf], 0));
		rc = produce(pc);
		if (rc) return rc;
		QVZ_CUDA(h, cudaEventRecord(h->ev_copied[pc.buf], h->stream));
		QVZ_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_copied[pc.buf], 0));
		const size_t bytes = (size_t) (pc.l1 - pc.l0 - 1) * stride + row_bytes;
		uint8_t *dst = host + (size_t) pc.l0 * stride;
		// (bounce: pin[pc.buf] was drained before piece k-1 was enqueued, i.e. before this point)
		QVZ_CUDA(h, cudaMemcpyAsync(bounce ? h->pin[pc.buf] : dst, h->stage[pc.buf], bytes, cudaMemcpyDeviceToHost, h->copy_stream));
		QVZ_CUDA(h, cudaEventRecord(h->ev_consumed[pc.buf], h->copy_stream));
		used[pc.buf] = true;
		if (bounce) {
			rc = drain();                            // piece k-1 -> the caller's memory, while piece k crosses PCIe
			if (rc) return rc;
			pend_dst = dst;
			pend_bytes = bytes;
			pend_buf = pc.buf;
		}
	}
	rc = drain();
	if (rc) return rc;
	QVZ_CUDA(h, cudaStreamSynchronize(h->copy_stream));
	QVZ_CUDA(h, cudaStreamSynchronize(h->stream));
	return QVZ_OK;
}

static void release_kmeans(qvz_gpu *h) {
	free_dev(h->means_b); h->means_b = nullptr;
	free_dev(h->means_w); h->means_w = nullptr;
	free_dev(h->means_sq); h->means_sq = nullptr;
	free_dev(h->means_t); h->means_t = nullptr; h->means_t_cap = 0;
	free_dev(h->sums); h->sums = nullptr;
	free_dev(h->moved); h->moved = nullptr;
	free_dev(h->counts_dev); h->counts_dev = nullptr; h->counts_cached = 0;
	free_dev(h->k1_sums); h->k1_sums = nullptr; h->k1_cap = 0; h->k1_valid = 0;
	free_dev(h->km_ctl); h->km_ctl = nullptr;
	free_dev(h->moved_log); h->moved_log = nullptr; h->moved_log_cap = 0;
	free_dev(h->last_counts); h->last_counts = nullptr; h->last_counts_cap = 0;
	if (h->h_ctl) cudaFreeHost(h->h_ctl);
	h->h_ctl = nullptr;
	for (int i = 0; i < QVZ_KM_RING; ++i)
		if (h->ev_iter[i]) { cudaEventDestroy(h->ev_iter[i]); h->ev_iter[i] = nullptr; }
	if (h->km_ev) {
		for (cudaEvent_t e : *h->km_ev) cudaEventDestroy(e);
		delete h->km_ev;
		h->km_ev = nullptr;
	}
	h->means_b_cap = h->means_w_cap = h->means_sq_cap = h->sums_cap = h->moved_cap = h->counts_cap = 0;
	if (h->h_moved) cudaFreeHost(h->h_moved);
	if (h->h_counts) cudaFreeHost(h->h_counts);
	h->h_moved = nullptr;
	h->h_counts = nullptr;
	h->km_K = 0;
}

static float ev_ms(qvz_gpu *h, int a, int b) {
	float ms = 0.f;
	cudaEventElapsedTime(&ms, h->ev[a], h->ev[b]);
	return ms;
}

// read-and-clear one device flag (after the stream has been synchronised)
static int take_flag(qvz_gpu *h, int idx, int *value) {
	QVZ_CUDA(h, cudaMemcpyAsync(h->h_flags, h->flags, QVZ_NFLAGS * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
	QVZ_CUDA(h, cudaStreamSynchronize(h->stream));
	*value = h->h_flags[idx];
	if (*value) {
		h->h_flags[idx] = 0;
		QVZ_CUDA(h, cudaMemcpyAsync(h->flags + idx, h->h_flags + idx, sizeof(int), cudaMemcpyHostToDevice, h->stream));
		QVZ_CUDA(h, cudaStreamSynchronize(h->stream));
	}
	return QVZ_OK;
}

extern "C" int qvz_gpu_open(qvz_gpu **out, int device) {
	if (!out) return QVZ_ERR_ARG;
	*out = nullptr;
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return QVZ_ERR_CUDA;
	qvz_gpu *h = (qvz_gpu *) calloc(1, sizeof(qvz_gpu));
	if (!h) return QVZ_ERR_ARG;
	h->device = device;
	h->cm_slot = -1;
	*out = h;       // returned even on failure below so the caller can read last_error, then close
	QVZ_CUDA(h, cudaSetDevice(device));
	QVZ_CUDA(h, cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device));
	QVZ_CUDA(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
	QVZ_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
	{                                            // the draw generator is issue-bound: at high priority it shares the SMs with the
		int lo = 0, hi = 0;                      // memory-bound stages it overlaps instead of queueing behind their CTAs
		QVZ_CUDA(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
		QVZ_CUDA(h, cudaStreamCreateWithPriority(&h->aux_stream, cudaStreamNonBlocking, hi));
	}
	QVZ_CUDA(h, cudaEventCreate(&h->ev_draws_start));
	QVZ_CUDA(h, cudaEventCreate(&h->ev_draws));
	QVZ_CUDA(h, cudaEventCreate(&h->ev_jump_done));
	QVZ_CUDA(h, cudaEventCreateWithFlags(&h->ev_walk_done, cudaEventDisableTiming));
	for (int i = 0; i < 8; ++i) QVZ_CUDA(h, cudaEventCreate(&h->ev[i]));
	for (int i = 0; i < 2; ++i) {
		QVZ_CUDA(h, cudaEventCreate(&h->ev_km[i]));
		QVZ_CUDA(h, cudaEventCreate(&h->ev_cc[i]));
	}
	for (int i = 0; i < 4; ++i) QVZ_CUDA(h, cudaEventCreate(&h->ev_q[i]));
	for (int b = 0; b < 2; ++b) {
		QVZ_CUDA(h, cudaEventCreateWithFlags(&h->ev_copied[b], cudaEventDisableTiming));
		QVZ_CUDA(h, cudaEventCreateWithFlags(&h->ev_consumed[b], cudaEventDisableTiming));
	}
	QVZ_CUDA(h, cudaMalloc(&h->flags, QVZ_NFLAGS * sizeof(int)));
	QVZ_CUDA(h, cudaMemsetAsync(h->flags, 0, QVZ_NFLAGS * sizeof(int), h->stream));
	QVZ_CUDA(h, cudaMallocHost(&h->h_flags, QVZ_NFLAGS * sizeof(int)));
	QVZ_CUDA(h, cudaMalloc(&h->D, 72 * 72 * sizeof(double)));
	h->cm_slot = qvz_kmeans_slot_acquire();
	return qvz_well_init(h);
}

extern "C" void qvz_gpu_close(qvz_gpu *h) {
	if (!h) return;
	cudaSetDevice(h->device);
	if (h->stream) cudaStreamSynchronize(h->stream);
	if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
	if (h->aux_stream) cudaStreamSynchronize(h->aux_stream);
	release_rows(h);
	if (h->ev_draws_start) cudaEventDestroy(h->ev_draws_start);
	if (h->ev_draws) cudaEventDestroy(h->ev_draws);
	if (h->ev_jump_done) cudaEventDestroy(h->ev_jump_done);
	if (h->ev_walk_done) cudaEventDestroy(h->ev_walk_done);
	if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
	for (int b = 0; b < 2; ++b) {
		if (h->pin[b]) cudaFreeHost(h->pin[b]);
		free_dev(h->stage[b]);
		if (h->ev_copied[b]) cudaEventDestroy(h->ev_copied[b]);
		if (h->ev_consumed[b]) cudaEventDestroy(h->ev_consumed[b]);
	}
	if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
	release_kmeans(h);
	qvz_kmeans_slot_release(h->cm_slot);
	qvz_well_free(h);
	free_dev(h->W);
	free_dev(h->flat);
	free_dev(h->G);
	free_dev(h->rowmap);
	free_dev(h->reach);
	free_dev(h->start);
	free_dev(h->support);
	free_dev(h->R);
	free_dev(h->D);
	free_dev(h->flags);
	if (h->h_flags) cudaFreeHost(h->h_flags);
	for (int i = 0; i < 8; ++i)
		if (h->ev[i]) cudaEventDestroy(h->ev[i]);
	for (int i = 0; i < 2; ++i) {
		if (h->ev_km[i]) cudaEventDestroy(h->ev_km[i]);
		if (h->ev_cc[i]) cudaEventDestroy(h->ev_cc[i]);
	}
	for (int i = 0; i < 4; ++i)
		if (h->ev_q[i]) cudaEventDestroy(h->ev_q[i]);
	if (h->stream) cudaStreamDestroy(h->stream);
	free(h);
}

extern "C" const char *qvz_gpu_last_error(const qvz_gpu *h) { return h ? h->err : "null handle"; }
extern "C" void *qvz_gpu_stream(qvz_gpu *h) { return h ? (void *) h->stream : nullptr; }

// The k-means and counting stages only RECORD events (no host synchronisation inside the stage calls); the
// milliseconds are worked out here, when somebody asks.
static int settle_timings(qvz_gpu *h) {
	if (h->tm_km_pending) {
		QVZ_CUDA(h, cudaEventSynchronize(h->ev_km[1]));
		float ms = 0.f, sum = 0.f;
		cudaEventElapsedTime(&ms, h->ev_km[0], h->ev_km[1]);
		h->tm.kmeans_ms = ms;
		for (uint32_t i = 0; h->km_ev && i + 1 < h->km_ev_used; i += 2) {
			cudaEventElapsedTime(&ms, (*h->km_ev)[i], (*h->km_ev)[i + 1]);
			if (getenv("QVZ_DEBUG_KM")) fprintf(stderr, "[kmeans] assign launch %u: %.3f ms\n", i / 2, ms);
			sum += ms;
		}
		h->tm.kmeans_assign_ms = sum;
		h->tm_km_pending = 0;
	}
	if (h->tm_cc_pending) {
		QVZ_CUDA(h, cudaEventSynchronize(h->ev_cc[1]));
		float ms = 0.f;
		cudaEventElapsedTime(&ms, h->ev_cc[0], h->ev_cc[1]);
		h->tm.cond_counts_ms = ms;
		h->tm_cc_pending = 0;
	}
	return QVZ_OK;
}

extern "C" int qvz_gpu_get_timings(qvz_gpu *h, struct qvz_gpu_timings *out) {
	if (!h || !out) return QVZ_ERR_ARG;
	QVZ_CUDA(h, cudaSetDevice(h->device));
	int rc = settle_timings(h);
	if (rc) return rc;
	*out = h->tm;
	return QVZ_OK;
}

extern "C" int qvz_gpu_reset_launch_count(qvz_gpu *h) {
	if (!h) return QVZ_ERR_ARG;
	h->tm.kernel_launches = 0;
	return QVZ_OK;
}

extern "C" uint64_t qvz_gpu_cond_counts_len(uint32_t K, uint32_t columns) {
	return (uint64_t) K * (1 + (uint64_t) QVZ_ALPHABET * (columns - 1)) * QVZ_ALPHABET;
}

// ------------------------------------------------------------------------------------------ ingest
extern "C" int qvz_gpu_load_rows(qvz_gpu *h, const uint8_t *rows, uint64_t n_lines, uint32_t columns,
                                 uint32_t row_stride, uint64_t first_line)
{
	if (!h || !rows) return QVZ_ERR_ARG;
	if (n_lines == 0 || columns == 0 || columns > QVZ_MAX_COLUMNS || row_stride < columns)
		QVZ_FAIL(h, QVZ_ERR_ARG, "load_rows: need n_lines > 0, 0 < columns <= %u, row_stride >= columns", QVZ_MAX_COLUMNS);
	if (first_line & 3) QVZ_FAIL(h, QVZ_ERR_ARG, "load_rows: first_line must be a multiple of 4");
	QVZ_CUDA(h, cudaSetDevice(h->device));
	h->K = 0;                                    // the resident cluster ids belong to the previous rows
	QVZ_CUDA(h, cudaStreamSynchronize(h->aux_stream));
	h->draws_state = 0;                          // ... and so do prefetched draws
	h->k1_valid = 0;
	h->counts_cached = 0;
	h->support_valid = 0;
	h->tab_valid = 0;                            // table images are built for the resident rows (alphabet box, reachable rows)

	qvz_layout &L = h->L;
	L.n_lines = n_lines;
	L.first_line = first_line;
	L.C = columns;
	L.C4 = (columns + 3) / 4;
	uint64_t target_runs = QVZ_TARGET_RUNS;
	if (const char *e = getenv("QVZ_TARGET_RUNS")) target_runs = strtoull(e, nullptr, 10) ? strtoull(e, nullptr, 10) : target_runs;   // tuning knob
	uint64_t lr = (n_lines + target_runs - 1) / target_runs;
	lr = (lr + 3) & ~3ull;                        // multiple of 4: every run starts on a WELL word boundary (T % 4096 == 0 gives P % 4096 == 0)
	if (lr < 4) lr = 4;
	L.Lr = (uint32_t) lr;
	uint64_t runs = (n_lines + lr - 1) / lr;
	L.T = (uint32_t) ((runs + QVZ_RUN_ALIGN - 1) / QVZ_RUN_ALIGN * QVZ_RUN_ALIGN);
	L.P = (uint64_t) L.T * L.Lr;

	int rc = ensure_buf(h, &h->Xw, &h->Xw_cap, (size_t) L.C4 * L.P * sizeof(uint32_t));
	if (rc) return rc;
	h->Xb_valid = 0;                             // the byte planes of the one-cluster counting pass are made on first use (cond_counts.cu)
	rc = ensure_buf(h, &h->cl, &h->cl_cap, (size_t) L.P);
	if (rc) return rc;
	QVZ_CUDA(h, cudaMemsetAsync(h->flags + 5, 0, sizeof(int), h->stream));
	QVZ_CUDA(h, cudaEventRecord(h->ev[EV_A], h->stream));
	QVZ_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev[EV_A], 0));
	rc = pipeline_h2d(h, rows, row_stride, columns, [&](const piece &pc) {
		return qvz_layout_ingest(h, pc.r0, pc.nr, h->stage[pc.buf], row_stride);
	});
	if (rc) return rc;
	QVZ_CUDA(h, cudaEventRecord(h->ev[EV_B], h->copy_stream));      // the last byte has crossed PCIe
	QVZ_CUDA(h, cudaEventRecord(h->ev[EV_C], h->stream));           // ... and has been packed
	int bad = 0;
	rc = take_flag(h, 0, &bad);                  // synchronises the compute stream
	if (rc) return rc;
	QVZ_CUDA(h, cudaStreamSynchronize(h->copy_stream));
	h->smax = (uint32_t) h->h_flags[5];             // take_flag copied all flags back
	h->tm.load_h2d_ms = ev_ms(h, EV_A, EV_B);
	h->tm.load_layout_ms = ev_ms(h, EV_B, EV_C);    // re-layout time NOT hidden behind the copy
	if (h->tm.load_layout_ms < 0.f) h->tm.load_layout_ms = 0.f;
	if (bad) {
		release_rows(h);
		QVZ_FAIL(h, QVZ_ERR_SYMBOL_RANGE, "load_rows: a quality byte is outside ['!', '!'+71]");
	}
	return QVZ_OK;
}

// ------------------------------------------------------------------------------------------ k-means
extern "C" int qvz_gpu_kmeans_begin(qvz_gpu *h, uint32_t K, const uint8_t *init_means) {
	if (!h || !init_means) return QVZ_ERR_ARG;
	if (!h->Xw) QVZ_FAIL(h, QVZ_ERR_ARG, "kmeans: no rows loaded");
	if (K == 0 || K > QVZ_MAX_CLUSTERS) QVZ_FAIL(h, QVZ_ERR_ARG, "kmeans: 1 <= clusters <= %u (the cluster id is a uint8_t, include/lines.h:26)", QVZ_MAX_CLUSTERS);
	QVZ_CUDA(h, cudaSetDevice(h->device));
	const uint32_t C = h->L.C, C4 = h->L.C4;
	h->km_K = K;
	int rc = ensure_buf(h, &h->means_b, &h->means_b_cap, (size_t) K * C);
	if (!rc) rc = ensure_buf(h, &h->means_w, &h->means_w_cap, (size_t) K * C4 * sizeof(uint32_t));
	if (!rc) rc = ensure_buf(h, &h->means_sq, &h->means_sq_cap, K * sizeof(uint32_t));
	if (!rc && K <= QVZ_MAX_K) {
		const size_t tb = (size_t) C4 * qvz_kmeans_kp(K) * sizeof(uint32_t);
		rc = ensure_buf(h, &h->means_t, &h->means_t_cap, tb);
		if (!rc) QVZ_CUDA(h, cudaMemsetAsync(h->means_t, 0, tb, h->stream));      // the padding centroids are zero words
	}
	if (!rc) rc = ensure_buf(h, &h->sums, &h->sums_cap, ((size_t) K * C + K) * sizeof(int64_t));
	if (!rc) rc = ensure_buf(h, &h->moved, &h->moved_cap, K * sizeof(double));
	if (!rc) rc = ensure_buf(h, &h->k1_sums, &h->k1_cap, ((size_t) K * C + K) * sizeof(int64_t));   // the run's local running sums
	if (!rc) rc = ensure_buf(h, &h->moved_log, &h->moved_log_cap, (size_t) QVZ_MAX_KMEANS_ITER * K * sizeof(double));
	if (!rc) rc = ensure_buf(h, &h->last_counts, &h->last_counts_cap, K * sizeof(int64_t));
	if (!rc && K == 1)                           // K == 1 takes its column sums from the count table (kmeans.cu)
		rc = ensure_buf(h, &h->counts_dev, &h->counts_cap, (size_t) qvz_gpu_cond_counts_len(1, C) * sizeof(uint32_t));
	if (rc) return rc;
	h->k1_valid = 0;                             // a new run reads the rows again
	h->counts_cached = 0;
	h->support_valid = 0;                        // the ids are about to change
	if (!h->h_moved) QVZ_CUDA(h, cudaMallocHost(&h->h_moved, QVZ_MAX_CLUSTERS * sizeof(double)));
	if (!h->h_counts) QVZ_CUDA(h, cudaMallocHost(&h->h_counts, QVZ_MAX_CLUSTERS * sizeof(int64_t)));
	if (!h->km_ctl) QVZ_CUDA(h, cudaMalloc(&h->km_ctl, QVZ_CTL_WORDS * sizeof(uint32_t)));
	if (!h->h_ctl) QVZ_CUDA(h, cudaMallocHost(&h->h_ctl, QVZ_KM_RING * QVZ_CTL_WORDS * sizeof(uint32_t)));
	for (int i = 0; i < QVZ_KM_RING; ++i)
		if (!h->ev_iter[i]) QVZ_CUDA(h, cudaEventCreateWithFlags(&h->ev_iter[i], cudaEventDisableTiming));
	if (!h->km_ev) h->km_ev = new std::vector<cudaEvent_t>();
	h->km_enq = 0;
	h->km_ev_used = 0;
	QVZ_CUDA(h, cudaMemsetAsync(h->km_ctl, 0, QVZ_CTL_WORDS * sizeof(uint32_t), h->stream));
	QVZ_CUDA(h, cudaMemcpyAsync(h->means_b, init_means, (size_t) K * C, cudaMemcpyHostToDevice, h->stream));
	h->K = K;
	QVZ_CUDA(h, cudaEventRecord(h->ev_km[0], h->stream));
	return qvz_kmeans_launch_update(h, nullptr, 0.0, 0);     // pack the initial centroids
}

extern "C" int qvz_gpu_kmeans_assign_dev(qvz_gpu *h, int64_t *sums_dev) {
	if (!h || !sums_dev || !h->km_K) return QVZ_ERR_ARG;
	QVZ_CUDA(h, cudaSetDevice(h->device));
	std::vector<cudaEvent_t> &pool = *h->km_ev;              // one event pair per assign launch, read in settle_timings
	while (pool.size() < (size_t) h->km_ev_used + 2) {
		cudaEvent_t e;
		QVZ_CUDA(h, cudaEventCreate(&e));
		pool.push_back(e);
	}
	QVZ_CUDA(h, cudaEventRecord(pool[h->km_ev_used], h->stream));
	int rc = h->km_K > QVZ_MAX_K ? qvz_kmeans_launch_assign_wide(h, sums_dev) : qvz_kmeans_launch_assign(h, sums_dev);
	if (rc) return rc;
	QVZ_CUDA(h, cudaEventRecord(pool[h->km_ev_used + 1], h->stream));
	h->km_ev_used += 2;
	return QVZ_OK;
}

// recalculate_means + the loop decision of do_kmeans_clustering (src/cluster.c:221-234) on the device: nothing here
// waits for the GPU.  The caller may enqueue the NEXT iteration right away (its kernels return at once if this one
// ended the run) and ask for the outcome later with qvz_gpu_kmeans_poll.
extern "C" int qvz_gpu_kmeans_update_async(qvz_gpu *h, const int64_t *sums_dev, double threshold, uint32_t max_iter) {
	if (!h || !sums_dev || !h->km_K) return QVZ_ERR_ARG;
	QVZ_CUDA(h, cudaSetDevice(h->device));
	int rc = qvz_kmeans_launch_update(h, sums_dev, threshold, max_iter);
	if (rc) return rc;
	const uint32_t slot = h->km_enq % QVZ_KM_RING;
	QVZ_CUDA(h, cudaMemcpyAsync(h->h_ctl + slot * QVZ_CTL_WORDS, h->km_ctl, QVZ_CTL_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
	QVZ_CUDA(h, cudaEventRecord(h->ev_iter[slot], h->stream));
	QVZ_CUDA(h, cudaEventRecord(h->ev_km[1], h->stream));
	h->tm_km_pending = 1;
	h->km_enq += 1;
	return QVZ_OK;
}

// outcome of the idx-th update of this run (0-based); at most QVZ_KM_RING - 1 later updates may have been enqueued
extern "C" int qvz_gpu_kmeans_poll(qvz_gpu *h, uint32_t idx, int *done, uint32_t *iters) {
	if (!h || !h->km_K || idx >= h->km_enq || h->km_enq - idx > QVZ_KM_RING) return QVZ_ERR_ARG;
	QVZ_CUDA(h, cudaSetDevice(h->device));
	const uint32_t slot = idx % QVZ_KM_RING;
	QVZ_CUDA(h, cudaEventSynchronize(h->ev_iter[slot]));
	if (done) *done = (int) h->h_ctl[slot * QVZ_CTL_WORDS + QVZ_CTL_DONE];
	if (iters) *iters = h->h_ctl[slot * QVZ_CTL_WORDS + QVZ_CTL_ITER];
	return QVZ_OK;
}

// after the run: iteration count, the "Cluster %d moved %f." log (iters x K, at most QVZ_MAX_KMEANS_ITER rows) and cluster_t.count
extern "C" int qvz_gpu_kmeans_result(qvz_gpu *h, uint32_t *iters_out, double *moved_log_out, uint32_t *counts_out) {
	if (!h || !h->km_K) return QVZ_ERR_ARG;
	QVZ_CUDA(h, cudaSetDevice(h->device));
	const uint32_t K = h->km_K;
	QVZ_CUDA(h, cudaMemcpyAsync(h->h_ctl, h->km_ctl, QVZ_CTL_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
	QVZ_CUDA(h, cudaMemcpyAsync(h->h_counts, h->last_counts, K * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
	int empty = 0;
	int rc = take_flag(h, 1, &empty);                // synchronises the stream
	if (rc) return rc;
	const uint32_t iters = h->h_ctl[QVZ_CTL_ITER];
	h->tm.kmeans_iters = iters;
	if (iters_out) *iters_out = iters;
	if (counts_out)
		for (uint32_t k = 0; k < K; ++k) counts_out[k] = (uint32_t) h->h_counts[k];
	if (moved_log_out && iters) {
		const uint32_t rows = iters < QVZ_MAX_KMEANS_ITER ? iters : QVZ_MAX_KMEANS_ITER;
		QVZ_CUDA(h, cudaMemcpyAsync(moved_log_out, h->moved_log, (size_t) rows * K * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
		QVZ_CUDA(h, cudaStreamSynchronize(h->stream));
	}
	if (empty) QVZ_FAIL(h, QVZ_ERR_EMPTY_CLUSTER, "kmeans: a cluster has no lines (the reference divides by zero, src/cluster.c:113)");
	return QVZ_OK;
}

// blocking form of one recentering: the caller decides whether to go on (kept for callers that add the sums themselves)
extern "C" int qvz_gpu_kmeans_update_dev(qvz_gpu *h, const int64_t *sums_dev, double *moved_out, uint32_t *counts_out) {
	if (!h || !sums_dev || !h->km_K) return QVZ_ERR_ARG;
	QVZ_CUDA(h, cudaSetDevice(h->device));
	const uint32_t K = h->km_K;
	int rc = qvz_gpu_kmeans_update_async(h, sums_dev, -1.0, 0xFFFFFFFFu);      // moved >= 0 > -1: never ends the run by itself
	if (rc) return rc;
	QVZ_CUDA(h, cudaMemcpyAsync(h->h_moved, h->moved, K * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
	QVZ_CUDA(h, cudaMemcpyAsync(h->h_counts, h->last_counts, K * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
	int empty = 0;
	rc = take_flag(h, 1, &empty);                    // synchronises the stream
	if (rc) return rc;
	for (uint32_t k = 0; k < K; ++k) {
		if (moved_out) moved_out[k] = h->h_moved[k];
		if (counts_out) counts_out[k] = (uint32_t) h->h_counts[k];
	}
	if (empty) QVZ_FAIL(h, QVZ_ERR_EMPTY_CLUSTER, "kmeans: a cluster has no lines (the reference divides by zero, src/cluster.c:113)");
	return QVZ_OK;
}

extern "C" int qvz_gpu_kmeans_assign_host(qvz_gpu *h, int64_t *sums_out) {
	if (!h || !sums_out || !h->km_K) return QVZ_ERR_ARG;
	int rc = qvz_gpu_kmeans_assign_dev(h, h->sums);
	if (rc) return rc;
	QVZ_CUDA(h, cudaMemcpyAsync(sums_out, h->sums, ((size_t) h->km_K * h->L.C + h->km_K) * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
	QVZ_CUDA(h, cudaStreamSynchronize(h->stream));
	return QVZ_OK;
}

extern "C" int qvz_gpu_kmeans_update_host(qvz_gpu *h, const int64_t *sums_in, double *moved_out, uint32_t *counts_out) {
	if (!h || !sums_in || !h->km_K) return QVZ_ERR_ARG;
	QVZ_CUDA(h, cudaSetDevice(h->device));
	QVZ_CUDA(h, cudaMemcpyAsync(h->sums, sums_in, ((size_t) h->km_K * h->L.C + h->km_K) * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
	return qvz_gpu_kmeans_update_dev(h, h->sums, moved_out, counts_out);
}

static int ids_to_host(qvz_gpu *h, uint8_t *ids_out) {
	return pipeline_d2h(h, ids_out, 1, 1, [&](const piece &pc) {
		return qvz_layout_ids_to_lines(h, pc.r0, pc.nr, h->stage[pc.buf]);
	});
}

extern "C" int qvz_gpu_kmeans_end(qvz_gpu *h, uint8_t *cluster_ids_out, uint8_t *means_out) {
	if (!h || !h->km_K) return QVZ_ERR_ARG;
	QVZ_CUDA(h, cudaSetDevice(h->device));
	if (means_out) {
		QVZ_CUDA(h, cudaMemcpyAsync(means_out, h->means_b, (size_t) h->km_K * h->L.C, cudaMemcpyDeviceToHost, h->stream));
		QVZ_CUDA(h, cudaStreamSynchronize(h->stream));
	}
	if (cluster_ids_out) return ids_to_host(h, cluster_ids_out);
	return QVZ_OK;
}

extern "C" int qvz_gpu_kmeans(qvz_gpu *h, uint32_t K, const uint8_t *init_means, double threshold,
                              uint32_t max_iter, uint8_t *cluster_ids_out, uint8_t *means_out,
                              uint32_t *counts_out, double *moved_log_out, uint32_t *iters_out)
{
	int rc = qvz_gpu_kmeans_begin(h, K, init_means);
	if (rc) return rc;
	// do_kmeans_clustering: while (iter_count < MAX_KMEANS_ITERATIONS && loop)  (src/cluster.c:221).  The loop
	// condition is evaluated on the device (kmeans.cu: update kernel), and iteration i+1 is enqueued BEFORE the
	// host looks at the outcome of iteration i: if i ended the run, the extra launches return immediately; if
	// not, the GPU never waited for the host.
	auto enqueue = [&]() -> int {
		int r = qvz_gpu_kmeans_assign_dev(h, h->sums);
		return r ? r : qvz_gpu_kmeans_update_async(h, h->sums, threshold, max_iter);
	};
	if (max_iter) {
		rc = enqueue();
		if (rc) return rc;
		for (uint32_t i = 0;; ++i) {
			if (i + 1 < max_iter) {
				rc = enqueue();
				if (rc) return rc;
			}
			int done = 0;
			rc = qvz_gpu_kmeans_poll(h, i, &done, nullptr);
			if (rc) return rc;
			if (done || i + 1 >= max_iter) break;
		}
	}
	rc = qvz_gpu_kmeans_result(h, iters_out, moved_log_out, counts_out);
	if (rc) return rc;
	return qvz_gpu_kmeans_end(h, cluster_ids_out, means_out);
}

extern "C" int qvz_gpu_set_clusters(qvz_gpu *h, uint32_t K, const uint8_t *cluster_ids) {
	if (!h || !cluster_ids) return QVZ_ERR_ARG;
	if (!h->Xw) QVZ_FAIL(h, QVZ_ERR_ARG, "set_clusters: no rows loaded");
	if (K == 0 || K > QVZ_MAX_CLUSTERS) QVZ_FAIL(h, QVZ_ERR_ARG, "set_clusters: bad cluster count");
	QVZ_CUDA(h, cudaSetDevice(h->device));
	h->counts_cached = 0;
	h->support_valid = 0;
	h->k1_valid = 0;                             // running sums of a k-means run in progress no longer match the ids
	h->K = 0;
	int rc = pipeline_h2d(h, cluster_ids, 1, 1, [&](const piece &pc) {
		return qvz_layout_ids_from_lines(h, pc.r0, pc.nr, h->stage[pc.buf], K);
	});
	if (rc) return rc;
	int bad = 0;
	rc = take_flag(h, 6, &bad);                  // synchronises the stream
	if (rc) return rc;
	if (bad) QVZ_FAIL(h, QVZ_ERR_ARG, "set_clusters: a cluster id is >= %u", K);
	h->K = K;
	return QVZ_OK;
}

// ------------------------------------------------------------------------------------------ counts
extern "C" int qvz_gpu_cond_counts_dev(qvz_gpu *h, uint32_t *counts_dev) {
	if (!h || !counts_dev) return QVZ_ERR_ARG;
	if (!h->Xw || !h->K) QVZ_FAIL(h, QVZ_ERR_ARG, "cond_counts: rows and cluster ids must be resident first");
	QVZ_CUDA(h, cudaSetDevice(h->device));
	QVZ_CUDA(h, cudaEventRecord(h->ev_cc[0], h->stream));
	if (h->counts_cached && h->K == 1) {
		// the K == 1 k-means pass of these rows already counted them (kmeans.cu): same table, no second pass
		if (counts_dev != h->counts_dev)
			QVZ_CUDA(h, cudaMemcpyAsync(counts_dev, h->counts_dev, (size_t) qvz_gpu_cond_counts_len(1, h->L.C) * sizeof(uint32_t),
			                            cudaMemcpyDeviceToDevice, h->stream));
	} else {
		int rc = qvz_cond_counts_launch(h, counts_dev);
		if (rc) return rc;
	}
	QVZ_CUDA(h, cudaEventRecord(h->ev_cc[1], h->stream));
	{                                            // which values occur where: lets the quantize stage stage reachable table rows only
		int rc = qvz_cond_counts_support(h, (h->counts_cached && h->K == 1) ? h->counts_dev : counts_dev);
		if (rc) return rc;
	}
	h->tm_cc_pending = 1;                        // no host synchronisation here: the table is ordered on the handle's stream
	return QVZ_OK;
}

extern "C" int qvz_gpu_cond_counts(qvz_gpu *h, uint32_t *counts_out) {
	if (!h) return QVZ_ERR_ARG;
	if (!h->Xw || !h->K) QVZ_FAIL(h, QVZ_ERR_ARG, "cond_counts: rows and cluster ids must be resident first");
	QVZ_CUDA(h, cudaSetDevice(h->device));
	const size_t bytes = (size_t) qvz_gpu_cond_counts_len(h->K, h->L.C) * sizeof(uint32_t);
	int rc = ensure_buf(h, &h->counts_dev, &h->counts_cap, bytes);
	if (rc) return rc;
	rc = qvz_gpu_cond_counts_dev(h, h->counts_dev);
	if (!rc && counts_out) {
		QVZ_CUDA(h, cudaMemcpyAsync(counts_out, h->counts_dev, bytes, cudaMemcpyDeviceToHost, h->stream));
		QVZ_CUDA(h, cudaStreamSynchronize(h->stream));
	}
	return rc;
}

// ------------------------------------------------------------------------------------------ quantize
// Upload the caller's flat tables and compose W / R (see quantize.cu) on the device.
static int upload_tables(qvz_gpu *h, const struct qvz_flat_tables *t, int *toeplitz) {
	const uint32_t K = t->clusters, C = t->columns;
	const size_t KC = (size_t) K * C;
	const uint64_t nq = t->q_off[KC - 1] + 2ull * t->nctx[KC - 1];
	// device staging layout (all offsets 8-byte aligned)
	const size_t o_qoff = 0, o_nctx = o_qoff + KC * 8, o_ctx = (o_nctx + KC * 4 + 7) & ~(size_t) 7;
	const size_t o_ratio = (o_ctx + KC * 72 + 7) & ~(size_t) 7, o_qmap = (o_ratio + nq / 2 + 7) & ~(size_t) 7;
	const size_t o_smap = (o_qmap + nq * 72 + 7) & ~(size_t) 7, total = o_smap + nq * 72;
	if (h->flat_cap < total) {
		free_dev(h->flat);
		h->flat = nullptr;
		QVZ_CUDA(h, cudaMalloc(&h->flat, total));
		h->flat_cap = total;
	}
	const size_t w_bytes = KC * 72 * 72 * sizeof(uint32_t), r_bytes = KC * 72;
	if (h->W_cap < w_bytes) {
		free_dev(h->W);
		h->W = nullptr;
		QVZ_CUDA(h, cudaMalloc(&h->W, w_bytes));
		h->W_cap = w_bytes;
	}
	if (h->R_cap < r_bytes) {
		free_dev(h->R);
		h->R = nullptr;
		QVZ_CUDA(h, cudaMalloc(&h->R, r_bytes));
		h->R_cap = r_bytes;
	}
	const cudaMemcpyKind H2D = cudaMemcpyHostToDevice;
	QVZ_CUDA(h, cudaMemcpyAsync(h->flat + o_qoff, t->q_off, KC * 8, H2D, h->stream));
	QVZ_CUDA(h, cudaMemcpyAsync(h->flat + o_nctx, t->nctx, KC * 4, H2D, h->stream));
	QVZ_CUDA(h, cudaMemcpyAsync(h->flat + o_ctx, t->ctx_of, KC * 72, H2D, h->stream));
	QVZ_CUDA(h, cudaMemcpyAsync(h->flat + o_ratio, t->qratio, nq / 2, H2D, h->stream));
	QVZ_CUDA(h, cudaMemcpyAsync(h->flat + o_qmap, t->qmap, nq * 72, H2D, h->stream));
	QVZ_CUDA(h, cudaMemcpyAsync(h->flat + o_smap, t->smap, nq * 72, H2D, h->stream));
	QVZ_CUDA(h, cudaMemcpyAsync(h->D, t->distortion, 72 * 72 * sizeof(double), H2D, h->stream));
	// distortion mode: 1 = a function of |x - y| only (true for -d M / L / A; a custom -D matrix may not be),
	// 2 = additionally integer-valued and small enough that a line's sum fits uint32 (-d M, -d A)
	*toeplitz = 1;
	for (uint32_t y = 0; y < 72 && *toeplitz; ++y)
		for (uint32_t x = 0; x < 72; ++x) {
			const uint32_t d = x > y ? x - y : y - x;
			if (memcmp(&t->distortion[x + 72 * y], &t->distortion[d], sizeof(double)) != 0) {
				*toeplitz = 0;
				break;
			}
		}
	if (*toeplitz) {
		bool integral = true;
		for (uint32_t d = 0; d < 72; ++d) {
			const double f = t->distortion[d];
			if (!(f >= 0.0 && f <= 2097152.0 && f == floor(f))) integral = false;     // 1022 columns * 2^21 < 2^32
		}
		if (integral) *toeplitz = 2;
	}
	return qvz_quantize_compose(h, (uint32_t) KC, (const uint32_t *) (h->flat + o_nctx), h->flat + o_ctx,
	                            (const uint64_t *) (h->flat + o_qoff), h->flat + o_ratio, h->flat + o_qmap,
	                            h->flat + o_smap);
}

// WELL jump-ahead + draw generation depend on the seed and the layout only -- not on the rows, the clusters or the
// tables -- so they run on their own stream, overlapped with whatever the main stream is doing (table upload and
// composition inside qvz_gpu_quantize; k-means / counts when the caller prefetches).  draws_state: 0 = nothing,
// 1 = run states only, 2 = run states + draws, for draws_seed.
static int start_draws(qvz_gpu *h, const uint32_t seed[32], bool with_draws) {
	const qvz_layout &L = h->L;
	int rc = ensure_buf(h, &h->run_states, &h->rs_cap, (size_t) L.T * 32 * sizeof(uint32_t));
	// draw words of a run in sequence order, Lr*C/4 rows of T words (quantize.cu); the walk reads up to two rows ahead
	const size_t draw_words = (size_t) L.Lr * L.C / 4 * L.T;
	if (!rc && with_draws) rc = ensure_buf(h, &h->Dw, &h->Dw_cap, (draw_words + 2 * (size_t) L.T) * sizeof(uint32_t));
	if (rc) return rc;
	if (h->walk_recorded) QVZ_CUDA(h, cudaStreamWaitEvent(h->aux_stream, h->ev_walk_done, 0));   // the last walk still reads Dw
	// the two spare rows are only ever read into bytes the walk discards; keep them defined
	if (with_draws) QVZ_CUDA(h, cudaMemsetAsync(h->Dw + draw_words, 0, 2 * (size_t) L.T * sizeof(uint32_t), h->aux_stream));
	cudaStream_t main_stream = h->stream;
	h->stream = h->aux_stream;                   // well.cu / quantize.cu launch on h->stream
	cudaError_t e = cudaEventRecord(h->ev_draws_start, h->aux_stream);
	if (e == cudaSuccess) rc = qvz_well_run_states(h, seed);
	if (e == cudaSuccess && !rc) e = cudaEventRecord(h->ev_jump_done, h->aux_stream);
	if (e == cudaSuccess && !rc && with_draws) rc = qvz_quantize_draws(h);
	if (e == cudaSuccess && !rc) e = cudaEventRecord(h->ev_draws, h->aux_stream);
	h->stream = main_stream;
	if (rc) return rc;
	if (e != cudaSuccess) QVZ_FAIL(h, QVZ_ERR_CUDA, "draw generation: %s", cudaGetErrorString(e));
	memcpy(h->draws_seed, seed, sizeof(h->draws_seed));
	h->draws_state = with_draws ? 2 : 1;
	return QVZ_OK;
}

// Could the batched walk be possible?  (decided for good in qvz_gpu_upload_tables; a wrong guess costs an unused draw pass)
static bool batched_possible(const qvz_gpu *h) {
	uint32_t A = (h->smax + 2) & ~1u;
	return !getenv("QVZ_FORCE_LINE_MAJOR") && A <= 62;
}

extern "C" int qvz_gpu_prefetch_draws(qvz_gpu *h, const uint32_t well_seed[32]) {
	if (!h || !well_seed) return QVZ_ERR_ARG;
	if (!h->Xw) QVZ_FAIL(h, QVZ_ERR_ARG, "prefetch_draws: no rows loaded");
	QVZ_CUDA(h, cudaSetDevice(h->device));
	return start_draws(h, well_seed, h->tab_valid ? h->tab_A != 0 : batched_possible(h));
}

// The per-column images of the batched walk from the resident W / R (quantize.cu).  use_support: seed the reachability
// pass with the data values the counting stage saw per (cluster, column); the images then hold exactly the rows these
// lines can reach.  Synchronises once (the image geometry depends on the tables' content).
static int build_images(qvz_gpu *h, bool use_support) {
	const uint32_t K = h->tab_K, C = h->tab_C, A = h->tab_box;
	uint32_t rows = 0, hrows = 0;
	int compact = 0;
	h->tab_support_used = 0;
	if (!getenv("QVZ_FORCE_LINE_MAJOR") && A <= 62) {
		compact = !getenv("QVZ_NO_REACH");
		const uint32_t *support = (use_support && compact && h->support_valid && h->support_K >= K && !getenv("QVZ_NO_SUPPORT")) ? h->support : nullptr;
		QVZ_CUDA(h, cudaMemsetAsync(h->flags + 7, 0, sizeof(int), h->stream));
		QVZ_CUDA(h, cudaMemsetAsync(h->flags + 4, 0, sizeof(int), h->stream));
		int rc = qvz_quantize_rows(h, K, C, A, compact, support);
		if (rc) return rc;
		QVZ_CUDA(h, cudaMemcpyAsync(h->h_flags, h->flags, QVZ_NFLAGS * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
		QVZ_CUDA(h, cudaStreamSynchronize(h->stream));
		rows = 1 + (uint32_t) h->h_flags[7];         // + the poison row
		hrows = compact ? 1 + (uint32_t) h->h_flags[4] : rows;      // rows of the hi plane: the poison row + the contexts that mix lo and hi
		if (rows > 256 || qvz_quantize_batched_group(rows, hrows, A) == 0) rows = 0;
		if (rows && support) h->tab_support_used = 1;
	}
	if (rows) {
		int rc = ensure_buf(h, &h->G, &h->G_cap, qvz_quantize_image_bytes(C, rows, hrows, A));
		if (rc) return rc;
		rc = qvz_quantize_compact(h, K, C, A, rows, hrows, compact);
		if (rc) return rc;
	}
	h->tab_A = rows ? A : 0;
	h->tab_rows = rows;
	h->tab_hrows = hrows;
	// the hi plane holds the poison row only: every context the resident rows can reach always takes the same one of its two
	// quantizers (qratio 0 or 128), folded into the lo plane with the ratio byte "always lo" -- no draw can change a symbol
	h->tab_nodraw = rows && compact && hrows == 1 && !getenv("QVZ_FORCE_DRAWS");
	h->tab_dmode = rows ? h->tab_dm : h->tab_toeplitz;
	return QVZ_OK;
}

// Tables -> device: the full 72 x 72 composition (W, R), the distortion mode, and -- when they fit -- the per-column
// images of the batched walk (quantize.cu).  Synchronises once (the image geometry depends on the tables' content).
extern "C" int qvz_gpu_upload_tables(qvz_gpu *h, const struct qvz_flat_tables *t) {
	if (!h || !t) return QVZ_ERR_ARG;
	if (!h->Xw) QVZ_FAIL(h, QVZ_ERR_ARG, "upload_tables: no rows loaded");
	if (t->columns != h->L.C || t->clusters == 0 || t->clusters > QVZ_MAX_CLUSTERS)
		QVZ_FAIL(h, QVZ_ERR_ARG, "upload_tables: tables are for %u clusters x %u columns, the rows have %u columns", t->clusters, t->columns, h->L.C);
	QVZ_CUDA(h, cudaSetDevice(h->device));
	h->tab_valid = 0;
	if (h->walk_recorded) QVZ_CUDA(h, cudaEventSynchronize(h->ev_walk_done));       // a previous walk may still read the images
	int toeplitz = 0;
	int rc = upload_tables(h, t, &toeplitz);
	if (rc) return rc;
	const uint32_t K = t->clusters, C = t->columns, KC = K * C;
	// distortion mode of the batched walk: 3 = (x-y)^2 (-d M), 4 = |x-y| (-d A), 2 = other integers of |x-y|, 1 = doubles of |x-y| (-d L), 0 = any matrix
	int dm = toeplitz;
	if (toeplitz == 2) {
		bool sq = true, ab = true;
		for (uint32_t d = 0; d < 72; ++d) {
			if (t->distortion[d] != (double) (d * d)) sq = false;
			if (t->distortion[d] != (double) d) ab = false;
		}
		dm = sq ? 3 : ab ? 4 : 2;
	}
	// A-1 = max(largest symbol in the rows, largest quantized value a present context can emit for such a symbol)
	QVZ_CUDA(h, cudaMemsetAsync(h->flags + 4, 0, sizeof(int), h->stream));
	QVZ_CUDA(h, cudaMemsetAsync(h->flags + 7, 0, sizeof(int), h->stream));
	rc = qvz_quantize_vmax(h, KC, h->smax);
	if (rc) return rc;
	rc = ensure_buf(h, &h->rowmap, &h->rowmap_cap, (size_t) KC * 72);
	if (!rc) rc = ensure_buf(h, &h->reach, &h->reach_cap, (size_t) KC * 72);
	if (!rc) rc = ensure_buf(h, &h->start, &h->start_cap, (size_t) K * sizeof(uint32_t));
	if (rc) return rc;
	QVZ_CUDA(h, cudaMemcpyAsync(h->h_flags, h->flags, QVZ_NFLAGS * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
	QVZ_CUDA(h, cudaStreamSynchronize(h->stream));
	if (h->h_flags[3]) {
		h->h_flags[3] = 0;
		QVZ_CUDA(h, cudaMemcpyAsync(h->flags + 3, h->h_flags + 3, sizeof(int), cudaMemcpyHostToDevice, h->stream));
		QVZ_FAIL(h, QVZ_ERR_ARG, "quantize: malformed flat tables (context index or quantized value out of range)");
	}
	const uint32_t vmax = (uint32_t) h->h_flags[4] > h->smax ? (uint32_t) h->h_flags[4] : h->smax;
	h->tab_K = K;
	h->tab_C = C;
	h->tab_box = (vmax + 2) & ~1u;               // even, >= vmax + 1
	h->tab_toeplitz = toeplitz;
	h->tab_dm = dm;
	rc = build_images(h, true);
	if (rc) return rc;
	h->tab_valid = 1;
	return QVZ_OK;
}

extern "C" int qvz_gpu_quantize(qvz_gpu *h, const struct qvz_flat_tables *t, const uint32_t well_seed[32],
                                uint8_t *symbols_out, uint8_t *qv_out, double *line_err_out)
{
	if (!h || !well_seed) return QVZ_ERR_ARG;
	if (!h->Xw || !h->K) QVZ_FAIL(h, QVZ_ERR_ARG, "quantize: rows and cluster ids must be resident first");
	if (!t && !h->tab_valid) QVZ_FAIL(h, QVZ_ERR_ARG, "quantize: no tables given and none uploaded (qvz_gpu_upload_tables)");
	QVZ_CUDA(h, cudaSetDevice(h->device));
	const qvz_layout &L = h->L;
	const size_t wbytes = (size_t) L.C4 * L.P * sizeof(uint32_t);
	int rc = ensure_buf(h, &h->Yw, &h->Yw_cap, wbytes);
	if (!rc && qv_out) rc = ensure_buf(h, &h->Qw, &h->Qw_cap, wbytes);
	if (!rc) rc = ensure_buf(h, &h->Ep, &h->Ep_cap, (size_t) L.P * sizeof(double));   // the walk always sums the distortion, like the reference
	if (rc) return rc;

	QVZ_CUDA(h, cudaEventRecord(h->ev[EV_A], h->stream));
	// draws first (aux stream), unless a prefetch for this seed is already in flight or done
	const bool prefetched = h->draws_state && memcmp(h->draws_seed, well_seed, sizeof(h->draws_seed)) == 0;
	if (!prefetched && (t || !(h->tab_A && h->tab_nodraw))) {     // (resident tables whose walk needs no draws: nothing to generate)
		rc = start_draws(h, well_seed, t ? batched_possible(h) : h->tab_A != 0);
		if (rc) return rc;
	}
	if (t) {
		rc = qvz_gpu_upload_tables(h, t);            // (synchronises the main stream; the draws keep running on theirs)
		if (rc) return rc;
	}
	if (h->tab_C != L.C || h->tab_K < h->K)
		QVZ_FAIL(h, QVZ_ERR_ARG, "quantize: tables are for %u clusters x %u columns, data has %u x %u", h->tab_K, h->tab_C, h->K, L.C);
	bool retried = false, batched = false, use_draws = true;
walk_again:
	batched = h->tab_A != 0;
	use_draws = !(batched && h->tab_nodraw);         // no context mixes its quantizers: the walk reads no draws (quantize.cu)
	if (use_draws && batched && h->draws_state != 2) {   // the guess said "line-major" but the tables allow the batched walk
		rc = start_draws(h, well_seed, true);
		if (rc) return rc;
	}
	QVZ_CUDA(h, cudaEventRecord(h->ev[EV_B], h->stream));
	if (use_draws) QVZ_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_draws, 0));      // run states (+ draws) are ready
	qvz_well_debug(h, "after run_states");
	QVZ_CUDA(h, cudaEventRecord(h->ev[EV_F], h->stream));
	if (batched) rc = qvz_quantize_launch_batched(h, h->tab_rows, h->tab_hrows, h->tab_A, qv_out != nullptr, h->tab_dmode);
	else rc = qvz_quantize_launch(h, qv_out != nullptr, 1, h->tab_dmode);
	if (rc) return rc;
	QVZ_CUDA(h, cudaEventRecord(h->ev[EV_C], h->stream));
	QVZ_CUDA(h, cudaEventRecord(h->ev_walk_done, h->stream));
	h->walk_recorded = 1;
	h->draws_state = 0;                          // one generation serves one walk (a new call may bring a new seed)

	// egress: packed -> line-major pieces on the device, each copied to the host while the next is re-laid out
	if (symbols_out) {
		rc = pipeline_d2h(h, symbols_out, L.C, L.C, [&](const piece &pc) {
			return qvz_layout_words_to_lines(h, pc.r0, pc.nr, h->Yw, h->stage[pc.buf], L.C, 0);
		});
		if (rc) return rc;
	}
	if (qv_out) {
		rc = pipeline_d2h(h, qv_out, L.C + 1, L.C + 1, [&](const piece &pc) {
			return qvz_layout_words_to_lines(h, pc.r0, pc.nr, h->Qw, h->stage[pc.buf], L.C + 1, 1);
		});
		if (rc) return rc;
	}
	if (line_err_out) {
		rc = pipeline_d2h(h, (uint8_t *) line_err_out, sizeof(double), sizeof(double), [&](const piece &pc) {
			return qvz_layout_doubles_to_lines(h, pc.r0, pc.nr, h->Ep, (double *) h->stage[pc.buf]);
		});
		if (rc) return rc;
	}
	qvz_well_debug(h, "end of quantize");
	QVZ_CUDA(h, cudaEventRecord(h->ev[EV_D], h->stream));
	int missing = 0;
	rc = take_flag(h, 2, &missing);              // the one host synchronisation of the stage
	if (rc) return rc;
	if (missing && batched && h->tab_support_used && !retried) {
		// The images were built for the values the counting stage saw under the cluster ids of THAT time; if the ids have
		// changed since, a line may need a row that was left out (it then ends in the poison row: never a wrong symbol).
		// Rebuild the images without that assumption and walk again.
		rc = build_images(h, false);
		if (rc) return rc;
		rc = start_draws(h, well_seed, h->tab_A != 0);
		if (rc) return rc;
		retried = true;
		goto walk_again;
	}
	// setup = everything on the main stream before the walk (table upload/composition when tables came with the call, waiting
	// for the draws); draws = draw generator kernel on the aux stream; quantize = draws + walk kernel durations
	float draws_ms = 0.f;
	if (use_draws && batched) cudaEventElapsedTime(&draws_ms, h->ev_jump_done, h->ev_draws);      // the draw generator kernel alone (jump-ahead is setup)
	h->tm.quantize_setup_ms = ev_ms(h, EV_A, EV_F);
	h->tm.quantize_draws_ms = draws_ms;
	h->tm.quantize_ms = h->tm.quantize_draws_ms + ev_ms(h, EV_F, EV_C);
	h->tm.quantize_d2h_ms = ev_ms(h, EV_C, EV_D);
	if (missing) QVZ_FAIL(h, QVZ_ERR_CONTEXT, "quantize: reached a context without a quantizer (the reference asserts, src/codebook.c:164)");
	return QVZ_OK;
}

extern "C" int qvz_gpu_well_jump(qvz_gpu *h, const uint32_t seed[32], uint64_t words, uint32_t state_out[32]) {
	if (!h || !seed || !state_out) return QVZ_ERR_ARG;
	QVZ_CUDA(h, cudaSetDevice(h->device));
	uint32_t *dev = nullptr;
	QVZ_CUDA(h, cudaMalloc(&dev, 32 * sizeof(uint32_t)));
	int rc = qvz_well_jump_state(h, seed, words, dev);
	cudaError_t e = cudaSuccess;
	if (!rc) e = cudaMemcpyAsync(state_out, dev, 32 * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream);
	if (!rc && e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
	cudaFree(dev);
	if (rc) return rc;
	if (e != cudaSuccess) QVZ_FAIL(h, QVZ_ERR_CUDA, "well_jump: %s", cudaGetErrorString(e));
	return QVZ_OK;
}
