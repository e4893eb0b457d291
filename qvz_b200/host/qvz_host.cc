// qvz_host.cc -- host side of the encoder behind include/qvz_host.h: codebook design from the GPU's conditional
// counts, the flat quantizer tables for the GPU walk, the codebook text and the adaptive arithmetic coder that
// consumes the GPU's symbol stream.
//
// Every double below is produced by the same operations in the same order as in the reference (build with
// -ffp-contract=off: the reference's x86-64 build has no fused multiply-add), so ratios, tables and the final
// bytes are identical.  Reference lines are cited per function (paths relative to the reference tree).
//
// One deliberate restructuring: the reference recomputes the inner sum of compute_qpmf_list
// (src/codebook.c:318-321) for every output symbol although it does not depend on it -- a 72^4 loop per column
// that is ~all of its 30-50 s codebook time.  Here that sum is computed once per (k, j) and reused; the value
// added is the identical double, so the result is bit-identical at 1/72 of the work.
#include "../../include/qvz_host.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace {

constexpr uint32_t A = QVZ_ALPHABET;            // 72 symbols (ALPHABET_SIZE, src/main.c:13)
constexpr uint32_t NOT_FOUND = 0xFFFFFFFFu;     // ALPHABET_SYMBOL_NOT_FOUND (include/pmf.h:9)
constexpr uint32_t MAX_ITER = 100;              // QUANTIZER_MAX_ITER (include/quantizer.h:10)

// struct alphabet_t (include/pmf.h:20-24): symbol list + 72-entry reverse index
struct Alphabet {
	std::vector<uint8_t> sym;
	uint32_t idx[A];
	void reindex() {                             // alphabet_compute_index (src/pmf.c:365-382): the LAST position of a repeated symbol wins
		for (uint32_t i = 0; i < A; ++i) idx[i] = NOT_FOUND;
		for (uint32_t i = 0; i < sym.size(); ++i) idx[sym[i]] = i;
	}
};

// alphabet_union (src/pmf.c:312-357): the reference's merge, kept verbatim in behaviour even for unsorted input
Alphabet alphabet_union(const Alphabet &a, const Alphabet &b) {
	Alphabet r;
	size_t i = 0, j = 0;
	while (i < a.sym.size() && j < b.sym.size()) {
		if (a.sym[i] < b.sym[j]) r.sym.push_back(a.sym[i++]);
		else if (a.sym[i] == b.sym[j]) { r.sym.push_back(a.sym[i]); ++i; ++j; }
		else r.sym.push_back(b.sym[j++]);
	}
	while (i < a.sym.size()) r.sym.push_back(a.sym[i++]);
	while (j < b.sym.size()) r.sym.push_back(b.sym[j++]);
	r.reindex();
	return r;
}

// struct quantizer_t (include/quantizer.h:16-22)
struct Quantizer {
	uint8_t q[A];
	Alphabet out;
	double ratio = 0.0, mse = 0.0;
};

inline double dist_at(const double *D, uint32_t x, uint32_t y) { return D[x + A * y]; }   // get_distortion (src/distortion.c:151-153)

// generate_quantizer (src/quantizer.c:34-132): Lloyd-Max over the 72-symbol alphabet for `states` regions
Quantizer generate_quantizer(const double *p, const double *D, uint32_t states) {
	Quantizer Q;
	memset(Q.q, 0, sizeof(Q.q));
	uint8_t bounds[A + 2], rec[A + 1];
	bounds[0] = 0;
	bounds[states] = (uint8_t) A;
	for (uint32_t j = 1; j < states; ++j) bounds[j] = (uint8_t) ((j * A) / states);
	for (uint32_t j = 0; j < states; ++j) rec[j] = (uint8_t) ((bounds[j] + bounds[j + 1] - 1) / 2);

	uint32_t changed = 1, iter = 0;
	while (changed && iter < MAX_ITER) {
		changed = 0;
		iter += 1;
		for (uint32_t j = 0; j < states; ++j) {       // reconstruction points for fixed bounds (:58-85)
			double min_mse = DBL_MAX;
			uint32_t min_r = bounds[j];
			for (uint32_t r = bounds[j]; r < bounds[j + 1]; ++r) {
				double mse = 0.0;
				for (uint32_t i = bounds[j]; i < bounds[j + 1]; ++i) mse += p[i] * dist_at(D, i, r);
				if (mse < min_mse) {
					min_r = r;
					min_mse = mse;
				}
			}
			if (min_r != rec[j]) {
				changed = 1;
				rec[j] = (uint8_t) min_r;
			}
		}
		uint32_t r = 0;                               // bounds for fixed reconstruction points (:91-104); bounds past r keep their old value
		for (uint32_t j = 1; j < A - 1 && r < states - 1; ++j) {
			const double mse = dist_at(D, j, rec[r]), next_mse = dist_at(D, j, rec[r + 1]);
			if (next_mse < mse) {
				r += 1;
				bounds[r] = (uint8_t) j;
			}
		}
	}
	for (uint32_t j = 0; j < states; ++j)             // input -> reconstruction point (:109-113)
		for (uint32_t i = bounds[j]; i < bounds[j + 1]; ++i) Q.q[i] = rec[j];
	Q.out.sym.assign(rec, rec + states);              // output alphabet = the reconstruction points (:116-118)
	Q.out.reindex();
	Q.mse = 0.0;                                      // expected distortion (:121-126)
	for (uint32_t j = 0; j < states; ++j)
		for (uint32_t i = bounds[j]; i < bounds[j + 1]; ++i) Q.mse += dist_at(D, i, rec[j]) * p[i];
	return Q;
}

// get_entropy (src/pmf.c:141-155) of apply_quantizer(q, p) (src/quantizer.c:139-161)
double quantized_entropy(const Quantizer &Q, const double *p) {
	double out[A];
	for (uint32_t i = 0; i < A; ++i) out[i] = 0.0;
	for (uint32_t i = 0; i < A; ++i) out[Q.q[i]] += p[i];
	double e = 0.0;
	for (uint32_t i = 0; i < A; ++i)
		if (out[i] > 0.0) e -= out[i] * log2(out[i]);
	return e;
}

double entropy(const double *p) {
	double e = 0.0;
	for (uint32_t i = 0; i < A; ++i)
		if (p[i] > 0.0) e -= p[i] * log2(p[i]);
	return e;
}

// optimize_for_entropy (src/codebook.c:230-269): the pair of quantizers around the target entropy + mixing ratio
double optimize_for_entropy(const double *p, const double *D, double target, Quantizer &lo, Quantizer &hi) {
	if (target == 0.0) {
		lo = generate_quantizer(p, D, 1);
		hi = generate_quantizer(p, D, 1);
		return 1.0;
	}
	uint32_t states = 1;
	hi = generate_quantizer(p, D, states);
	double hi_entropy = quantized_entropy(hi, p), lo_entropy;
	do {
		lo = std::move(hi);
		lo_entropy = hi_entropy;
		states += 1;
		hi = generate_quantizer(p, D, states);
		hi_entropy = quantized_entropy(hi, p);
	} while (hi_entropy < target && states < A);
	if (hi_entropy < target) return 0.0;
	if (lo_entropy >= target || hi_entropy == lo_entropy) return 1.0;
	return (target - hi_entropy) / (lo_entropy - hi_entropy);
}

// renormalize_pmf (src/pmf.c:235-254)
void renormalize(double *p, size_t n) {
	double total = 0;
	for (size_t i = 0; i < n; ++i) total += p[i];
	if (total > 0)
		for (size_t i = 0; i < n; ++i) p[i] = p[i] / total;
}

// one cluster's struct cond_quantizer_list_t (include/codebook.h:61-69)
struct ClusterBook {
	std::vector<Alphabet> in;                        // input_alphabets[col]
	std::vector<std::vector<Quantizer>> q;           // q[col][2*ctx + hi]
	std::vector<std::vector<uint8_t>> qratio;        // qratio[col][ctx]
};

// A few persistent worker threads for the loops of design_cluster whose iterations are independent (every double is still
// computed by one thread with the reference's own order of operations, so the results do not depend on the thread count).
class ForkJoin {
	std::vector<std::thread> th;
	std::mutex mu;
	std::condition_variable cv, done_cv;
	const std::function<void(size_t)> *job = nullptr;
	size_t n = 0;
	std::atomic<size_t> next{0};
	size_t pending = 0;
	uint64_t gen = 0;
	bool stop = false;
	void work() {
		for (;;) {
			const size_t i = next.fetch_add(1);
			if (i >= n) break;
			(*job)(i);
		}
	}
	void worker() {
		uint64_t seen = 0;
		for (;;) {
			{
				std::unique_lock<std::mutex> lk(mu);
				cv.wait(lk, [&] { return stop || gen != seen; });
				if (stop) return;
				seen = gen;
			}
			work();
			{
				std::lock_guard<std::mutex> lk(mu);
				if (--pending == 0) done_cv.notify_one();
			}
		}
	}
public:
	explicit ForkJoin(unsigned threads) {
		for (unsigned t = 1; t < threads; ++t) th.emplace_back([this] { worker(); });
	}
	~ForkJoin() {
		{
			std::lock_guard<std::mutex> lk(mu);
			stop = true;
		}
		cv.notify_all();
		for (auto &t : th) t.join();
	}
	void run(size_t count, const std::function<void(size_t)> &fn) {
		if (th.empty() || count < 2) {
			for (size_t i = 0; i < count; ++i) fn(i);
			return;
		}
		{
			std::lock_guard<std::mutex> lk(mu);
			job = &fn;
			n = count;
			next.store(0);
			pending = th.size();
			++gen;
		}
		cv.notify_all();
		work();
		std::unique_lock<std::mutex> lk(mu);
		done_cv.wait(lk, [&] { return pending == 0; });
	}
};

// the marginals of calculate_statistics (src/codebook.c:208-219) + generate_codebooks (:355-468) for one cluster
void design_cluster(const uint32_t *counts, uint32_t C, int mode, double target, const double *D, ClusterBook &B, unsigned inner_threads) {
	ForkJoin pool(inner_threads);
	const size_t rows = 1 + (size_t) A * (C - 1);
	// pmf_t.pmf of every conditional pmf: counts / total, all zero when the row was never seen (recalculate_pmf, src/pmf.c:219-230)
	std::vector<double> condp(rows * A, 0.0);
	for (size_t r = 0; r < rows; ++r) {
		uint32_t total = 0;
		for (uint32_t i = 0; i < A; ++i) total += counts[r * A + i];      // pmf_t.total (uint32, src/pmf.c:213)
		if (!total) continue;
		const double t = (double) total;
		for (uint32_t i = 0; i < A; ++i) condp[r * A + i] = ((double) counts[r * A + i]) / t;
	}
	auto cond = [&](uint32_t col, uint32_t prev) -> const double * {     // get_cond_pmf (src/codebook.c:116-120)
		return &condp[(col == 0 ? 0 : 1 + (size_t) (col - 1) * A + prev) * A];
	};
	// marginal_pmfs: combine_pmfs (src/pmf.c:189-205) chained over the previous column's values
	std::vector<double> marg((size_t) C * A, 0.0);
	for (uint32_t i = 0; i < A; ++i) marg[i] = 1.0 * cond(0, 0)[i] + 0.0 * marg[i];
	for (uint32_t col = 1; col < C; ++col)
		for (uint32_t j = 0; j < A; ++j) {
			const double w = marg[(size_t) (col - 1) * A + j];
			const double *b = cond(col, j);
			double *m = &marg[(size_t) col * A];
			for (uint32_t i = 0; i < A; ++i) m[i] = 1.0 * m[i] + w * b[i];
		}

	B.in.resize(C);
	B.q.resize(C);
	B.qratio.resize(C);
	auto store = [&](uint32_t col, uint32_t ctx, Quantizer &&lo, Quantizer &&hi, double ratio) {    // store_cond_quantizers_indexed (:152-157)
		lo.ratio = ratio;
		hi.ratio = 1 - ratio;
		B.q[col][2 * ctx] = std::move(lo);
		B.q[col][2 * ctx + 1] = std::move(hi);
		B.qratio[col][ctx] = (uint8_t) (ratio * 128.);
	};
	auto entropy_target = [&](const double *p) { return mode == QVZ_MODE_RATIO ? entropy(p) * target : target; };

	// column 0: the single context {0} (:383-398)
	B.in[0].sym.assign(1, 0);
	B.in[0].reindex();
	B.q[0].resize(2);
	B.qratio[0].resize(1);
	double ratio0;
	{
		Quantizer lo, hi;
		ratio0 = optimize_for_entropy(cond(0, 0), D, entropy_target(cond(0, 0)), lo, hi);
		store(0, 0, std::move(lo), std::move(hi), ratio0);
	}

	std::vector<double> prev_qpmf, qpmf, xpmf, ptemp;    // qpmf[k][idx] = P(Q_{c-1} = U[idx] | X_{c-1} = k);  xpmf[idx][k] = P(X_c = k | Q_{c-1} = U[idx])
	for (uint32_t col = 1; col < C; ++col) {
		// contexts of this column = union of the output alphabets of the previous column's quantizers (:411-416)
		const size_t nprev = B.in[col - 1].sym.size();
		Alphabet U = B.q[col - 1][0].out;
		for (size_t j = 1; j < 2 * nprev; ++j) U = alphabet_union(U, B.q[col - 1][j].out);
		const size_t nu = U.sym.size();
		B.in[col] = U;
		B.q[col].resize(2 * nu);
		B.qratio[col].resize(nu);

		qpmf.assign((size_t) A * nu, 0.0);
		if (col == 1) {                                  // compute_qpmf_quan_list (:274-289)
			const Quantizer &lo = B.q[0][0], &hi = B.q[0][1];
			for (uint32_t x = 0; x < A; ++x)
				for (size_t idx = 0; idx < nu; ++idx) {
					const uint32_t s = U.sym[idx];
					if (lo.q[x] == s) qpmf[x * nu + idx] += ratio0;
					if (hi.q[x] == s) qpmf[x * nu + idx] += (1 - ratio0);
				}
		} else {                                         // compute_qpmf_list (:291-330)
			// p_temp(k, j) = sum_x P(Q_{c-2}=j | X_{c-2}=x) * P(X_{c-1}=k | X_{c-2}=x) * P(X_{c-2}=x): independent of idx
			ptemp.assign((size_t) A * nprev, 0.0);
			pool.run(A, [&](size_t k) {
				for (size_t j = 0; j < nprev; ++j) {
					double t = 0;
					for (uint32_t x = 0; x < A; ++x)
						t += prev_qpmf[x * nprev + j] * cond(col - 1, x)[k] * marg[(size_t) (col - 2) * A + x];
					ptemp[k * nprev + j] = t;
				}
			});
			pool.run(A, [&](size_t k) {
				for (size_t idx = 0; idx < nu; ++idx) {
					const uint32_t s = U.sym[idx];
					double acc = 0.0;
					for (size_t j = 0; j < nprev; ++j) {
						const Quantizer &lo = B.q[col - 1][2 * j], &hi = B.q[col - 1][2 * j + 1];
						double p_q_xq = 0.0;
						if (lo.q[k] == s) p_q_xq += lo.ratio;
						if (hi.q[k] == s) p_q_xq += hi.ratio;
						acc += p_q_xq * ptemp[k * nprev + j];
					}
					qpmf[k * nu + idx] = acc;
				}
				renormalize(&qpmf[k * nu], nu);
			});
		}

		xpmf.assign(nu * A, 0.0);                        // compute_xpmf_list (:332-349)
		pool.run(nu, [&](size_t idx) {
			for (uint32_t k = 0; k < A; ++k) {
				double t = 0.0;
				for (uint32_t x = 0; x < A; ++x)
					t += qpmf[x * nu + idx] * cond(col, x)[k] * marg[(size_t) (col - 1) * A + x];
				xpmf[idx * A + k] = t;
			}
			renormalize(&xpmf[idx * A], A);
		});

		pool.run(nu, [&](size_t j) {                     // one quantizer pair per context (:432-446), contexts are independent
			Quantizer lo, hi;
			const double *p = &xpmf[j * A];
			const double ratio = optimize_for_entropy(p, D, entropy_target(p), lo, hi);
			store(col, (uint32_t) j, std::move(lo), std::move(hi), ratio);
		});
		prev_qpmf.swap(qpmf);
	}
}

}  // namespace

struct qvz_codebooks {
	uint32_t K = 0, C = 0;
	std::vector<ClusterBook> books;
	std::vector<double> distortion;
	// flat mirror (struct qvz_flat_tables)
	std::vector<uint32_t> nctx;
	std::vector<uint8_t> ctx_of, qratio, qmap, smap;
	std::vector<uint64_t> q_off;
};

extern "C" int qvz_host_distortion(int type, double *out) {
	// gen_manhattan_distortion / gen_mse_distortion / gen_lorentzian_distortion (src/distortion.c:50-93)
	for (uint32_t x = 0; x < A; ++x)
		for (uint32_t y = 0; y < A; ++y) {
			const int d = abs((int) x - (int) y);
			double v;
			if (type == QVZ_DIST_MANHATTAN) v = d;
			else if (type == QVZ_DIST_MSE) v = ((int) x - (int) y) * ((int) x - (int) y);
			else if (type == QVZ_DIST_LORENTZ) v = log2(1.0 + (double) d);
			else return -1;
			out[x + y * A] = v;
		}
	return 0;
}

extern "C" int qvz_host_distortion_file(const char *path, double *out) {
	// gen_custom_distortion (src/distortion.c:100-145): rows of comma separated doubles, '#' lines are comments.
	// (The reference spins forever on a short row -- its fill loop never advances; short rows are zero-filled here.)
	FILE *fp = fopen(path, "rt");
	if (!fp) return -1;
	for (uint32_t i = 0; i < A * A; ++i) out[i] = 0.0;
	char line[1024];
	uint32_t x = 0;
	while (x < A && fgets(line, sizeof(line), fp) != NULL) {
		if (line[0] == '#') continue;
		char *field = line - 1;
		uint32_t y = 0;
		while (y < A && field != NULL) {
			field += 1;
			out[x + A * y] = atof(field);
			field = strchr(field, ',');
			y += 1;
		}
		x += 1;
	}
	fclose(fp);
	return 0;
}

extern "C" qvz_codebooks *qvz_host_design(const uint32_t *counts, uint32_t clusters, uint32_t columns, int mode, double target,
                                          const double *distortion, int threads) {
	if (!counts || !distortion || clusters == 0 || clusters > 255 || columns == 0 || columns > QVZ_MAX_COLUMNS) return nullptr;
	if (mode != QVZ_MODE_RATIO && mode != QVZ_MODE_FIXED) return nullptr;
	qvz_codebooks *cb = new qvz_codebooks();
	cb->K = clusters;
	cb->C = columns;
	cb->distortion.assign(distortion, distortion + A * A);
	cb->books.resize(clusters);
	const size_t per_cluster = (1 + (size_t) A * (columns - 1)) * A;
	// clusters are independent (one thread each); inside a cluster the columns are sequential but the contexts of a column
	// are not: the threads that are left over (<= 8 per cluster) share them
	const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
	unsigned nt = threads > 0 ? (unsigned) threads : std::min<unsigned>(clusters, hw);
	nt = std::min<unsigned>(nt, clusters);
	unsigned inner = threads > 0 ? std::max(1u, (unsigned) threads / nt) : std::max(1u, hw / nt);
	if (inner > 8) inner = 8;
	std::vector<std::thread> pool;
	for (unsigned t = 0; t < nt; ++t)
		pool.emplace_back([&, t]() {
			for (uint32_t k = t; k < clusters; k += nt)
				design_cluster(counts + k * per_cluster, columns, mode, target, cb->distortion.data(), cb->books[k], inner);
		});
	for (auto &th : pool) th.join();

	// flatten: contexts, ratios, input->value maps and value->state maps, always copied
	const size_t KC = (size_t) clusters * columns;
	cb->nctx.resize(KC);
	cb->q_off.resize(KC);
	cb->ctx_of.assign(KC * A, QVZ_CTX_ABSENT);
	uint64_t nq = 0;
	for (uint32_t k = 0; k < clusters; ++k)
		for (uint32_t c = 0; c < columns; ++c) {
			const size_t kc = (size_t) k * columns + c;
			const ClusterBook &B = cb->books[k];
			cb->nctx[kc] = (uint32_t) B.in[c].sym.size();
			cb->q_off[kc] = nq;
			nq += 2 * B.in[c].sym.size();
			for (uint32_t v = 0; v < A; ++v)
				if (B.in[c].idx[v] != NOT_FOUND) cb->ctx_of[kc * A + v] = (uint8_t) B.in[c].idx[v];
		}
	cb->qratio.resize(nq / 2);
	cb->qmap.resize(nq * A);
	cb->smap.assign(nq * A, 0xFF);
	for (uint32_t k = 0; k < clusters; ++k)
		for (uint32_t c = 0; c < columns; ++c) {
			const size_t kc = (size_t) k * columns + c;
			const ClusterBook &B = cb->books[k];
			for (size_t j = 0; j < B.q[c].size(); ++j) {
				const uint64_t qi = cb->q_off[kc] + j;
				memcpy(&cb->qmap[qi * A], B.q[c][j].q, A);
				for (uint32_t v = 0; v < A; ++v)
					if (B.q[c][j].out.idx[v] != NOT_FOUND) cb->smap[qi * A + v] = (uint8_t) B.q[c][j].out.idx[v];
			}
			for (size_t j = 0; j < B.qratio[c].size(); ++j) cb->qratio[cb->q_off[kc] / 2 + j] = B.qratio[c][j];
		}
	return cb;
}

extern "C" void qvz_host_free(qvz_codebooks *cb) { delete cb; }

extern "C" int qvz_host_tables(const qvz_codebooks *cb, struct qvz_flat_tables *out) {
	if (!cb || !out) return -1;
	out->clusters = cb->K;
	out->columns = cb->C;
	out->nctx = cb->nctx.data();
	out->ctx_of = cb->ctx_of.data();
	out->q_off = cb->q_off.data();
	out->qratio = cb->qratio.data();
	out->qmap = cb->qmap.data();
	out->smap = cb->smap.data();
	out->distortion = cb->distortion.data();
	return 0;
}

// write_codebooks / write_codebook (src/codebook.c:474-555)
extern "C" uint64_t qvz_host_codebook_bytes(const qvz_codebooks *cb) {
	if (!cb) return 0;
	uint64_t n = 9;                                  // cluster count, columns, lines
	for (const ClusterBook &B : cb->books) {
		n += 2 + 2 * (A + 1);
		for (uint32_t c = 1; c < cb->C; ++c) {
			const uint64_t nc = B.in[c].sym.size();
			n += (nc + 1) + 2 * (nc * A + 1);
		}
	}
	return n;
}

extern "C" int qvz_host_write_codebooks(const qvz_codebooks *cb, uint64_t n_lines, uint8_t *out) {
	if (!cb || !out) return -1;
	uint8_t *p = out;
	*p++ = (uint8_t) cb->K;
	const uint32_t cols = cb->C, lines = (uint32_t) n_lines;    // htonl: big endian, lines truncated to 32 bits (:481-482)
	for (int s = 24; s >= 0; s -= 8) *p++ = (uint8_t) (cols >> s);
	for (int s = 24; s >= 0; s -= 8) *p++ = (uint8_t) (lines >> s);
	auto put_q = [&](const Quantizer &Q) {
		for (uint32_t i = 0; i < A; ++i) *p++ = (uint8_t) (Q.q[i] + 33);     // COPY_Q_TO_LINE (include/codebook.h:101)
	};
	for (const ClusterBook &B : cb->books) {
		*p++ = (uint8_t) (B.qratio[0][0] + 33);
		*p++ = '\n';
		put_q(B.q[0][0]);
		*p++ = '\n';
		put_q(B.q[0][1]);
		*p++ = '\n';
		for (uint32_t c = 1; c < cb->C; ++c) {
			const size_t nc = B.in[c].sym.size();
			for (size_t j = 0; j < nc; ++j) *p++ = (uint8_t) (B.qratio[c][j] + 33);
			*p++ = '\n';
			for (size_t j = 0; j < nc; ++j) put_q(B.q[c][2 * j]);
			*p++ = '\n';
			for (size_t j = 0; j < nc; ++j) put_q(B.q[c][2 * j + 1]);
			*p++ = '\n';
		}
	}
	return (uint64_t) (p - out) == qvz_host_codebook_bytes(cb) ? 0 : -1;
}

// ------------------------------------------------------------------------------------------------------ coder
namespace {

// struct os_stream_t + stream_write_bit / stream_finish_byte / stream_write_buffer (src/os_stream.c:72-120), MSB first
struct BitWriter {
	FILE *fp;
	std::vector<uint8_t> buf;
	size_t pos = 0, limit;
	uint64_t acc = 0;                                // bits not yet written, right aligned (always < 8 of them between calls)
	uint32_t nacc = 0;
	uint64_t written = 0;
	bool ok = true;
	explicit BitWriter(FILE *f) : fp(f), buf((size_t) 1 << 22), limit(buf.size() - 16) {}
	void flush() {
		if (pos && fwrite(buf.data(), 1, pos, fp) != pos) ok = false;
		written += pos;
		pos = 0;
	}
	// n <= 32 bits of v, most significant first (stream_write_bits, one bit at a time there).  The whole bytes among the
	// pending bits are stored with one unconditional 8-byte big-endian store; only the position advances by their count.
	inline void bits(uint32_t v, uint32_t n) {
		acc = (acc << n) | v;
		nacc += n;                                   // <= 7 + 32
		if (nacc == 0) return;                       // nothing pending (n == 0 on a byte boundary): a shift by 64 is undefined
		const uint64_t be = __builtin_bswap64(acc << (64 - nacc));
		memcpy(buf.data() + pos, &be, 8);
		pos += nacc >> 3;
		nacc &= 7;
		acc &= (1ull << nacc) - 1;
		if (pos >= limit) flush();
	}
	inline void bit(uint32_t b) { bits(b & 1, 1); }
	inline void run(uint32_t b, uint32_t n) {        // n copies of bit b
		const uint32_t ones = b ? 0xFFFFFFFFu : 0u;
		for (; n >= 32; n -= 32) bits(ones, 32);
		if (n) bits(ones >> (32 - n), n);
	}
	void finish() {                                  // stream_finish_byte: pads the byte in progress; on a byte boundary it still emits one zero byte
		buf[pos++] = (uint8_t) (acc << (8 - nacc));
		acc = 0;
		nacc = 0;
		flush();
	}
};

// struct stream_stats_t + update_stats (src/qv_stream.c:9-25)
struct Stats {
	uint32_t *counts;
	uint32_t card, n;
};

// Arithmetic_code_t + arithmetic_encoder_step (src/arith.c:5-97), m = 22 (include/qv_compressor.h:19)
struct Coder {
	static constexpr uint32_t M = 22, MSB = M - 1, SMSB = M - 2, CLEAR = (1u << MSB) - 1, R = 1u << (M - 3), STEP = 8;
	uint32_t l = 0, u = (1u << M) - 1;
	int32_t scale3 = 0;
	BitWriter &os;
	explicit Coder(BitWriter &w) : os(w) {}
	// cumulative counts of symbol x in model s, then update_stats(stats, x, a->r) (src/qv_stream.c:9-25).  This half of
	// arithmetic_encoder_step never looks at the interval, so it can run ahead of the coder (qvz_host_encode).
	static inline void model(Stats &s, uint32_t x, uint32_t &below, uint32_t &upto, uint32_t &n) {
		uint32_t b = 0;
		for (uint32_t i = 0; i < x; ++i) b += s.counts[i];
		below = b;
		upto = b + s.counts[x];
		n = s.n;
		s.counts[x] += STEP;
		s.n += STEP;
		if (s.n > R) {
			s.n = 0;
			for (uint32_t i = 0; i < s.card; ++i)
				if (s.counts[i]) {
					s.counts[i] >>= 1;
					s.counts[i] += 1;
					s.n += s.counts[i];
				}
		}
	}
	// the interval half (src/arith.c:38-96)
	inline void narrow(uint32_t below, uint32_t upto, uint32_t n) {
		// floor(range*cum / n) without the 64-bit integer divide (the longest latency on the coder's dependency chain):
		// range*cum < 2^43 and n < 2^21 are exact doubles; v = cum*range * fl(1/n) + 2^-25 is within 2^-29 of t + 2^-25 for
		// the true quotient t <= 2^22, a non-integer t is at least 1/n >= 2^-21 away from both neighbouring integers and an
		// integer t lands strictly inside (t, t+1): truncation gives floor(t) in every case.  1/n does not depend on the
		// interval, so the division itself is off the chain.
		const uint64_t range = (uint64_t) (u - l + 1);
		const double inv = 1.0 / (double) n;
		u = l + (uint32_t) ((double) (range * upto) * inv + 0x1p-25) - 1;
		l = l + (uint32_t) ((double) (range * below) * inv + 0x1p-25);
		// The reference rescales one bit per loop turn (src/arith.c:56-96).  Its turns come in two runs, each with a closed
		// form: first E1/E2 while the top bits of l and u agree -- k turns, k = the number of leading bits they share, which
		// are emitted together (the pending E3 complements follow the first of them, as in the reference) --, then E3 while
		// l = 01.. and u = 10.. -- j turns, j = how many bits after the top one have l = 1 and u = 0; each drops that bit.
		// After an E3 turn the top bits still differ, so no E1/E2 turn can follow: one pass, no loop, the same bits.
		const uint32_t MASK = (1u << M) - 1u;
		const uint32_t diff = l ^ u;
		const uint32_t k = diff ? (uint32_t) __builtin_clz(diff << (32 - M)) : M;
		if (scale3 > 0 && k) {
			const uint32_t top = l >> (M - k), first = top >> (k - 1);
			os.bits(first, 1);
			os.run(first ^ 1u, (uint32_t) scale3);
			scale3 = 0;
			os.bits(top & ((1u << (k - 1)) - 1u), k - 1);
		} else os.bits(l >> (M - k), k);                                     // k == 0: nothing (l < 2^M)
		l = (l << k) & MASK;
		u = ((u << k) & MASK) | ((1u << k) - 1u);
		// now l = 0..., u = 1...  (k == M, l == u before the shift, leaves l = 0, u = MASK)
		const uint32_t y = ((l & ~u) << 1) & MASK;                           // bit MSB-i set: i-th bit after the top has l = 1, u = 0
		const uint32_t j = (uint32_t) __builtin_clz(~(y << (32 - M)) | 1u);  // leading ones of the M-bit window (< M: bit 0 of y is 0)
		scale3 += (int32_t) j;
		l = (l << j) & CLEAR;
		u = ((u << j) & CLEAR) | (1u << MSB) | ((1u << j) - 1u);
	}
	inline void step(Stats &s, uint32_t x) {
		uint32_t below, upto, n;
		model(s, x, below, upto, n);
		narrow(below, upto, n);
	}
	void last() {                                    // encoder_last_step (src/arith.c:99-116)
		const uint32_t msbL = l >> MSB;
		os.bit(msbL);
		if (scale3 > 0) os.run(msbL ^ 1u, (uint32_t) scale3);
		scale3 = 0;
		os.bits(l & CLEAR, MSB);                     // stream_write_bits(os, a->l, m - 1): the low m-1 bits, most significant first
		os.finish();
	}
};

}  // namespace

extern "C" int qvz_host_encode(const qvz_codebooks *cb, const char *path, uint64_t n_lines, const uint8_t *cluster_ids,
                               const uint8_t *symbols, const uint32_t well_seed[32], uint64_t *stream_bytes_out) {
	if (!cb || !path || !cluster_ids || !symbols || !well_seed) return -1;
	FILE *fp = fopen(path, "wb");
	if (!fp) return -1;
	const uint32_t K = cb->K, C = cb->C;
	{
		std::vector<uint8_t> head(qvz_host_codebook_bytes(cb));
		qvz_host_write_codebooks(cb, n_lines, head.data());
		bool ok = fwrite(head.data(), 1, head.size(), fp) == head.size();
		ok = ok && fwrite(well_seed, sizeof(uint32_t), 32, fp) == 32;       // host-endian words, as the reference writes them (src/qv_stream.c:90)
		if (!ok) {
			fclose(fp);
			return -1;
		}
	}
	// initialize_stream_stats (src/qv_stream.c:32-61): one adaptive model per (cluster, column, quantizer), uniform start
	const size_t KC = (size_t) K * C;
	const uint64_t nq = cb->q_off[KC - 1] + 2ull * cb->nctx[KC - 1];
	std::vector<uint64_t> count_off(nq + 1, 0);
	std::vector<uint32_t> card(nq);
	for (uint32_t k = 0; k < K; ++k)
		for (uint32_t c = 0; c < C; ++c)
			for (size_t j = 0; j < cb->books[k].q[c].size(); ++j) card[cb->q_off[(size_t) k * C + c] + j] = (uint32_t) cb->books[k].q[c][j].out.sym.size();
	for (uint64_t i = 0; i < nq; ++i) count_off[i + 1] = count_off[i] + card[i];
	std::vector<uint32_t> counts(count_off[nq], 1u);
	std::vector<Stats> stats(nq);
	for (uint64_t i = 0; i < nq; ++i) stats[i] = Stats{&counts[count_off[i]], card[i], card[i]};
	std::vector<uint32_t> ccounts(K, 1u);                                   // cluster_stats (src/qv_stream.c:98-108)
	Stats cstats{ccounts.data(), K, K};
	// value reached by (quantizer, state): output_alphabet->symbols[state], needed for the next column's context
	std::vector<uint8_t> out_sym(count_off[nq]);
	for (uint32_t k = 0; k < K; ++k)
		for (uint32_t c = 0; c < C; ++c)
			for (size_t j = 0; j < cb->books[k].q[c].size(); ++j) {
				const uint64_t qi = cb->q_off[(size_t) k * C + c] + j;
				memcpy(&out_sym[count_off[qi]], cb->books[k].q[c][j].out.sym.data(), card[qi]);
			}

	BitWriter os(fp);
	Coder coder(os);
	int rc = 0;
	unsigned nthreads = std::thread::hardware_concurrency();
	if (nthreads > 8) nthreads = 8;
	if (n_lines * C < (1u << 20)) nthreads = 1;                             // small inputs: not worth a pipeline
	if (const char *e = getenv("QVZ_CODER_THREADS")) nthreads = (unsigned) atoi(e);     // explicit: any size (tests)
	if (nthreads <= 1) {
		for (uint64_t line = 0; line < n_lines && !rc; ++line) {            // src/qv_compressor.c:76-135 minus the quantization itself
			const uint32_t k = cluster_ids[line];
			if (k >= K) {
				rc = -2;
				break;
			}
			coder.step(cstats, k);                                          // qv_write_cluster (:86)
			const uint8_t *sym = symbols + line * C;
			uint32_t prev = 0;
			for (uint32_t c = 0; c < C; ++c) {
				const size_t kc = (size_t) k * C + c;
				const uint32_t ctx = cb->ctx_of[kc * A + prev];             // choose_quantizer's context lookup (src/codebook.c:163)
				const uint32_t hi = sym[c] >> 7, state = sym[c] & 0x7Fu;
				if (ctx == QVZ_CTX_ABSENT) {
					rc = -2;
					break;
				}
				const uint64_t qi = cb->q_off[kc] + 2 * ctx + hi;
				if (state >= card[qi]) {
					rc = -2;
					break;
				}
				coder.step(stats[qi], state);                               // compress_qv (:8-11)
				prev = out_sym[count_off[qi] + state];
			}
		}
	} else {
		// The adaptive models are one per (cluster, column, quantizer): what a symbol's model says (cumulative counts, total)
		// depends on the earlier symbols of the SAME column only, never on the coder's interval.  So blocks of lines go
		// through two parallel passes on worker threads -- (A) per line: which model codes each symbol (the context chain
		// along the line); (B) per column, lines in order: the model's answer, then its update -- and the main thread only
		// narrows the interval with the precomputed (below, upto, n) triples, in the reference's symbol order.  Same bits.
		uint64_t BL = 8192;                                                 // lines per block
		if (const char *e = getenv("QVZ_CODER_BLOCK")) BL = strtoull(e, nullptr, 10) ? strtoull(e, nullptr, 10) : BL;
		const uint64_t per_line = (uint64_t) C + 1;                         // the cluster id, then the columns
		std::vector<uint32_t> qbuf[2];
		std::vector<uint64_t> trip[2];                                      // below | upto << 21 | n << 42  (all <= R + STEP < 2^21)
		for (int b = 0; b < 2; ++b) {
			qbuf[b].resize(BL * C);
			trip[b].resize(BL * per_line);
		}
		const uint64_t nblocks = (n_lines + BL - 1) / BL;
		std::mutex mu;
		std::condition_variable cv;
		uint64_t produced = 0, consumed = 0;                                // blocks; guarded by mu
		std::atomic<int> bad(0);
		auto parallel = [&](auto fn) {
			std::vector<std::thread> th;
			for (unsigned t = 1; t < nthreads - 1; ++t) th.emplace_back(fn, t, nthreads - 1);
			fn(0u, nthreads - 1);
			for (auto &x : th) x.join();
		};
		std::thread producer([&]() {
			for (uint64_t b = 0; b < nblocks; ++b) {
				{
					std::unique_lock<std::mutex> lk(mu);
					cv.wait(lk, [&] { return b < consumed + 2 || bad.load(); });
				}
				if (bad.load()) break;
				const uint64_t l0 = b * BL, nl = std::min(BL, n_lines - l0);
				uint32_t *qb = qbuf[b & 1].data();
				uint64_t *tb = trip[b & 1].data();
				parallel([&](unsigned t, unsigned nt) {                     // (A) lines [la, lb): the model of every symbol
					const uint64_t la = nl * t / nt, lb = nl * (t + 1) / nt;
					for (uint64_t i = la; i < lb; ++i) {
						const uint32_t k = cluster_ids[l0 + i];
						if (k >= K) {
							bad.store(1);
							return;
						}
						const uint8_t *sym = symbols + (l0 + i) * C;
						uint32_t prev = 0;
						for (uint32_t c = 0; c < C; ++c) {
							const size_t kc = (size_t) k * C + c;
							const uint32_t ctx = cb->ctx_of[kc * A + prev];
							const uint32_t hi = sym[c] >> 7, state = sym[c] & 0x7Fu;
							if (ctx == QVZ_CTX_ABSENT) {
								bad.store(1);
								return;
							}
							const uint64_t qi = cb->q_off[kc] + 2 * ctx + hi;
							if (state >= card[qi]) {
								bad.store(1);
								return;
							}
							qb[i * C + c] = (uint32_t) qi;
							prev = out_sym[count_off[qi] + state];
						}
					}
				});
				if (bad.load()) break;
				parallel([&](unsigned t, unsigned nt) {                     // (B) columns [ca, cb): every model in line order
					const uint32_t ca = (uint32_t) ((uint64_t) C * t / nt), cbnd = (uint32_t) ((uint64_t) C * (t + 1) / nt);
					for (uint64_t i = 0; i < nl; ++i) {
						uint32_t below, upto, n;
						if (t == 0) {                                       // the cluster id is coded first (qv_write_cluster)
							Coder::model(cstats, cluster_ids[l0 + i], below, upto, n);
							tb[i * per_line] = (uint64_t) below | ((uint64_t) upto << 21) | ((uint64_t) n << 42);
						}
						const uint8_t *sym = symbols + (l0 + i) * C;
						for (uint32_t c = ca; c < cbnd; ++c) {
							Coder::model(stats[qb[i * C + c]], sym[c] & 0x7Fu, below, upto, n);
							tb[i * per_line + 1 + c] = (uint64_t) below | ((uint64_t) upto << 21) | ((uint64_t) n << 42);
						}
					}
				});
				{
					std::lock_guard<std::mutex> lk(mu);
					produced = b + 1;
				}
				cv.notify_all();
			}
			{
				std::lock_guard<std::mutex> lk(mu);
				produced = nblocks;                                         // also on error: releases the consumer
			}
			cv.notify_all();
		});
		for (uint64_t b = 0; b < nblocks; ++b) {
			{
				std::unique_lock<std::mutex> lk(mu);
				cv.wait(lk, [&] { return produced > b; });
			}
			if (bad.load()) break;
			const uint64_t l0 = b * BL, nl = std::min(BL, n_lines - l0);
			const uint64_t *tb = trip[b & 1].data();
			for (uint64_t j = 0; j < nl * per_line; ++j) {
				__builtin_prefetch(tb + j + 128);                           // the triples were written by other cores
				const uint64_t v = tb[j];
				coder.narrow((uint32_t) (v & 0x1FFFFFu), (uint32_t) ((v >> 21) & 0x1FFFFFu), (uint32_t) (v >> 42));
			}
			{
				std::lock_guard<std::mutex> lk(mu);
				consumed = b + 1;
			}
			cv.notify_all();
		}
		{
			std::lock_guard<std::mutex> lk(mu);
			consumed = nblocks + 2;                                         // on error: lets the producer leave its wait
		}
		cv.notify_all();
		producer.join();
		if (bad.load()) rc = -2;
	}
	if (!rc) coder.last();
	if (stream_bytes_out) *stream_bytes_out = os.written;
	if (!os.ok) rc = -1;
	if (fclose(fp) != 0 && !rc) rc = -1;
	return rc;
}

// ------------------------------------------------------------------------------------------------------ decoder
// decode() (src/main.c:132-160): read_codebooks (src/codebook.c:560-669) + start_qv_decompression
// (src/qv_compressor.c:145-231) with the arithmetic decoder of src/arith.c:118-205.  One sequential chain of
// adaptive-coder steps: host work by nature; here so that the command line is complete (qvz -x).
namespace {

struct BitReader {                                   // stream_read_bit (src/os_stream.c:35-51); zeros past the end of the file
	std::vector<uint8_t> buf;
	size_t pos = 0;
	uint32_t bit = 0;
	inline uint32_t next() {
		uint32_t v = 0;
		if (pos < buf.size()) v = (buf[pos] >> (7 - bit)) & 1u;
		if (++bit == 8) {
			bit = 0;
			++pos;
		}
		return v;
	}
	inline uint32_t bits(uint32_t k) {               // the next k <= 24 bits, first one most significant
		uint32_t v = 0;
		while (k) {
			const uint32_t take = k < 8 - bit ? k : 8 - bit;
			const uint32_t byte = pos < buf.size() ? buf[pos] : 0u;
			v = (v << take) | ((byte >> (8 - bit - take)) & ((1u << take) - 1u));
			bit += take;
			if (bit == 8) {
				bit = 0;
				++pos;
			}
			k -= take;
		}
		return v;
	}
};

struct Decoder {
	static constexpr uint32_t M = 22, MSB = M - 1, SMSB = M - 2, CLEAR = (1u << MSB) - 1, R = 1u << (M - 3), STEP = 8;
	uint32_t l = 0, u = (1u << M) - 1, t = 0;
	BitReader &is;
	explicit Decoder(BitReader &r) : is(r) {
		for (int b = (int) M - 1; b >= 0; --b) t |= is.next() << b;       // a->t = stream_read_bits(os, m) (src/qv_stream.c:113)
	}
	static void update(Stats &s, uint32_t x) {                              // update_stats (src/qv_stream.c:9-25)
		s.counts[x] += STEP;
		s.n += STEP;
		if (s.n > R) {
			s.n = 0;
			for (uint32_t i = 0; i < s.card; ++i)
				if (s.counts[i]) {
					s.counts[i] >>= 1;
					s.counts[i] += 1;
					s.n += s.counts[i];
				}
		}
	}
	inline uint32_t symbol(const Stats &s) const {                           // the search shared by both decoder steps
		const uint64_t range = (uint64_t) (u - l + 1), gap = (uint64_t) (t - l + 1);
		// floor through a double division (exact, see Coder::narrow): numerator < 2^43, quotient < 2^21, and a non-integer
		// quotient is at least 1/range >= 2^-22 away from the next integer
		const uint32_t sub = (uint32_t) ((double) (gap * s.n - 1) / (double) range);
		uint32_t k = 0, cum = 0;
		while (sub >= cum) cum += s.counts[k++];
		return k - 1;
	}
	inline uint32_t step(Stats &s) {                                         // arithmetic_decoder_step (src/arith.c:118-188)
		const uint64_t range = (uint64_t) (u - l + 1);
		const uint32_t x = symbol(s);
		uint32_t below = 0;
		for (uint32_t i = 0; i < x; ++i) below += s.counts[i];
		const uint32_t upto = below + s.counts[x];
		const double inv = 1.0 / (double) s.n;                              // as in Coder::narrow
		u = l + (uint32_t) ((double) (range * upto) * inv + 0x1p-25) - 1;
		l = l + (uint32_t) ((double) (range * below) * inv + 0x1p-25);
		const uint32_t MASK = (1u << M) - 1u;            // one run of E1/E2 turns, then one run of E3 turns: see Coder::narrow
		const uint32_t diff = l ^ u;
		const uint32_t k = diff ? (uint32_t) __builtin_clz(diff << (32 - M)) : M;
		if (k) {
			l = (l << k) & MASK;
			u = ((u << k) & MASK) | ((1u << k) - 1u);
			t = ((t << k) & MASK) | is.bits(k);
		}
		const uint32_t y = ((l & ~u) << 1) & MASK;
		const uint32_t j = (uint32_t) __builtin_clz(~(y << (32 - M)) | 1u);
		if (j) {                                         // every E3 turn flips the bit that becomes the top one; only the last flip stays inside
			l = (l << j) & CLEAR;
			u = ((u << j) & CLEAR) | (1u << MSB) | ((1u << j) - 1u);
			t = (((t << j) & MASK) ^ (1u << MSB)) | is.bits(j);
		}
		update(s, x);
		return x;
	}
};

// well_1024a + well_1024a_bits(7) (src/well.c:8-46)
struct Well {
	uint32_t s[32], n = 0, out = 0, left = 0;
	inline uint32_t word() {
		const uint32_t z0 = s[(n + 31) & 31], a = s[(n + 3) & 31], b = s[(n + 24) & 31], c = s[(n + 10) & 31];
		const uint32_t z1 = s[n] ^ (a ^ (a >> 8));
		const uint32_t z2 = (b ^ (b << 19)) ^ (c ^ (c << 14));
		s[n] = z1 ^ z2;
		n = (n + 31) & 31;
		s[n] = (z0 ^ (z0 << 11)) ^ (z1 ^ (z1 << 7)) ^ (z2 ^ (z2 << 13));
		return s[n];
	}
	inline uint32_t draw7() {
		if (left < 7) {
			out = word();
			left = 32;
		}
		const uint32_t r = out & 127u;
		out >>= 7;
		left -= 7;
		return r;
	}
};

// find_output_alphabet (src/quantizer.c:167-191): the decoder rebuilds output alphabets from runs of equal values
void find_output_alphabet(Quantizer &Q) {
	Q.out.sym.clear();
	uint8_t p = Q.q[0];
	Q.out.sym.push_back(p);
	for (uint32_t x = 1; x < A; ++x)
		if (Q.q[x] != p) {
			p = Q.q[x];
			Q.out.sym.push_back(p);
		}
	Q.out.reindex();
}

}  // namespace

extern "C" int qvz_host_decode(const char *in_path, const char *out_path, uint64_t *lines_out) {
	if (!in_path || !out_path) return -1;
	FILE *fin = fopen(in_path, "rb");
	if (!fin) return -1;
	std::vector<uint8_t> file;
	{
		fseek(fin, 0, SEEK_END);
		const long sz = ftell(fin);
		fseek(fin, 0, SEEK_SET);
		file.resize(sz > 0 ? (size_t) sz : 0);
		if (!file.empty() && fread(file.data(), 1, file.size(), fin) != file.size()) {
			fclose(fin);
			return -1;
		}
		fclose(fin);
	}
	size_t pos = 0;
	auto need = [&](size_t n) { return pos + n <= file.size(); };
	if (!need(9)) return -2;
	const uint32_t K = file[0];                      // read_codebooks (src/codebook.c:560-581): big-endian columns and lines
	const uint32_t C = (uint32_t) file[1] << 24 | (uint32_t) file[2] << 16 | (uint32_t) file[3] << 8 | file[4];
	const uint32_t lines = (uint32_t) file[5] << 24 | (uint32_t) file[6] << 16 | (uint32_t) file[7] << 8 | file[8];
	pos = 9;
	if (K == 0 || C == 0 || C > QVZ_MAX_COLUMNS) return -2;
	std::vector<ClusterBook> books(K);
	auto read_q = [&](Quantizer &Q) -> bool {        // COPY_Q_FROM_LINE (include/codebook.h:102)
		if (!need(A)) return false;
		for (uint32_t i = 0; i < A; ++i) {
			Q.q[i] = (uint8_t) (file[pos + i] - 33);
			if (Q.q[i] >= A) return false;
		}
		pos += A;
		find_output_alphabet(Q);
		return true;
	};
	auto eol = [&]() -> bool {                       // every codebook line ends with one '\n'
		if (!need(1) || file[pos] != '\n') return false;
		++pos;
		return true;
	};
	for (uint32_t k = 0; k < K; ++k) {               // read_codebook (src/codebook.c:586-669)
		ClusterBook &B = books[k];
		B.in.resize(C);
		B.q.resize(C);
		B.qratio.resize(C);
		B.in[0].sym.assign(1, 0);
		B.in[0].reindex();
		B.q[0].resize(2);
		B.qratio[0].resize(1);
		if (!need(2)) return -2;
		B.qratio[0][0] = (uint8_t) (file[pos] - 33);
		++pos;
		if (!eol() || !read_q(B.q[0][0]) || !eol() || !read_q(B.q[0][1]) || !eol()) return -2;
		Alphabet uniques = alphabet_union(B.q[0][0].out, B.q[0][1].out);
		for (uint32_t c = 1; c < C; ++c) {
			B.in[c] = uniques;
			const size_t nc = uniques.sym.size();
			B.q[c].resize(2 * nc);
			B.qratio[c].resize(nc);
			if (!need(nc)) return -2;
			for (size_t i = 0; i < nc; ++i) B.qratio[c][i] = (uint8_t) (file[pos + i] - 33);
			pos += nc;
			if (!eol()) return -2;
			Alphabet next;
			next.reindex();
			for (int hi = 0; hi < 2; ++hi) {
				for (size_t i = 0; i < nc; ++i) {
					if (!read_q(B.q[c][2 * i + hi])) return -2;
					next = alphabet_union(next, B.q[c][2 * i + hi].out);
				}
				if (!eol()) return -2;
			}
			uniques = next;
		}
	}
	if (!need(128)) return -2;
	Well well;                                       // initialize_arithStream, decompressor side (src/qv_stream.c:72-74, 93)
	memcpy(well.s, &file[pos], 128);
	pos += 128;
	BitReader is;
	is.buf.assign(file.begin() + pos, file.end());

	// adaptive models, exactly as the encoder sets them up but from the decoder's output alphabets -- and everything the
	// per-symbol loop touches flattened like in qvz_host_encode (one dependent load per lookup instead of a chain of
	// vector headers): ctx_flat[kc][prev], ratio_flat[kc][ctx], models qoff[kc] + 2*ctx + hi, out_flat[model][state]
	const size_t KC = (size_t) K * C;
	std::vector<uint8_t> ctx_flat(KC * A, 0xFF), ratio_flat(KC * A, 0);
	std::vector<uint64_t> qoff(KC + 1, 0);
	for (uint32_t k = 0; k < K; ++k)
		for (uint32_t c = 0; c < C; ++c) {
			const size_t kc = (size_t) k * C + c;
			const ClusterBook &B = books[k];
			for (uint32_t v = 0; v < A; ++v) {
				const uint32_t ctx = B.in[c].idx[v];
				if (ctx != NOT_FOUND) {
					ctx_flat[kc * A + v] = (uint8_t) ctx;
					ratio_flat[kc * A + ctx] = B.qratio[c][ctx];
				}
			}
			qoff[kc + 1] = qoff[kc] + B.q[c].size();
		}
	const uint64_t nq = qoff[KC];
	std::vector<uint64_t> sym_off(nq + 1, 0);
	for (uint32_t k = 0; k < K; ++k)
		for (uint32_t c = 0; c < C; ++c)
			for (size_t j = 0; j < books[k].q[c].size(); ++j) {
				const uint64_t qi = qoff[(size_t) k * C + c] + j;
				sym_off[qi + 1] = books[k].q[c][j].out.sym.size();
			}
	for (uint64_t i = 0; i < nq; ++i) sym_off[i + 1] += sym_off[i];
	std::vector<uint32_t> cnt(sym_off[nq], 1u);
	std::vector<uint8_t> out_flat(sym_off[nq]);
	std::vector<Stats> st(nq);
	for (uint32_t k = 0; k < K; ++k)
		for (uint32_t c = 0; c < C; ++c)
			for (size_t j = 0; j < books[k].q[c].size(); ++j) {
				const uint64_t qi = qoff[(size_t) k * C + c] + j;
				const uint32_t card = (uint32_t) (sym_off[qi + 1] - sym_off[qi]);
				st[qi] = Stats{&cnt[sym_off[qi]], card, card};
				memcpy(&out_flat[sym_off[qi]], books[k].q[c][j].out.sym.data(), card);
			}
	std::vector<uint32_t> ccounts(K, 1u);
	Stats cstats{ccounts.data(), K, K};

	FILE *fout = fopen(out_path, "wt");
	if (!fout) return -1;
	Decoder dec(is);
	std::vector<uint8_t> line(C + 1);
	line[C] = '\n';
	int rc = 0;
	for (uint64_t ln = 0; ln < lines && !rc; ++ln) {
		const uint32_t k = dec.step(cstats);                                 // qv_read_cluster
		if (k >= K) {
			rc = -2;
			break;
		}
		uint32_t prev = 0;
		for (uint32_t c = 0; c < C; ++c) {
			const size_t kc = (size_t) k * C + c;
			const uint32_t ctx = ctx_flat[kc * A + prev];                    // choose_quantizer (src/codebook.c:162-171)
			if (ctx == 0xFF) {
				rc = -2;
				break;
			}
			const uint64_t qi = qoff[kc] + 2 * ctx + (well.draw7() >= ratio_flat[kc * A + ctx] ? 1u : 0u);
			Stats &s = st[qi];
			uint32_t state;
			if (ln + 1 == lines && c + 1 == C) state = dec.symbol(s);        // decoder_last_step (src/arith.c:190-205): no more bits are read
			else state = dec.step(s);
			if (state >= s.card) {
				rc = -2;
				break;
			}
			prev = out_flat[sym_off[qi] + state];
			line[c] = (uint8_t) (prev + 33);
		}
		if (!rc && fwrite(line.data(), 1, C + 1, fout) != C + 1) rc = -1;
	}
	if (fclose(fout) != 0 && !rc) rc = -1;
	if (lines_out) *lines_out = lines;
	return rc;
}
