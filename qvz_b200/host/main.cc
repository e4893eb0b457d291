// qvz -- the reference's command line (src/main.c:166-355: -q -x -f -r -d -D -c -T -u -h -s -v, same parsing, same
// messages) over the B200 front end.  encode() keeps the reference's flow (src/main.c:18-127):
//
//   load_file                 -> mmap + qvz_gpu_load_rows (rows packed once into HBM; no per-line pointer table)
//   do_kmeans_clustering      -> initial centroids picked on the host with libc rand() exactly like
//                                initialize_kmeans_clustering (src/cluster.c:192-206), iterations on the GPU
//   calculate_statistics      -> qvz_gpu_cond_counts
//   generate_codebooks        -> qvz_host_design   (host, small alphabets)
//   write_codebooks +
//   start_qv_compression      -> qvz_gpu_quantize (symbol stream, -u image, per-line distortion) + qvz_host_encode
//                                (codebook text, WELL seed, arithmetic coder)
//
// QVZ_GPUS=n (environment, default 1) shards the lines over n devices of the box from this one process: the integer
// centroid sums (<= 6 KB) and the count tables are added on the host between the stepping calls, each device
// quantizes its shard from its own position in the WELL stream.  The .qvz bytes do not depend on n.
// There is no CPU fallback: without a CUDA device the program stops with the library's error.
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <string>
#include <thread>
#include <vector>

#include "../../include/qvz_gpu.h"
#include "../../include/qvz_host.h"

#define MAX_LINES_PER_BLOCK 1000000ull   /* include/lines.h:12 */
#define DISTORTION_CUSTOM 4              /* include/distortion.h:10 */

struct options {
	uint8_t verbose = 0, stats = 0, uncompressed = 0, distortion = QVZ_DIST_MSE, mode = QVZ_MODE_RATIO;
	double ratio = 0.5, cluster_threshold = 4;
	uint32_t clusters = 1;
	const char *uncompressed_name = nullptr, *dist_file = nullptr;
};

static double now() {
	struct timespec ts;
	clock_gettime(CLOCK_REALTIME, &ts);
	return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static void die_gpu(qvz_gpu *h, const char *what, int rc) {
	printf("%s failed (%d): %s\n", what, rc, qvz_gpu_last_error(h));
	exit(1);
}

struct shard {
	qvz_gpu *h = nullptr;
	uint64_t l0 = 0, l1 = 0;
};

// one host thread per device for the calls that block on their device
template <class F>
static void for_each_shard(std::vector<shard> &sh, F fn) {
	if (sh.size() == 1) {
		fn(sh[0], 0);
		return;
	}
	std::vector<std::thread> pool;
	for (size_t g = 0; g < sh.size(); ++g) pool.emplace_back([&, g]() { fn(sh[g], g); });
	for (auto &t : pool) t.join();
}

static void encode(const char *input_name, const char *output_name, options *opts) {
	const double t_total = now();
	double D[QVZ_ALPHABET * QVZ_ALPHABET];
	if (opts->distortion == DISTORTION_CUSTOM) {
		if (qvz_host_distortion_file(opts->dist_file, D)) {
			perror("Unable to open distortion definition file");
			exit(1);
		}
	} else qvz_host_distortion(opts->distortion, D);

	// load_file (src/lines.c:27-82): columns from the first line, lines from the file size
	FILE *fp = fopen(input_name, "rt");
	const int fd = open(input_name, O_RDONLY);
	if (!fp || fd == -1) {
		printf("load_file returned error: %d\n", 1);
		exit(1);
	}
	char first[1024];
	if (!fgets(first, sizeof(first), fp)) first[0] = 0;
	fclose(fp);
	const size_t len = strlen(first);
	const uint32_t columns = len ? (uint32_t) (len - 1) : 0;
	if (columns > QVZ_MAX_COLUMNS || columns == 0) {
		printf("load_file returned error: %d\n", 2);
		exit(1);
	}
	struct stat finfo;
	fstat(fd, &finfo);
	const uint64_t lines = (uint64_t) finfo.st_size / ((uint64_t) columns + 1);
	if (lines == 0) {
		printf("load_file returned error: %d\n", 1);
		exit(1);
	}
	const uint8_t *file = (const uint8_t *) mmap(NULL, finfo.st_size, PROT_READ, MAP_SHARED, fd, 0);
	if (file == MAP_FAILED) {
		perror("mmap");
		exit(1);
	}

	// shards of lines, one device each; inner boundaries on multiples of 4 lines (WELL word boundaries)
	int ngpu = getenv("QVZ_GPUS") ? atoi(getenv("QVZ_GPUS")) : 1;
	if (ngpu < 1) ngpu = 1;
	if ((uint64_t) ngpu * 4 > lines) ngpu = 1;
	std::vector<shard> sh(ngpu);
	for (int g = 0; g < ngpu; ++g) {
		sh[g].l0 = g ? ((lines * g / ngpu) + 3) & ~3ull : 0;
		if (g) sh[g - 1].l1 = sh[g].l0;
	}
	sh[ngpu - 1].l1 = lines;
	const uint32_t K = opts->clusters, C = columns;
	for_each_shard(sh, [&](shard &s, size_t g) {
		int rc = qvz_gpu_open(&s.h, (int) g);
		if (rc) die_gpu(s.h, "qvz_gpu_open", rc);
		rc = qvz_gpu_load_rows(s.h, file + s.l0 * (C + 1), s.l1 - s.l0, C, C + 1, s.l0);
		if (rc) die_gpu(s.h, "qvz_gpu_load_rows", rc);
	});

	// ---- do_kmeans_clustering (src/cluster.c:212-244)
	const double t_cluster = now();
	const uint64_t block_count = (lines + MAX_LINES_PER_BLOCK - 1) / MAX_LINES_PER_BLOCK;
	std::vector<uint8_t> means((size_t) K * C);
	for (uint32_t j = 0; j < K; ++j) {               // initialize_kmeans_clustering (:192-206): two rand() per cluster
		const uint32_t block_id = rand() % block_count;
		const uint64_t in_block = (block_id + 1 < block_count || lines % MAX_LINES_PER_BLOCK == 0) ? MAX_LINES_PER_BLOCK : lines % MAX_LINES_PER_BLOCK;
		const uint32_t line_id = rand() % in_block;
		memcpy(&means[(size_t) j * C], file + (block_id * MAX_LINES_PER_BLOCK + line_id) * (C + 1), C);
		if (opts->verbose) printf("Chose block %d, line %d.\n", block_id, line_id);
	}
	// initialize_arithStream (src/qv_stream.c:76-90): the WELL seed is 32 rand() values after srand(time(0)), drawn
	// AFTER the k-means picks above consumed the unseeded rand() stream; -DDEBUG builds of the reference use
	// 0x55555555 instead, selected here with QVZ_DEBUG_SEED=1.  Nothing else depends on the seed, so it is drawn
	// now and the devices generate their draws underneath k-means, counting and codebook design.
	uint32_t seed[32];
	srand((uint32_t) time(0));
	for (int i = 0; i < 32; ++i) seed[i] = getenv("QVZ_DEBUG_SEED") ? 0x55555555u : (uint32_t) rand();
	for (auto &s : sh) {
		int rc = qvz_gpu_prefetch_draws(s.h, seed);
		if (rc) die_gpu(s.h, "qvz_gpu_prefetch_draws", rc);
	}
	std::vector<uint8_t> ids(lines);
	uint32_t iter_count = 0;
	if (ngpu == 1) {
		std::vector<double> moved((size_t) QVZ_MAX_KMEANS_ITER * K);
		int rc = qvz_gpu_kmeans(sh[0].h, K, means.data(), opts->cluster_threshold, QVZ_MAX_KMEANS_ITER, ids.data(), nullptr, nullptr,
		                        moved.data(), &iter_count);
		if (rc) die_gpu(sh[0].h, "qvz_gpu_kmeans", rc);
		if (opts->verbose)
			for (uint32_t it = 0; it < iter_count; ++it) {
				for (uint32_t k = 0; k < K; ++k) printf("Cluster %d moved %f.\n", k, moved[(size_t) it * K + k]);
				printf("\n");
			}
	} else {
		const size_t ns = (size_t) K * C + K;
		std::vector<int64_t> total(ns);
		std::vector<std::vector<int64_t>> part(ngpu, std::vector<int64_t>(ns));
		std::vector<std::vector<double>> moved_g(ngpu, std::vector<double>(K));
		std::vector<double> &moved = moved_g[0];
		for (auto &s : sh) {
			int rc = qvz_gpu_kmeans_begin(s.h, K, means.data());
			if (rc) die_gpu(s.h, "qvz_gpu_kmeans_begin", rc);
		}
		bool loop = true;
		while (iter_count < QVZ_MAX_KMEANS_ITER && loop) {
			std::fill(total.begin(), total.end(), 0);
			for_each_shard(sh, [&](shard &s, size_t g) {     // every device assigns its shard ...
				int rc = qvz_gpu_kmeans_assign_host(s.h, part[g].data());
				if (rc) die_gpu(s.h, "qvz_gpu_kmeans_assign_host", rc);
			});
			for (int g = 0; g < ngpu; ++g)                   // ... the integer sums are added here ...
				for (size_t i = 0; i < ns; ++i) total[i] += part[g][i];
			for_each_shard(sh, [&](shard &s, size_t g) {     // ... and every device recentres on the same totals
				int rc = qvz_gpu_kmeans_update_host(s.h, total.data(), moved_g[g].data(), nullptr);
				if (rc) die_gpu(s.h, "qvz_gpu_kmeans_update_host", rc);
			});
			double move_max = 0;
			for (uint32_t k = 0; k < K; ++k) {
				if (moved[k] > move_max) move_max = moved[k];
				if (opts->verbose) printf("Cluster %d moved %f.\n", k, moved[k]);
			}
			loop = move_max > opts->cluster_threshold;
			iter_count += 1;
			if (opts->verbose) printf("\n");
		}
		for_each_shard(sh, [&](shard &s, size_t) {
			int rc = qvz_gpu_kmeans_end(s.h, ids.data() + s.l0, nullptr);
			if (rc) die_gpu(s.h, "qvz_gpu_kmeans_end", rc);
		});
	}
	if (opts->verbose) {
		printf("\nTotal number of iterations: %d.\n", iter_count);
		printf("Clustering took %.4f seconds\n", now() - t_cluster);
	}

	// ---- calculate_statistics + generate_codebooks
	const double t_stats = now();
	const uint64_t ncount = qvz_gpu_cond_counts_len(K, C);
	std::vector<uint32_t> counts(ncount, 0);
	if (ngpu == 1) {
		int rc = qvz_gpu_cond_counts(sh[0].h, counts.data());
		if (rc) die_gpu(sh[0].h, "qvz_gpu_cond_counts", rc);
	} else {
		std::vector<std::vector<uint32_t>> part(ngpu, std::vector<uint32_t>(ncount));
		for_each_shard(sh, [&](shard &s, size_t g) {
			int rc = qvz_gpu_cond_counts(s.h, part[g].data());
			if (rc) die_gpu(s.h, "qvz_gpu_cond_counts", rc);
		});
		for (int g = 0; g < ngpu; ++g)
			for (uint64_t i = 0; i < ncount; ++i) counts[i] += part[g][i];
	}
	qvz_codebooks *cb = qvz_host_design(counts.data(), K, C, opts->mode, opts->ratio, D, 0);
	if (!cb) {
		printf("codebook design rejected its arguments\n");
		exit(1);
	}
	if (opts->verbose) printf("Stats and codebook generation took %.4f seconds\n", now() - t_stats);

	// ---- write_codebooks + start_qv_compression
	struct qvz_flat_tables tables;
	qvz_host_tables(cb, &tables);
	std::vector<uint8_t> symbols((size_t) lines * C), qv;
	std::vector<double> line_err(lines);
	if (opts->uncompressed) qv.resize((size_t) lines * (C + 1));
	for_each_shard(sh, [&](shard &s, size_t) {
		int rc = qvz_gpu_quantize(s.h, &tables, seed, symbols.data() + s.l0 * C, opts->uncompressed ? qv.data() + s.l0 * (C + 1) : nullptr,
		                          line_err.data() + s.l0);
		if (rc) die_gpu(s.h, "qvz_gpu_quantize", rc);
	});
	if (opts->verbose)
		for (uint64_t b = 0; b < block_count; ++b) printf("Line: %dM\n", (int) b);
	if (opts->uncompressed) {
		FILE *fu = fopen(opts->uncompressed_name, "w");
		if (!fu) {
			perror("Unable to open uncompressed file");
			exit(1);
		}
		fwrite(qv.data(), 1, qv.size(), fu);
		fclose(fu);
	}
	uint64_t stream_bytes = 0;
	int rc = qvz_host_encode(cb, output_name, lines, ids.data(), symbols.data(), seed, &stream_bytes);
	if (rc == -1) {
		perror("Unable to open output file");
		exit(1);
	} else if (rc) {
		printf("arithmetic coder rejected the symbol stream (%d)\n", rc);
		exit(1);
	}
	double distortion = 0.0;                         // distortion += error / columns, in line order (src/qv_compressor.c:127,140)
	for (uint64_t i = 0; i < lines; ++i) distortion += line_err[i];
	distortion = distortion / ((double) lines);
	const uint64_t bytes_used = (uint32_t) stream_bytes;      // start_qv_compression returns uint32_t (src/qv_compressor.c:48)
	const double elapsed = now() - t_total;

	if (opts->verbose) {
		switch (opts->distortion) {
		case QVZ_DIST_MANHATTAN: printf("L1 distortion: %f\n", distortion); break;
		case QVZ_DIST_MSE: printf("MSE distortion: %f\n", distortion); break;
		case QVZ_DIST_LORENTZ: printf("log(1+L1) distortion: %f\n", distortion); break;
		case DISTORTION_CUSTOM: printf("Custom distortion: %f\n", distortion); break;
		default: break;
		}
		printf("Lines: %llu\n", (unsigned long long) lines);
		printf("Columns: %u\n", C);
		printf("Total bytes used: %llu\n", (unsigned long long) bytes_used);
		printf("Encoding took %.4f seconds.\n", elapsed);
		printf("Total time elapsed: %.4f seconds.\n", elapsed);
	}
	if (opts->stats)
		printf("rate, %.4f, distortion, %.4f, time, %.4f, size, %llu \n", (bytes_used * 8.) / ((double) (lines) * C), distortion, elapsed,
		       (unsigned long long) bytes_used);
	qvz_host_free(cb);
	for (auto &s : sh) qvz_gpu_close(s.h);
	munmap((void *) file, finfo.st_size);
	close(fd);
}

static void usage(const char *name) {
	printf("Usage: %s (options) [input file] [output file]\n", name);
	printf("Options are:\n");
	printf("   -q           : Store quality values in compressed file (default)\n");
	printf("   -x           : Extract quality values from compressed file\n");
	printf("   -f [ratio]   : Compress using [ratio] bits per bit of input entropy per symbol\n");
	printf("   -r [rate]    : Compress using fixed [rate] bits per symbol\n");
	printf("   -d [M|L|A]   : Optimize for MSE, Log(1+L1), L1 distortions, respectively (default: MSE)\n");
	printf("   -D [FILE]    : Optimize using the custom distortion matrix specified in FILE\n");
	printf("   -c [#]       : Compress using [#] clusters (default: 1)\n");
	printf("   -T [#]       : Use [#] as a threshold for cluster center movement (L2 norm) to declare a stable solution (default: 4).\n");
	printf("   -u [FILE]    : Write the uncompressed lossy values to FILE (default: off)\n");
	printf("   -h           : Print this help\n");
	printf("   -s           : Print summary stats\n");
	printf("   -v           : Enable verbose output\n");
	printf("\nFor custom distortion matrices, a 72x72 matrix of values must be provided as the cost of reconstructing\n");
	printf("the x-th row as the y-th column, where x and y range from 0 to 71 (inclusive) corresponding to the possible\n");
	printf("Phred scores.\n");
}

int main(int argc, char **argv) {
	const char *input_name = 0, *output_name = 0;
	options opts;
	uint8_t extract = 0, file_idx = 0;
	int i = 1;
	auto need = [&](int) {                           // the reference reads argv[i+1] unchecked; stop cleanly instead
		if (i + 1 >= argc) {
			printf("Option %s needs a value.\n", argv[i]);
			usage(argv[0]);
			exit(1);
		}
	};
	while (i < argc) {
		if (argv[i][0] != '-') {
			switch (file_idx) {
			case 0: input_name = argv[i]; file_idx = 1; break;
			case 1: output_name = argv[i]; file_idx = 2; break;
			default:
				printf("Garbage argument \"%s\" detected.\n", argv[i]);
				usage(argv[0]);
				exit(1);
			}
			i += 1;
			continue;
		}
		switch (argv[i][1]) {
		case 'x': extract = 1; i += 1; break;
		case 'q': extract = 0; i += 1; break;
		case 'f': need(1); extract = 0; opts.ratio = atof(argv[i + 1]); opts.mode = QVZ_MODE_RATIO; i += 2; break;
		case 'r': need(1); extract = 0; opts.ratio = atof(argv[i + 1]); opts.mode = QVZ_MODE_FIXED; i += 2; break;
		case 'c': need(1); opts.clusters = (uint8_t) atoi(argv[i + 1]); i += 2; break;      // qv_options_t.clusters is a uint8_t (include/codebook.h:31)
		case 'v': opts.verbose = 1; i += 1; break;
		case 'h': usage(argv[0]); exit(0);
		case 's': opts.stats = 1; i += 1; break;
		case 'u': need(1); opts.uncompressed = 1; opts.uncompressed_name = argv[i + 1]; i += 2; break;
		case 'T': need(1); opts.cluster_threshold = atoi(argv[i + 1]); i += 2; break;
		case 'd':
			need(1);
			switch (argv[i + 1][0]) {
			case 'M': opts.distortion = QVZ_DIST_MSE; break;
			case 'L': opts.distortion = QVZ_DIST_LORENTZ; break;
			case 'A': opts.distortion = QVZ_DIST_MANHATTAN; break;
			default: printf("Distortion measure not supported, using MSE.\n"); break;
			}
			i += 2;
			break;
		case 'D': need(1); opts.distortion = DISTORTION_CUSTOM; opts.dist_file = argv[i + 1]; i += 2; break;
		default:
			printf("Unrecognized option -%c.\n", argv[i][1]);
			usage(argv[0]);
			exit(1);
		}
	}
	if (file_idx != 2) {
		printf("Missing required filenames.\n");
		usage(argv[0]);
		exit(1);
	}
	if (opts.clusters < 1) {                         // the reference allocates zero clusters and crashes
		printf("At least one cluster is needed.\n");
		exit(1);
	}
	if (opts.verbose) {
		if (extract) printf("%s will be decoded to %s.\n", input_name, output_name);
		else {
			printf("%s will be encoded as %s.\n", input_name, output_name);
			if (opts.mode == QVZ_MODE_RATIO) printf("Ratio mode selected, targeting %f compression ratio.\n", opts.ratio);
			else printf("Fixed-rate mode selected, targeting %f bits per symbol.\n", opts.ratio);
			switch (opts.distortion) {
			case QVZ_DIST_MSE: printf("MSE will be used as a distortion metric.\n"); break;
			case QVZ_DIST_LORENTZ: printf("log(1+L1) will be used as a distortion metric.\n"); break;
			case QVZ_DIST_MANHATTAN: printf("L1 will be used as a distortion metric.\n"); break;
			case DISTORTION_CUSTOM: printf("A custom distortion metric stored in %s will be used.\n", opts.dist_file); break;
			}
			printf("Compression will use %d clusters, with a movement threshold of %.0f.\n", opts.clusters, opts.cluster_threshold);
		}
	}
	if (extract) {                                   // decode() (src/main.c:132-160): sequential host work
		const double t0 = now();
		uint64_t lines = 0;
		const int rc = qvz_host_decode(input_name, output_name, &lines);
		if (rc == -1) {
			perror("Unable to open input or output files");
			exit(1);
		} else if (rc) {
			printf("%s is not a valid qvz file.\n", input_name);
			exit(1);
		}
		if (opts.verbose) printf("Decoded %llu lines in %f seconds.\n", (unsigned long long) lines, now() - t0);
		return 0;
	}
	encode(input_name, output_name, &opts);
	return 0;
}
