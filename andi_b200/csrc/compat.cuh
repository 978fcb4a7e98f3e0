// andi_b200/csrc/compat.cuh -- part B of the C ABI (include/andi_compat.h); included by
// andi_b200.cu. The reference's own symbols (src/esa.h:61-64, src/process.c:141) as thin,
// mutex-serialised wrappers around one lazily created GPU context.
#pragma once
#include "../../include/andi_compat.h"

#include <mutex>
#include <unordered_map>

// the reference's globals (src/global.h:20-48, defined in src/andi.c:45-50) and the host-side functions
// calculate_distances hands its results to, if the host program has them
extern "C" {
extern int MODEL __attribute__((weak));
extern int FLAGS __attribute__((weak));
extern double ANCHOR_P_VALUE __attribute__((weak));
extern long unsigned int BOOTSTRAP __attribute__((weak));
void print_distances(const model *, const seq_t *, size_t, int) __attribute__((weak));
void print_coverages(const model *, size_t) __attribute__((weak));
model model_average(const model *, const model *) __attribute__((weak));
model model_bootstrap(model) __attribute__((weak));
}

namespace {
std::mutex g_compat_mutex;
andi_ctx *g_compat_ctx = nullptr;
std::unordered_map<const void *, andi_esa *> g_compat_index;  // key: the host SA array of the esa_s
int g_compat_model = ANDI_M_JC;

andi_ctx *compat_ctx() {
	if (!g_compat_ctx) {
		const char *env = getenv("ANDI_B200_DEVICE");
		int device = env ? atoi(env) : 0;
		if (andi_ctx_create(device, nullptr, &g_compat_ctx) != ANDI_OK) {
			fprintf(stderr, "andi_b200: cannot create a CUDA context: %s\n", andi_last_error(nullptr));
			g_compat_ctx = nullptr;
		}
	}
	return g_compat_ctx;
}

andi_esa *compat_lookup(const esa_s *C) {
	if (!C || !C->SA) return nullptr;
	auto it = g_compat_index.find((const void *)C->SA);
	return it == g_compat_index.end() ? nullptr : it->second;
}
}  // namespace

extern "C" void andi_compat_set_model(int model_id) {
	std::lock_guard<std::mutex> lock(g_compat_mutex);
	g_compat_model = model_id;
}

extern "C" int esa_init(esa_s *C, const seq_subject *S) {
	if (!C || !S || !S->RS) return 1;  // src/esa.c:255
	std::lock_guard<std::mutex> lock(g_compat_mutex);
	memset(C, 0, sizeof *C);
	C->S = S->RS;
	C->len = (saidx_t)S->RSlen;
	andi_ctx *ctx = compat_ctx();
	if (!ctx) return ANDI_ERR_CUDA;
	andi_esa *E = nullptr;
	int rc = andi_esa_build_rs(ctx, S->RS, S->RSlen, ANDI_ESA_FULL, &E);
	if (rc) return rc;
	const size_t N = S->RSlen;
	C->SA = (saidx_t *)malloc(N * sizeof(saidx_t));
	C->LCP = (saidx_t *)malloc((N + 1) * sizeof(saidx_t));
	C->CLD = (saidx_t *)malloc((N + 1) * sizeof(saidx_t));
	C->FVC = (char *)malloc(N);
	C->cache = (lcp_inter_t *)malloc(sizeof(lcp_inter_t) << 20);
	if (!C->SA || !C->LCP || !C->CLD || !C->FVC || !C->cache) {
		// the reference exits through err(errno, "Out of memory") (src/global.h:74-79)
		fprintf(stderr, "andi_b200: Out of memory\n");
		exit(12);
	}
	rc = andi_esa_download(E, C->SA, C->LCP, C->CLD, C->FVC, (andi_lcp_inter *)C->cache);
	if (rc) {
		andi_esa_free(E);
		free(C->SA), free(C->LCP), free(C->CLD), free(C->FVC), free(C->cache);
		const char *keep = C->S;
		memset(C, 0, sizeof *C);
		C->S = keep;
		return rc;
	}
	if (S->threshold) E->threshold = (u32)S->threshold;
	g_compat_index[(const void *)C->SA] = E;
	return 0;
}

extern "C" void esa_free(esa_s *C) {
	if (!C) return;
	std::lock_guard<std::mutex> lock(g_compat_mutex);
	andi_esa *E = compat_lookup(C);
	if (E) {
		g_compat_index.erase((const void *)C->SA);
		andi_esa_free(E);
	}
	free(C->SA), free(C->LCP), free(C->CLD), free(C->cache), free(C->FVC);
	memset(C, 0, sizeof *C);
}

static lcp_inter_t compat_match(const esa_s *C, const char *query, size_t qlen) {
	lcp_inter_t bad = {-1, -1, -1, -1};
	if (!C || !query || !C->len || !C->SA || !C->LCP || !C->S || !C->CLD) return bad;  // src/esa.c:616-618
	std::lock_guard<std::mutex> lock(g_compat_mutex);
	andi_esa *E = compat_lookup(C);
	if (!E) return bad;
	andi_lcp_inter r;
	if (qlen == 0) {
		lcp_inter_t whole = {0, 0, C->len - 1, -1};
		return whole;
	}
	if (andi_esa_get_match(E, &query, &qlen, 1, &r) != ANDI_OK) return bad;
	lcp_inter_t out = {r.l, r.i, r.j, r.m};
	return out;
}

extern "C" lcp_inter_t get_match(const esa_s *C, const char *query, size_t qlen) {
	return compat_match(C, query, qlen);
}

extern "C" lcp_inter_t get_match_cached(const esa_s *C, const char *query, size_t qlen) {
	return compat_match(C, query, qlen);
}

extern "C" model dist_anchor(const esa_s *C, const char *query, size_t query_length, size_t threshold) {
	model ret;
	memset(&ret, 0, sizeof ret);
	ret.seq_len = (unsigned int)query_length;
	std::lock_guard<std::mutex> lock(g_compat_mutex);
	andi_esa *E = compat_lookup(C);
	if (!E) {
		fprintf(stderr, "andi_b200: dist_anchor called with an index that esa_init did not build\n");
		exit(1);  // the reference has no error return here; a wrong answer would be worse
	}
	int model_id = (&MODEL != nullptr) ? MODEL : g_compat_model;
	andi_model m;
	int rc = andi_dist_anchor(E->ctx, E, query, query_length, threshold, model_id, &m);
	if (rc) {
		fprintf(stderr, "andi_b200: dist_anchor failed: %s\n", andi_last_error(E->ctx));
		exit(1);
	}
	memcpy(&ret, &m, sizeof ret);
	return ret;
}

// ------------------------------------------------------------------ the driver level
// src/dist_hack.h:34-96 and src/process.c:230-321 as ONE batched GPU run.

static int compat_devices(int *out, int cap) {
	const char *list = getenv("ANDI_B200_DEVICES");
	int n = 0;
	if (list && *list) {
		if (!strcmp(list, "all")) {
			for (int d = 0, c = andi_device_count(); d < c && n < cap; d++) out[n++] = d;
			return n;
		}
		const char *p = list;
		while (*p && n < cap) {
			char *end;
			long a = strtol(p, &end, 10), b = a;
			if (end == p || a < 0) break;
			if (*end == '-') {
				p = end + 1;
				b = strtol(p, &end, 10);
				if (end == p || b < a) break;
			}
			for (long d = a; d <= b && n < cap; d++) out[n++] = (int)d;
			p = *end == ',' ? end + 1 : end;
			if (*end && *end != ',') break;
		}
		if (n) return n;
	}
	const char *one = getenv("ANDI_B200_DEVICE");
	out[0] = one ? atoi(one) : 0;
	return 1;
}

static void compat_progress(size_t done, size_t total, void *user) {
	// src/dist_hack.h:74-90
	fprintf(stderr, "\rComparing %zu sequences: %5.1f%% (%zu/%zu)", *(const size_t *)user,
			total ? 100.0 * (double)done / (double)total : 100.0, done, total);
}

static void compat_matrix(model *M, const seq_t *sequences, size_t n, int low_memory) {
	if (!M || !sequences || !n) return;
	std::vector<const char *> ptr(n);
	std::vector<size_t> len(n);
	for (size_t i = 0; i < n; i++) ptr[i] = sequences[i].S, len[i] = sequences[i].len;
	int devices[64];
	const int nd = compat_devices(devices, 64);
	const bool show = (&FLAGS != nullptr) && (FLAGS & 128);	 // F_PRINT_PROGRESS, src/global.h:65
	const double p = (&ANCHOR_P_VALUE != nullptr) ? ANCHOR_P_VALUE : 0.025;
	const int model_id = (&MODEL != nullptr) ? MODEL : g_compat_model;
	if (show) compat_progress(0, n * n - n, &n);
	char msg[512];
	static_assert(sizeof(model) == sizeof(andi_model), "struct model layout");
	int rc = andi_dist_matrix_multi(devices, nd, ptr.data(), len.data(), n, p, model_id, low_memory, (andi_model *)M,
									show ? compat_progress : nullptr, &n, msg, sizeof msg);
	if (rc) {
		// src/dist_hack.h:53: errx(1, "Failed to create index for %s.", ...)
		fprintf(stderr, "andi: Failed to create index: %s\n", msg);
		exit(1);
	}
	if (show) fprintf(stderr, ", done.\n");
}

extern "C" void distMatrix(model *M, const seq_t *sequences, size_t n) { compat_matrix(M, sequences, n, 0); }
extern "C" void distMatrixLM(model *M, const seq_t *sequences, size_t n) { compat_matrix(M, sequences, n, 1); }

extern "C" void calculate_distances(seq_t *sequences, size_t n) {
	if (!print_distances || !model_average) {
		fprintf(stderr, "andi_b200: calculate_distances needs the host program's print_distances / model_average "
						"(link the reference's io.c and model.c, with -rdynamic)\n");
		exit(1);
	}
	// src/process.c:233-245
	if (n == 0 || SIZE_MAX / sizeof(model) / n < n) {
		fprintf(stderr, "andi: Comparison is limited to %zu sequences (%zu given).\n", (size_t)sqrt((double)(SIZE_MAX / sizeof(model))), n);
		exit(1);
	}
	model *M = (model *)malloc(n * n * sizeof(model));
	if (!M) {
		fprintf(stderr, "andi: Could not allocate enough memory for the comparison matrix. Try using --join or --low-memory.\n");
		exit(errno ? errno : 1);
	}
	const int flags = (&FLAGS != nullptr) ? FLAGS : 0;
	compat_matrix(M, sequences, n, (flags & 32) != 0);	// F_LOW_MEMORY
	print_distances(M, sequences, n, 1);
	if ((flags & 2) && print_coverages) print_coverages(M, n);	// F_VERBOSE
	// src/process.c:289-321: bootstrap matrices from the averaged cells, through the host's own sampler
	if (&BOOTSTRAP != nullptr && BOOTSTRAP && model_bootstrap) {
		model *B = (model *)malloc(n * n * sizeof(model));
		if (!B) {
			fprintf(stderr, "andi: Out of memory\n");
			exit(errno ? errno : 1);
		}
		while (BOOTSTRAP--) {
			for (size_t i = 0; i < n; i++) {
				for (size_t j = i; j < n; j++) {
					if (i == j) {
						memset(&B[i * n + j], 0, sizeof(model));
						B[i * n + j].seq_len = 1, B[i * n + j].counts[0] = 1;
						continue;
					}
					model datum = model_average(&M[i * n + j], &M[j * n + i]);
					datum = model_bootstrap(datum);
					B[j * n + i] = B[i * n + j] = datum;
				}
			}
			print_distances(B, sequences, n, 0);
		}
		free(B);
	}
	free(M);
}
