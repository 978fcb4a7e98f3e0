// andi_b200/csrc/compat.cuh -- part B of the C ABI (include/andi_compat.h); included by
// andi_b200.cu. The reference's own symbols (src/esa.h:61-64, src/process.c:141) as thin,
// mutex-serialised wrappers around one lazily created GPU context.
#pragma once
#include "../../include/andi_compat.h"

#include <mutex>
#include <unordered_map>

extern "C" int MODEL __attribute__((weak));	 // the reference's global (src/andi.c:50), if the host program has one

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
