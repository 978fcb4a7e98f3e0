// andi_b200/csrc/multi_host.cuh -- the matrix over several GPUs from one process; included by
// andi_b200.cu. The reference parallelises distMatrix over subjects with OpenMP threads
// (src/dist_hack.h:8,46-47); here the subjects go to devices: one host thread per GPU, the pool
// uploaded and packed once and handed on packed (peer copies over NVLink), subjects dispensed
// in small batches from an atomic queue, rows written into the caller's matrix.
#pragma once
#include <atomic>
#include <mutex>
#include <thread>

extern "C" int andi_pool_export(const andi_ctx *ctx, andi_pool_view *out) {
	if (!ctx || !out || !ctx->n) return ANDI_ERR_ARG;
	out->d_code = ctx->pool_code, out->d_spec = ctx->pool_spec, out->words = ctx->pool_words, out->n = ctx->n;
	out->lens = ctx->len.data(), out->gc = ctx->gc.data(), out->has_separator = ctx->has_sep.data();
	out->any_separator = ctx->any_sep ? 1 : 0;
	return ANDI_OK;
}

extern "C" int andi_pool_import(andi_ctx *ctx, const andi_pool_view *v, int src_device) {
	if (!ctx || !v || !v->d_code || !v->n || !v->lens || !v->gc || !v->has_separator) return ANDI_ERR_ARG;
	int rc = pool_check(ctx, v->lens, v->n);
	if (rc) return rc;
	CK(cudaSetDevice(ctx->device));
	// take copies of the facts first: the view may describe this very context
	const size_t n = v->n, words = v->words;
	std::vector<size_t> len(v->lens, v->lens + n);
	std::vector<double> gc(v->gc, v->gc + n);
	std::vector<int> has_sep(v->has_separator, v->has_separator + n);
	const void *src_code = v->d_code, *src_spec = v->d_spec;
	const bool any = v->any_separator != 0;
	if (src_code == ctx->pool_code) return ANDI_OK;	 // importing one's own pool
	pool_release(ctx);
	{
		int rc2 = planes_ensure(ctx, words);
		if (rc2) return rc2;
	}
	u64 *code = ctx->pool_code, *spec = ctx->pool_spec;
	const size_t bytes = words * sizeof(u64);
	if (src_device < 0 || src_device == ctx->device) {
		CK(cudaMemcpyAsync(code, src_code, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
		if (any && src_spec) CK(cudaMemcpyAsync(spec, src_spec, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
	} else {
		int can = 0;
		if (cudaDeviceCanAccessPeer(&can, ctx->device, src_device) == cudaSuccess && can) {
			cudaError_t e = cudaDeviceEnablePeerAccess(src_device, 0);
			if (e != cudaSuccess) cudaGetLastError();  // already enabled is fine; without it the copy is staged
		}
		CK(cudaMemcpyPeerAsync(code, ctx->device, src_code, src_device, bytes, ctx->stream));
		if (any && src_spec) CK(cudaMemcpyPeerAsync(spec, ctx->device, src_spec, src_device, bytes, ctx->stream));
	}
	if (!(any && src_spec)) CK(cudaMemsetAsync(spec, 0, bytes, ctx->stream));  // a pool without separators: nothing to move
	CK(cudaStreamSynchronize(ctx->stream));
	ctx->n = n, ctx->len = len, ctx->gc = gc, ctx->has_sep = has_sep, ctx->any_sep = any;
	ctx->word_off.resize(n);
	std::vector<QueryView> qv(n);
	size_t w = 0;
	for (size_t k = 0; k < n; k++) {
		ctx->word_off[k] = w;
		w += (plane_words(len[k]) + 1) & ~(size_t)1;
		qv[k].t.code = code + ctx->word_off[k], qv[k].t.spec = spec + ctx->word_off[k];
		qv[k].t.len = (u32)len[k], qv[k].t.mid = 0xffffffffu, qv[k].has_sep = has_sep[k];
	}
	if (w != words) {
		ctx->err = "pool view does not match its sequence lengths";
		pool_release(ctx);
		return ANDI_ERR_ARG;
	}
	CK(dalloc(ctx, &ctx->d_queries, n));
	CK(cudaMemcpyAsync(ctx->d_queries, qv.data(), n * sizeof(QueryView), cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	ctx->st.p2p_bytes += bytes * ((any && src_spec) ? 2 : 1);
	return ANDI_OK;
}

extern "C" int andi_dist_matrix_multi(const int *devices, int nd, const char *const *seqs, const size_t *lens, size_t n,
									  double p_value, int model, int low_memory, andi_model *out, andi_progress_fn progress,
									  void *user, char *errbuf, size_t errbuf_len) {
	auto fail = [&](int rc, const char *msg) {
		if (errbuf && errbuf_len) snprintf(errbuf, errbuf_len, "%s", msg ? msg : "");
		return rc;
	};
	if (!devices || nd < 1 || !seqs || !lens || n == 0 || !out) return fail(ANDI_ERR_ARG, "bad argument");
	std::vector<andi_ctx *> ctx((size_t)nd, nullptr);
	auto destroy_all = [&]() {
		for (auto c : ctx) andi_ctx_destroy(c);
	};
	for (int d = 0; d < nd; d++) {
		int rc = andi_ctx_create(devices[d], nullptr, &ctx[(size_t)d]);
		if (rc) {
			std::string m = andi_last_error(nullptr);
			destroy_all();
			return fail(rc, m.c_str());
		}
	}
	// the pool: one upload, one pack (src/sequence.c:196-207,260-282 on devices[0])
	int rc = andi_pool_set_host(ctx[0], seqs, lens, n);
	if (rc) {
		std::string m = andi_last_error(ctx[0]);
		destroy_all();
		return fail(rc, m.c_str());
	}
	andi_pool_view view;
	andi_pool_export(ctx[0], &view);

	// subjects in batches: small enough that every device gets many (the tail of the run is one
	// batch long), large enough that the per-call row download does not matter
	size_t batch = n / ((size_t)nd * 24);
	batch = std::max<size_t>(1, std::min<size_t>(batch, 16));
	std::atomic<size_t> next{0};
	std::atomic<int> first_rc{ANDI_OK};
	std::mutex mu;
	std::string first_msg;
	size_t pairs_done = 0;
	const size_t pairs_total = n * n - n;
	auto worker = [&](int d) {
		andi_ctx *c = ctx[(size_t)d];
		int r = d == 0 ? ANDI_OK : andi_pool_import(c, &view, devices[0]);
		while (!r && first_rc.load() == ANDI_OK) {
			size_t i = next.fetch_add(batch);
			if (i >= n) break;
			size_t e = std::min(n, i + batch);
			r = andi_dist_rows(c, i, e, p_value, model, low_memory, out + i * n);
			if (!r) {
				std::lock_guard<std::mutex> lock(mu);
				pairs_done += (e - i) * (n - 1);
				if (progress) progress(pairs_done, pairs_total, user);
			}
		}
		if (r) {
			std::lock_guard<std::mutex> lock(mu);
			int expected = ANDI_OK;
			if (first_rc.compare_exchange_strong(expected, r)) first_msg = andi_last_error(c);
		}
	};
	std::vector<std::thread> threads;
	for (int d = 1; d < nd; d++) threads.emplace_back(worker, d);
	worker(0);
	for (auto &t : threads) t.join();
	destroy_all();
	if (first_rc.load()) return fail(first_rc.load(), first_msg.c_str());
	return fail(ANDI_OK, "");
}
