// andi_b200/csrc/walk_host.cuh -- host side of the anchor walk; included by andi_b200.cu.
//
// distMatrix / distMatrixLM (src/dist_hack.h:34-96): for every subject build its index, walk
// every query against it. One subject is resident at a time (its index lives in L2 while the
// queries stream), which is also what F_LOW_MEMORY asks for; the parallelism comes from
// splitting every query into chunks (walk_kernels.cuh).
#pragma once
#include "walk_fast.cuh"
#include "walk_v3.cuh"

typedef void (*chunks_fn)(const SubjectIndex, const QueryView *, const u32 *, u32, u32, u32, u32, u32 *,
						  unsigned long long *);
typedef void (*reduce_fn)(const SubjectIndex, const QueryView *, const u32 *, u32, u32, u32, u32, const u32 *, const u32 *, u32 *, u32);

static void pick_walk(int model, bool spec, chunks_fn &cf, reduce_fn &rf) {
	bool quarter = model == ANDI_M_RAW || model == ANDI_M_JC || model == ANDI_M_KIMURA;
	if (quarter && spec) cf = k_walk_chunks<true, true>, rf = k_walk_reduce_finish<true, true>;
	if (quarter && !spec) cf = k_walk_chunks<true, false>, rf = k_walk_reduce_finish<true, false>;
	if (!quarter && spec) cf = k_walk_chunks<false, true>, rf = k_walk_reduce_finish<false, true>;
	if (!quarter && !spec) cf = k_walk_chunks<false, false>, rf = k_walk_reduce_finish<false, false>;
}

// Chunk length: aim at ~5 units per resident thread so the dynamic unit queue balances, keep chunks
// long enough that the boundary replay (a handful of steps) and the per-unit record stay a small
// fraction.
static u32 pick_chunk(const andi_ctx *ctx, unsigned long long total_bases) {
	unsigned long long target_units = (unsigned long long)ctx->sm_count * 1024ULL * 11ULL / 2ULL;
	unsigned long long want = total_bases / target_units;
	// measured on the C4 shape, round 1 (phase pipeline): 2048 -> 310 k, 4096 -> 341 k, 5120 -> 345 k, 8192 -> 334 k pairs/s;
	// round 2 (k_walk_v3, two subjects in flight): 4096 -> 519 k, 5632 -> 536 k, 8192 -> 546 k, 11264 -> 544 k
	u32 chunk = (u32)std::min<unsigned long long>(16384ULL, std::max<unsigned long long>(1024ULL, (want + 511ULL) / 512ULL * 512ULL));
	const char *env = getenv("ANDI_B200_CHUNK");
	if (env && atoi(env) >= 64) chunk = (u32)atoi(env);
	return chunk;
}

struct WalkPlan {
	u32 chunk = 0, cpq = 0;
	size_t record_words = 0;
};

static WalkPlan plan_walk(const andi_ctx *ctx, const std::vector<size_t> &qlens) {
	WalkPlan p;
	unsigned long long total = 0;
	size_t maxlen = 0;
	for (size_t l : qlens) total += l, maxlen = std::max(maxlen, l);
	p.chunk = pick_chunk(ctx, total);
	p.cpq = (u32)((maxlen + p.chunk - 1) / p.chunk);
	if (p.cpq == 0) p.cpq = 1;
	p.record_words = qlens.size() * (size_t)p.cpq * ANDI_UNIT_WORDS;
	return p;
}

// The pool's prefix-composition table (k_comp_prefix), built on first use.
static int pool_comp_ensure(andi_ctx *ctx) {
	if (ctx->pool_comp || !ctx->n) return ANDI_OK;
	CK(dalloc(ctx, &ctx->pool_comp, ctx->pool_words));
	k_comp_prefix<<<(unsigned)ctx->n, 256, 0, ctx->stream>>>(ctx->d_queries, ctx->pool_code, ctx->pool_comp);
	ctx->st.esa_launches++;
	CK(cudaGetLastError());
	return ANDI_OK;
}

// The pool's separator hints (k_sep3), built on first use.
static int pool_sep3_ensure(andi_ctx *ctx) {
	if (ctx->pool_sep3 || !ctx->n) return ANDI_OK;
	CK(dalloc(ctx, &ctx->pool_sep3, ctx->pool_words));
	k_sep3<<<nblocks(ctx->pool_words, 256), 256, 0, ctx->stream>>>(ctx->pool_spec, ctx->pool_words, ctx->pool_sep3);
	ctx->st.esa_launches++;
	CK(cudaGetLastError());
	return ANDI_OK;
}

// Walk nq queries against one index; d_out gets nq cells of 17 words. `pool_queries`: the
// query views are the pool's own (ctx->d_queries), so the prefix-composition table applies.
static int launch_walk(andi_ctx *ctx, SubjectIndex S, const QueryView *d_queries, const u32 *d_query_ids,
					   u32 nq, const WalkPlan &plan, u32 threshold, int model, bool spec, bool pool_queries,
					   u32 *d_records, u32 *d_out) {
	chunks_fn cf = nullptr;
	reduce_fn rf = nullptr;
	pick_walk(model, spec, cf, rf);
	// the phase-pipeline kernels (headline: RAW/JC/KIMURA counting without separators; LOGDET/ANI
	// need the composition table, which only pool queries have)
	const char *force = getenv("ANDI_B200_WALK");
	bool quarter = model == ANDI_M_RAW || model == ANDI_M_JC || model == ANDI_M_KIMURA;
	if (!(force && strcmp(force, "basic") == 0)) {
		if (spec && pool_queries) {
			int rc = pool_sep3_ensure(ctx);
			if (rc) return rc;
			S.qcode_base = ctx->pool_code, S.qsep3_base = ctx->pool_sep3;
		}
		if (spec && !S.s_sep3) {
			ctx->err = "index without separator hints";
			return ANDI_ERR_ARG;
		}
		if (quarter) {
			cf = spec ? k_walk_chunks_fast<true, true> : k_walk_chunks_fast<true, false>;
		} else if (pool_queries) {
			int rc = pool_comp_ensure(ctx);
			if (rc) return rc;
			S.qcode_base = ctx->pool_code, S.qcomp_base = ctx->pool_comp;
			cf = spec ? k_walk_chunks_fast<false, true> : k_walk_chunks_fast<false, false>;
		}
	}
	// the headline configuration goes through k_walk_v3 (walk_v3.cuh): PHASE 1 over all units,
	// PHASE 2 over all chunk boundaries; ANDI_B200_WALK=pipeline keeps the round-1 kernel
	// (LOGDET / ANI: only pool queries have the prefix-composition table the anchor interiors need)
	const bool v3 = !spec && (quarter || (pool_queries && S.qcomp_base)) && v3_applies(S, threshold) &&
					!(force && (strcmp(force, "basic") == 0 || strcmp(force, "pipeline") == 0));
	int per_sm = 0;
	const unsigned threads = v3 ? V3_THREADS : ANDI_WALK_THREADS;
	if (v3)
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, quarter ? k_walk_v3<1, true, false> : k_walk_v3<1, false, false>, V3_THREADS, 0));
	else
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cf, ANDI_WALK_THREADS, 0));
	if (per_sm < 1) per_sm = 1;
	unsigned long long units = (unsigned long long)nq * plan.cpq;
	unsigned grid = (unsigned)std::min<unsigned long long>((units + threads - 1) / threads, (unsigned long long)per_sm * ctx->sm_count);
	cudaEvent_t e0 = get_event(ctx), e1 = get_event(ctx);
	mark(ctx, e0);
	if (!ctx->first_ev) {
		ctx->first_ev = get_event(ctx);
		mark(ctx, ctx->first_ev);
	}
	if (!ctx->walk_counter) CK(dalloc(ctx, &ctx->walk_counter, 2));
	CK(cudaMemsetAsync(ctx->walk_counter, 0, 2 * sizeof(unsigned long long), ctx->stream));
	if (v3) {
		// The mean length of the anchors the chunk walks of this lane's previous launch ended with (the
		// copy has arrived: the index build has synchronised the stream since): kilobases = a pool of
		// near-identical genomes -> the instantiation with EXT bursts; a few dozen bases -> without.
		if (ctx->h_walk_stat && ctx->h_walk_stat[1]) {
			const unsigned long long mean = ctx->h_walk_stat[0] / ctx->h_walk_stat[1];
			if (mean > 192ULL) ctx->walk_burst = true;
			if (mean < 128ULL) ctx->walk_burst = false;
		}
		if (const char *fb = getenv("ANDI_B200_BURST")) ctx->walk_burst = atoi(fb) != 0;  // experiments only
		const bool burst = ctx->walk_burst;
		auto p1 = quarter ? (burst ? k_walk_v3<1, true, true> : k_walk_v3<1, true, false>) : (burst ? k_walk_v3<1, false, true> : k_walk_v3<1, false, false>);
		auto p2 = quarter ? (burst ? k_walk_v3<2, true, true> : k_walk_v3<2, true, false>) : (burst ? k_walk_v3<2, false, true> : k_walk_v3<2, false, false>);
		p1<<<grid, V3_THREADS, 0, ctx->stream>>>(S, d_queries, d_query_ids, nq, plan.chunk, plan.cpq, threshold, d_records, ctx->walk_counter);
		if (plan.cpq > 1)
			p2<<<grid, V3_THREADS, 0, ctx->stream>>>(S, d_queries, d_query_ids, nq, plan.chunk, plan.cpq, threshold, d_records,
													 ctx->walk_counter + 1);
		ctx->st.walk_launches += plan.cpq > 1 ? 1 : 0;
	} else {
		cf<<<grid, ANDI_WALK_THREADS, 0, ctx->stream>>>(S, d_queries, d_query_ids, nq, plan.chunk, plan.cpq, threshold,
														 d_records, ctx->walk_counter);
	}
	// the records of every pair: summed by a grid of (slice, pair) CTAs, finished (tail, diagonal cell,
	// pairs with a boundary that did not synchronise) by one warp per pair
	if (nq > ctx->walk_bad_cap) {
		dfree(ctx, ctx->walk_bad);
		ctx->walk_bad_cap = 0;
		CK(dalloc(ctx, &ctx->walk_bad, nq));
		ctx->walk_bad_cap = nq;
	}
	CK(cudaMemsetAsync(ctx->walk_bad, 0, (size_t)nq * sizeof(u32), ctx->stream));
	if (!ctx->walk_stat) {
		CK(dalloc(ctx, &ctx->walk_stat, 2));
		CK(cudaHostAlloc((void **)&ctx->h_walk_stat, 2 * sizeof(unsigned long long), cudaHostAllocDefault));
		ctx->h_walk_stat[0] = ctx->h_walk_stat[1] = 0;
	}
	CK(cudaMemsetAsync(ctx->walk_stat, 0, 2 * sizeof(unsigned long long), ctx->stream));
	CK(cudaMemsetAsync(d_out, 0, (size_t)nq * 17 * sizeof(u32), ctx->stream));
	k_walk_reduce_sum<<<dim3(nq, nblocks(plan.cpq, ANDI_REDUCE_SLICE)), 256, 0, ctx->stream>>>(d_queries, d_query_ids, S.self, plan.chunk,
																								plan.cpq, d_records, d_out, ctx->walk_bad, v3 ? ctx->walk_stat : nullptr);
	if (v3) CK(cudaMemcpyAsync(ctx->h_walk_stat, ctx->walk_stat, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
	rf<<<nblocks((size_t)nq * 32, 128), 128, 0, ctx->stream>>>(S, d_queries, d_query_ids, nq, plan.chunk, plan.cpq, threshold, d_records,
																ctx->walk_bad, d_out, v3 ? 1u : 0u);
	ctx->st.walk_launches += 1;
	static const bool debug_bad = getenv("ANDI_B200_DEBUG_BAD") != nullptr;  // experiments: which boundaries did not synchronise
	if (debug_bad) {
		std::vector<u32> bad(nq);
		CK(cudaMemcpyAsync(bad.data(), ctx->walk_bad, (size_t)nq * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
		u32 nbad = 0, first = nq;
		for (u32 k = 0; k < nq; k++)
			if (bad[k]) nbad++, first = first == nq ? k : first;
		fprintf(stderr, "[andi_b200] subject %u: %u of %u pairs with a boundary whose chains did not meet (k_walk_v3: before the end of the query; else: inside the next chunk)\n", S.self, nbad, nq);
		if (first < nq) {
			std::vector<u32> rec((size_t)plan.cpq * ANDI_UNIT_WORDS);
			CK(cudaMemcpy(rec.data(), d_records + (size_t)first * plan.cpq * ANDI_UNIT_WORDS, rec.size() * sizeof(u32), cudaMemcpyDeviceToHost));
			for (u32 ch = 0; ch + 1 < plan.cpq; ch++) {
				const u32 *r = &rec[(size_t)ch * ANDI_UNIT_WORDS], *n = r + ANDI_UNIT_WORDS;
				if (!r[37])
					fprintf(stderr, "  pair %u boundary after chunk %u (at %llu): E = pos %u, last (s %u, q %u, len %u), paired %u; next chunk's E = pos %u, last (s %u, q %u, len %u)\n",
							first, ch, (unsigned long long)(ch + 1) * plan.chunk, r[32], r[33], r[34], r[35], r[36], n[32], n[33], n[34], n[35]);
			}
		}
	}
	mark(ctx, e1);
	ctx->walk_ev.emplace_back(e0, e1);
	if (!ctx->last_ev) ctx->last_ev = get_event(ctx);
	mark(ctx, ctx->last_ev);
	ctx->st.walk_launches += 2;
	ctx->st.pairs += (!d_query_ids && S.self < nq) ? nq - 1 : nq;  // the subject's own cell is not a walk
	CK(cudaGetLastError());
	return ANDI_OK;
}

extern "C" int andi_dist_row(andi_ctx *ctx, const andi_esa *E, const size_t *query_ids, size_t nq,
							 size_t threshold, int model, andi_model *out) {
	if (!ctx || !E || !query_ids || !out || E->ctx != ctx) return ANDI_ERR_ARG;
	if (nq == 0) return ANDI_OK;
	CK(cudaSetDevice(ctx->device));
	std::vector<u32> ids(nq);
	std::vector<size_t> qlens(nq);
	bool spec = E->has_sep;
	for (size_t k = 0; k < nq; k++) {
		if (query_ids[k] >= ctx->n) return ANDI_ERR_ARG;
		ids[k] = (u32)query_ids[k];
		qlens[k] = ctx->len[ids[k]];
		spec |= ctx->has_sep[ids[k]] != 0;
	}
	SubjectIndex S = subject_index(E);
	S.self = 0xffffffffu;				   // dist_anchor itself has no notion of "self"
	S.qspec_delta = ctx->pool_spec - ctx->pool_code;
	if (threshold < (size_t)S.K) S.K = 0;  // the directory assumes K <= threshold
	WalkPlan plan = plan_walk(ctx, qlens);
	u32 *d_ids = nullptr, *d_out = nullptr, *d_rec = nullptr;
	CK(dalloc(ctx, &d_ids, nq));
	CK(dalloc(ctx, &d_out, nq * 17));
	CK(dalloc(ctx, &d_rec, plan.record_words));
	CK(cudaMemcpyAsync(d_ids, ids.data(), nq * 4, cudaMemcpyHostToDevice, ctx->stream));
	int rc = launch_walk(ctx, S, ctx->d_queries, d_ids, (u32)nq, plan, (u32)threshold, model, spec, true, d_rec, d_out);
	if (!rc) {
		CK(cudaMemcpyAsync(out, d_out, nq * sizeof(andi_model), cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
		ctx->st.d2h_bytes += nq * sizeof(andi_model);
		harvest_events(ctx);
	}
	dfree(ctx, d_ids), dfree(ctx, d_out), dfree(ctx, d_rec);
	return rc;
}

extern "C" int andi_dist_anchor(andi_ctx *ctx, const andi_esa *E, const char *query, size_t qlen,
								size_t threshold, int model, andi_model *out) {
	if (!ctx || !E || !query || !out || E->ctx != ctx) return ANDI_ERR_ARG;
	CK(cudaSetDevice(ctx->device));
	if (qlen == 0) {
		memset(out, 0, sizeof *out);
		return ANDI_OK;
	}
	if (qlen > (size_t)INT_MAX) return ANDI_ERR_TOO_LONG;
	TempQueries T;
	int rc = temp_queries(ctx, &query, &qlen, 1, T);
	if (rc) return rc;
	SubjectIndex S = subject_index(E);
	S.self = 0xffffffffu;
	S.qspec_delta = T.spec - T.code;
	S.qcode_base = T.code, S.qsep3_base = T.sep3;
	if (threshold < (size_t)S.K) S.K = 0;
	WalkPlan plan = plan_walk(ctx, std::vector<size_t>{qlen});
	u32 *d_out = nullptr, *d_rec = nullptr;
	CK(dalloc(ctx, &d_out, 17));
	CK(dalloc(ctx, &d_rec, plan.record_words));
	rc = launch_walk(ctx, S, T.d_views, nullptr, 1, plan, (u32)threshold, model, E->has_sep || T.any_sep, false, d_rec, d_out);
	if (!rc) {
		CK(cudaMemcpyAsync(out, d_out, sizeof(andi_model), cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
		ctx->st.d2h_bytes += sizeof(andi_model);
		harvest_events(ctx);
	}
	dfree(ctx, d_out), dfree(ctx, d_rec);
	temp_release(ctx, T);
	return rc;
}

// The helper lane of a context: same device, own stream, the owner's pool by reference.
static int helper_ensure(andi_ctx *ctx) {
	if (!ctx->helper) {
		andi_ctx *h = nullptr;
		int rc = andi_ctx_create(ctx->device, nullptr, &h);
		if (rc) {
			ctx->err = std::string("helper lane: ") + andi_last_error(nullptr);
			return rc;
		}
		ctx->helper = h;
		CK(cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&ctx->join_ev, cudaEventDisableTiming));
	}
	andi_ctx *h = ctx->helper;
	h->n = ctx->n, h->len = ctx->len, h->gc = ctx->gc, h->has_sep = ctx->has_sep, h->word_off = ctx->word_off;
	h->pool_code = ctx->pool_code, h->pool_spec = ctx->pool_spec, h->pool_words = ctx->pool_words;
	h->pool_comp = ctx->pool_comp, h->pool_sep3 = ctx->pool_sep3, h->d_queries = ctx->d_queries, h->any_sep = ctx->any_sep;
	h->borrowed_pool = true;
	return ANDI_OK;
}

static void merge_stats(andi_stats &a, const andi_stats &b) {
	a.esa_ms += b.esa_ms, a.walk_ms += b.walk_ms;
	a.esa_launches += b.esa_launches, a.cub_calls += b.cub_calls, a.walk_launches += b.walk_launches;
	a.pairs += b.pairs, a.subjects += b.subjects, a.sa_rounds += b.sa_rounds;
	a.h2d_bytes += b.h2d_bytes, a.d2h_bytes += b.d2h_bytes, a.p2p_bytes += b.p2p_bytes;
}

// Two subjects are in flight, on two streams ("lanes"): while the walk of subject i drains its last
// units -- a persistent grid whose tail leaves SMs idle -- the index build and the first walk
// blocks of subject i+1 take the free SMs. Every lane has its own index arrays, build scratch and
// unit records (lane 1 is a helper context that borrows the pool); within a lane everything is
// stream-ordered, so the only cross-lane synchronisation is fork / join around the whole call.
// ANDI_B200_LANES=1 keeps everything on the context's own stream (used to time kernels alone).
static int dist_rows_impl(andi_ctx *ctx, size_t s_begin, size_t s_end, double p_value, int model, int low_memory,
						  andi_model *out, bool out_on_device) {
	(void)low_memory;  // one index is resident per lane in either mode (src/dist_hack.h:14-16)
	if (!ctx || !out || s_begin > s_end || s_end > ctx->n) return ANDI_ERR_ARG;
	if (s_begin == s_end) return ANDI_OK;
	CK(cudaSetDevice(ctx->device));
	const size_t n = ctx->n, rows = s_end - s_begin;
	WalkPlan plan = plan_walk(ctx, ctx->len);
	unsigned long long total_bases = 0;
	for (size_t l : ctx->len) total_bases += l;
	const char *lanes_env = getenv("ANDI_B200_LANES");
	int nl = (lanes_env && atoi(lanes_env) == 1) || rows < 2 ? 1 : 2;
	u32 *d_out = nullptr;
	if (out_on_device)
		d_out = (u32 *)out;
	else
		CK(dalloc(ctx, &d_out, rows * n * 17));
	const bool quarter = model == ANDI_M_RAW || model == ANDI_M_JC || model == ANDI_M_KIMURA;
	int rc = ANDI_OK;
	if (nl == 2) {
		// the lazily built pool tables must exist before the helper borrows the pool
		if (!quarter) rc = pool_comp_ensure(ctx);
		if (!rc && ctx->any_sep) rc = pool_sep3_ensure(ctx);
		if (!rc) rc = helper_ensure(ctx);
		if (rc) nl = 1, rc = ANDI_OK;  // no second lane: carry on with one
	}
	andi_ctx *lane[2] = {ctx, nl == 2 ? ctx->helper : nullptr};
	if (nl == 2) {
		CK(cudaEventRecord(ctx->fork_ev, ctx->stream));
		CK(cudaStreamWaitEvent(lane[1]->stream, ctx->fork_ev, 0));
	}
	cudaEvent_t e_begin = get_event(ctx), e_end = get_event(ctx);
	mark(ctx, e_begin);
	u32 *d_rec[2] = {nullptr, nullptr};
	andi_esa E[2];	// rebuilt in place for every subject of the lane: the arrays are allocated once
	for (int x = 0; x < nl && !rc; x++) {
		E[x].ctx = lane[x];
		if (dalloc(lane[x], &d_rec[x], plan.record_words) != cudaSuccess) {
			ctx->err = "device allocation failed (walk records)";
			rc = ANDI_ERR_NOMEM;
		}
	}
	for (size_t i = s_begin; i < s_end && !rc; i++) {
		const int x = (int)((i - s_begin) % (size_t)nl);
		andi_ctx *L = lane[x];
		andi_esa &Ex = E[x];
		Ex.n = (u32)ctx->len[i];
		Ex.N = 2 * Ex.n + 1;
		Ex.has_sep = ctx->has_sep[i] != 0;
		Ex.self = (u32)i;
		Ex.threshold = (u32)andi_threshold(p_value, ctx->gc[i], Ex.N);
		Ex.K = choose_depth(Ex.N, Ex.threshold, total_bases);
		size_t nw = plane_words(Ex.N);
		rc = esa_ensure(L, &Ex);
		if (!rc) {
			k_build_rs<<<nblocks(nw, 256), 256, 0, L->stream>>>(ctx->pool_code + ctx->word_off[i], ctx->pool_spec + ctx->word_off[i],
																 Ex.n, Ex.code, Ex.spec, (u32)nw);
			L->st.esa_launches++;
			rc = build_index(L, &Ex, ANDI_ESA_SEARCH);
		}
		if (!rc) {
			SubjectIndex S = subject_index(&Ex);
			S.qspec_delta = ctx->pool_spec - ctx->pool_code;
			rc = launch_walk(L, S, ctx->d_queries, nullptr, (u32)n, plan, Ex.threshold, model, ctx->any_sep || Ex.has_sep, true,
							 d_rec[x], d_out + (i - s_begin) * n * 17);
		}
		if (rc && L != ctx) ctx->err = L->err;
	}
	for (int x = 0; x < nl; x++) {
		esa_release(&E[x]);
		dfree(lane[x], d_rec[x]);
	}
	if (nl == 2) {
		cudaEventRecord(ctx->join_ev, lane[1]->stream);
		cudaStreamWaitEvent(ctx->stream, ctx->join_ev, 0);
	}
	mark(ctx, e_end);
	if (!rc && !out_on_device) {
		CK(cudaMemcpyAsync(out, d_out, rows * n * sizeof(andi_model), cudaMemcpyDeviceToHost, ctx->stream));
		ctx->st.d2h_bytes += rows * n * sizeof(andi_model);
	}
	cudaError_t e = cudaStreamSynchronize(ctx->stream);
	if (nl == 2 && e == cudaSuccess) e = cudaStreamSynchronize(lane[1]->stream);
	if (!rc && e != cudaSuccess) {
		ctx->err = std::string("walk: ") + cudaGetErrorString(e);
		rc = ANDI_ERR_CUDA;
	}
	harvest_events(ctx);
	if (nl == 2) {
		harvest_events(lane[1]);
		merge_stats(ctx->st, lane[1]->st);
		lane[1]->st = andi_stats{};
	}
	// wall time of the whole call on the device (the per-kernel sums above overlap across lanes)
	float ms = 0.f;
	if (cudaEventElapsedTime(&ms, e_begin, e_end) == cudaSuccess) ctx->st.rows_ms += ms;
	ctx->free_ev.push_back(e_begin), ctx->free_ev.push_back(e_end);
	if (!out_on_device) dfree(ctx, d_out);
	return rc;
}

extern "C" int andi_dist_rows(andi_ctx *ctx, size_t s_begin, size_t s_end, double p_value, int model,
							  int low_memory, andi_model *out) {
	return dist_rows_impl(ctx, s_begin, s_end, p_value, model, low_memory, out, false);
}

extern "C" int andi_dist_rows_device(andi_ctx *ctx, size_t s_begin, size_t s_end, double p_value, int model,
									 int low_memory, andi_model *d_out) {
	return dist_rows_impl(ctx, s_begin, s_end, p_value, model, low_memory, d_out, true);
}
