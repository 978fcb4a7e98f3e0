// andi_b200/csrc/walk_v3.cuh -- k_walk_v3<PHASE, QUARTER, BURST>: the chunked anchor walk for texts without
// separators (QUARTER = RAW / JC / KIMURA counting, the headline configuration; else LOGDET / ANI). The per-lane logic, its rationale and
// its reference citations are in walk_v3_lane.h (the same text runs in the CPU emulation of
// emu/emu_v3.cpp); this file supplies the device primitives, the warp loop and the launch.
//
//   PHASE 1  every (query, chunk) unit from the amnesic state: record words [0,16) U_c, [32,37) E_c
//   PHASE 2  every chunk boundary: the true chain (from E_c) against the amnesic walk of the next
//            chunk until both are in the same state: record words [16,32) D_c, [37] flag
// k_walk_reduce (walk_kernels.cuh) then sums the records exactly as it does for k_walk_chunks.
#pragma once
#include "walk_kernels.cuh"

#define V3_FN __device__ __forceinline__
#ifndef V3_THREADS
#define V3_THREADS 256
#endif
#define V3_CELL_STRIDE V3_THREADS
#ifndef V3_SERVE_BATCH
#define V3_SERVE_BATCH 6u  // parked lanes that make a warp stop and serve them ...
#endif
#ifndef V3_SERVE_EVERY
#define V3_SERVE_EVERY 8u  // ... or every so many trips (a power of two)
#endif
#ifndef V3_BLOCKS_PER_SM
#define V3_BLOCKS_PER_SM 4
#endif
#define V3_STAT(name) \
	do {              \
	} while (0)

V3_FN u32 v3_ctz64(u64 x) { return (u32)(__ffsll((long long)x) - 1); }
V3_FN u32 v3_ctz32(u32 x) { return (u32)__clz((int)__brev(x)); }  // 32 for x == 0
V3_FN u32 v3_popc32(u32 x) { return (u32)__popc(x); }
V3_FN u64 v3_ld_fdir(const u64 *p) { return __ldg(p); }
V3_FN u32 v3_ld_sa(const u32 *p) { return __ldg(p); }
V3_FN void v3_window64(const u64 *__restrict__ w, u32 pos, u64 &lo, u64 &hi) { window64(w, pos, lo, hi); }
V3_FN u32 v3_kmer_key(u64 win, int k) { return kmer_key(win, k); }

struct V3Lane;
struct V3Const;
V3_FN void v3_count_slice(const V3Lane &L, const V3Const &c, u32 *col, u32 sign);

#define V3_COOP_IN_WARP_LOOP 1  // requests V3_SVC_COOP / V3_SVC_COOP2 are served by v3_coop_scan below, not by v3_service
#include "walk_v3_lane.h"

// v3_scan_full (walk_v3_lane.h) by the whole warp: the requests of the lanes in `want`; every lane
// takes one candidate (requests of up to eight candidates: four requests at a time, eight lanes each;
// larger buckets: one request at a time, 32 candidates per round) and all candidates grow together,
// one 64-column window per iteration, until the last of them has met its first mismatch or its
// limit. Called with all 32 lanes.
V3_FN void v3_coop_scan(V3Lane &L, const V3Const &c, unsigned want) {
	const u32 lane = threadIdx.x & 31u;
	// Requests with at most eight candidates (most buckets of three or more, every tag-2 pair) are
	// served FOUR AT A TIME, eight lanes each: where such buckets are common (5 Mbp genomes at K = 12,
	// divergent pairs with many lookups) several lanes of a warp ask at every service stop, and one
	// request after the other cost more than it saved there.
	unsigned small = __ballot_sync(0xffffffffu, ((want >> lane) & 1u) && (L.svc == V3_SVC_COOP2 || L.cand2 <= 8u));
	want &= ~small;
	while (small) {
		const u32 g = lane >> 3, sub = lane & 7u;
		const u32 r = __fns(small, 0u, (int)g + 1);	 // the request lane my group works for (0xffffffff: none)
		const bool has = r != 0xffffffffu;
		const u32 src = has ? r : 0u;
		const bool two = __shfl_sync(0xffffffffu, L.svc, src) == V3_SVC_COOP2;
		const u32 a = __shfl_sync(0xffffffffu, L.cand_p, src), b = __shfl_sync(0xffffffffu, L.cand2, src);
		const u32 pos = __shfl_sync(0xffffffffu, L.pos, src), qlen = __shfl_sync(0xffffffffu, L.qlen, src);
		const u64 *q_code = reinterpret_cast<const u64 *>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(L.q_code), src));
		const bool have = has && sub < (two ? 2u : b);
		u32 p = 0;
		if (have) p = two ? (sub ? b : a) : __ldg(c.SA + a + sub);
		const u32 run = p < c.mid ? c.mid - p : (p == c.mid ? 0u : c.N - p), rem = qlen - pos;
		const u32 lim = have ? (rem < run ? rem : run) : 0u;
		u32 len = 0;
		bool alive = len < lim;
		while (__any_sync(0xffffffffu, alive)) {
			if (alive) {
				u64 q0, q1, s0, s1;
				window64(q_code, pos + len, q0, q1);
				window64(c.s_code, p + len, s0, s1);
				const u32 D = v3_first_diff(q0 ^ s0, q1 ^ s1);
				len += D < lim - len ? D : lim - len;
				alive = D >= 64u && len < lim;
			}
		}
		u32 top = have ? len : 0u;
		top = max(top, __shfl_xor_sync(0xffffffffu, top, 1));
		top = max(top, __shfl_xor_sync(0xffffffffu, top, 2));
		top = max(top, __shfl_xor_sync(0xffffffffu, top, 4));
		const unsigned at_top = __ballot_sync(0xffffffffu, have && len == top) & (0xffu << (8u * g));
		const u32 n = (u32)__popc(at_top), p_top = __shfl_sync(0xffffffffu, p, at_top ? __ffs((int)at_top) - 1 : 0);
		// the result travels from lane 0 of the group to the lane that asked
		const u32 my_g = (u32)__popc(small & ((1u << lane) - 1u));
		const bool served = ((small >> lane) & 1u) && my_g < 4u;
		const u32 from = served ? 8u * my_g : 0u;
		const u32 r_top = __shfl_sync(0xffffffffu, top, from), r_n = __shfl_sync(0xffffffffu, n, from), r_p = __shfl_sync(0xffffffffu, p_top, from);
		if (served) L.cand_p = r_p, L.len1 = r_top, L.cand2 = r_n == 1u ? 1u : 0u, L.job = V3_RESOLVED, L.svc = V3_RUN;
		for (int k = 0; k < 4 && small; k++) small &= small - 1u;
	}
	while (want) {
		const int r = __ffs((int)want) - 1;
		want &= want - 1u;
		const bool two = __shfl_sync(0xffffffffu, L.svc, r) == V3_SVC_COOP2;
		const u32 a = __shfl_sync(0xffffffffu, L.cand_p, r), b = __shfl_sync(0xffffffffu, L.cand2, r);
		const u32 pos = __shfl_sync(0xffffffffu, L.pos, r), qlen = __shfl_sync(0xffffffffu, L.qlen, r);
		const u64 *q_code = reinterpret_cast<const u64 *>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(L.q_code), r));
		const u32 count = two ? 2u : b, rem = qlen - pos;
		u32 best = 0, best_p = 0, best_n = 0;
		for (u32 base = 0; base < count; base += 32u) {
			const u32 k = base + lane;
			const bool have = k < count;
			u32 p = 0;
			if (have) p = two ? (k ? b : a) : __ldg(c.SA + a + k);
			const u32 run = p < c.mid ? c.mid - p : (p == c.mid ? 0u : c.N - p);
			const u32 lim = have ? (rem < run ? rem : run) : 0u;
			u32 len = 0;
			bool alive = len < lim;
			while (__any_sync(0xffffffffu, alive)) {
				if (alive) {
					u64 q0, q1, s0, s1;
					window64(q_code, pos + len, q0, q1);
					window64(c.s_code, p + len, s0, s1);
					const u32 D = v3_first_diff(q0 ^ s0, q1 ^ s1);
					len += D < lim - len ? D : lim - len;
					alive = D >= 64u && len < lim;
				}
			}
			const u32 top = __reduce_max_sync(0xffffffffu, have ? len : 0u);
			const unsigned at_top = __ballot_sync(0xffffffffu, have && len == top);
			const u32 n = (u32)__popc(at_top), p_top = __shfl_sync(0xffffffffu, p, __ffs((int)at_top) - 1);
			best_n = top > best ? n : (top == best ? best_n + n : best_n);
			best_p = top > best ? p_top : best_p;
			best = top > best ? top : best;
		}
		if ((int)lane == r) L.cand_p = best_p, L.len1 = best, L.cand2 = best_n == 1u ? 1u : 0u, L.job = V3_RESOLVED, L.svc = V3_RUN;
	}
}

// model.c:259-278 in O(1): the composition of the query slice [lq, lq + ll) from the per-word prefix
// composition of the pool (k_comp_prefix): two table entries and the two partial words at the ends.
V3_FN void v3_count_slice(const V3Lane &L, const V3Const &c, u32 *col, u32 sign) {
	const uint4 *cp = reinterpret_cast<const uint4 *>(c.qcomp_base) + (L.q_code - c.qcode_base);
	const u32 b0 = L.lq, b1 = L.lq + L.ll;
	const uint4 p0 = __ldg(cp + (b0 >> 5)), p1 = __ldg(cp + (b1 >> 5));
	const u64 w0 = __ldg(L.q_code + (b0 >> 5)), w1 = __ldg(L.q_code + (b1 >> 5));
	const u64 m0 = ANDI_EVEN_BITS & ((1ULL << (2u * (b0 & 31u))) - 1ULL), m1 = ANDI_EVEN_BITS & ((1ULL << (2u * (b1 & 31u))) - 1ULL);
	const u64 l0 = w0 & m0, h0 = (w0 >> 1) & m0, l1 = w1 & m1, h1 = (w1 >> 1) & m1;
	col[0 * V3_CELL_STRIDE] += (p1.x - p0.x + (u32)__popcll(m1 & ~l1 & ~h1) - (u32)__popcll(m0 & ~l0 & ~h0)) * sign;
	col[5 * V3_CELL_STRIDE] += (p1.y - p0.y + (u32)__popcll(l1 & ~h1) - (u32)__popcll(l0 & ~h0)) * sign;
	col[10 * V3_CELL_STRIDE] += (p1.z - p0.z + (u32)__popcll(h1 & ~l1) - (u32)__popcll(h0 & ~l0)) * sign;
	col[15 * V3_CELL_STRIDE] += (p1.w - p0.w + (u32)__popcll(h1 & l1) - (u32)__popcll(h0 & l0)) * sign;
}

struct V3Acc {	// the count cells of this lane, as walk_step<> wants them
	u32 *col;
	u32 sign;
	__device__ __forceinline__ void add(u32 cell, u32 v) { col[cell * V3_CELL_STRIDE] += v * sign; }
};

// The generic step (walk_step of walk_kernels.cuh) for the few lanes the window jobs do not
// cover; a real call, so its registers do not weigh on the main loop.
template <bool QUARTER>
__device__ __noinline__ void v3_slow_step(const SubjectIndex &S, u32 t, const u64 *q_code, u32 qlen, u32 *col, u32 sign,
										  u32 &pos, u32 &ls, u32 &lq, u32 &ll, u32 &paired) {
	TextView q;
	q.code = q_code, q.spec = nullptr, q.len = qlen, q.mid = 0xffffffffu;
	WalkState w;
	w.pos_q = pos, w.last_s = ls, w.last_q = lq, w.last_len = ll, w.paired = paired;
	V3Acc acc = {col, sign};
	walk_step<QUARTER, false>(S, q, t, w, acc);
	pos = w.pos_q, ls = w.last_s, lq = w.last_q, ll = w.last_len, paired = w.paired;
}

struct V3Env {
	const SubjectIndex &S;
	u64 total;
	u32 *records;
	unsigned long long *counter;
	const QueryView *queries;
	const u32 *query_ids;
	u32 t;
	__device__ __forceinline__ u64 next_unit() { return atomicAdd(counter, 1ULL); }
	template <int PHASE>
	__device__ __forceinline__ bool open_unit(u64 unit, V3Lane &L, const V3Const &c, u32 *col) {
		u32 k, ch;
		v3_split_unit(unit, total, c.cpq, k, ch);
		const u32 qid = query_ids ? query_ids[k] : k;
		if (qid == S.self) return false;
		return v3_begin_unit<PHASE>(L, c, queries[qid].t.code, queries[qid].t.len, ch, records + unit * ANDI_UNIT_WORDS, col);
	}
	template <bool QUARTER>
	__device__ __forceinline__ void slow_step(V3Lane &L, u32 *col, u32 sign) {
		// (copies: a lane field whose address escapes into the call would live in local memory)
		u32 pos = L.pos, ls = L.ls, lq = L.lq, ll = L.ll, paired = L.paired;
		v3_slow_step<QUARTER>(S, t, L.q_code, L.qlen, col, sign, pos, ls, lq, ll, paired);
		L.pos = pos, L.ls = ls, L.lq = lq, L.ll = ll, L.paired = paired;
	}
};

// BURST: the instantiation for pools of near-identical genomes (v3_ext_round, walk_v3_lane.h); the host
// picks it per launch from the mean anchor length of the lane's previous walk (walk_host.cuh). Looking
// for bursts costs the ordinary walk 1.6 % (measured on one box: 565.9 k against 556.8 k pairs/s), so
// it is not in the instantiation that pool runs.
template <int PHASE, bool QUARTER, bool BURST>
__global__ void __launch_bounds__(V3_THREADS, V3_BLOCKS_PER_SM)
k_walk_v3(const SubjectIndex S, const QueryView *__restrict__ queries, const u32 *__restrict__ query_ids, u32 nq, u32 chunk,
		  u32 cpq, u32 threshold, u32 *__restrict__ records, unsigned long long *__restrict__ next_unit) {
	__shared__ u32 cells[16][V3_THREADS];
	__shared__ u32 pend_q[V3_PEND_SLOTS][V3_THREADS], pend_s[V3_PEND_SLOTS][V3_THREADS];
	__shared__ unsigned char pend_g[V3_PEND_SLOTS][V3_THREADS];
	u32 *col = &cells[0][threadIdx.x];
	const V3Pend P = {&pend_q[0][threadIdx.x], &pend_s[0][threadIdx.x], &pend_g[0][threadIdx.x]};
	V3Const c;
	c.t = threshold, c.N = S.rs.len, c.mid = S.rs.mid, c.border = S.rs.len / 2, c.chunk = chunk, c.cpq = cpq, c.K = S.K;
	c.s_code = S.rs.code, c.fdir = S.fdir, c.SA = S.SA, c.qcode_base = S.qcode_base, c.qcomp_base = S.qcomp_base;
	V3Env env = {S, (u64)nq * cpq, records, next_unit, queries, query_ids, threshold};
	V3Lane L;
	L.svc = V3_SVC_FETCH, L.job = V3_STEP;
	L.pos = L.ls = L.lq = L.ll = L.paired = L.cand_p = L.cand2 = L.len1 = L.sumq = L.sumr = L.npend = 0;
	L.q_code = nullptr, L.qlen = 0, L.c_end = 0, L.unit = 0;
	L.b_pos = L.b_ls = L.b_lq = L.b_ll = L.b_paired = 0, L.a_true = 1, L.flag = 1;
	for (u32 trip = 0;; trip++) {
		const unsigned parked = __ballot_sync(0xffffffffu, L.svc != V3_RUN && L.svc != V3_SVC_DONE);
		const unsigned running = __ballot_sync(0xffffffffu, L.svc == V3_RUN);
		if (!(parked | running)) break;
		if (v3_serve_now((u32)__popc(parked), (u32)__popc(running), trip)) {
			if (L.svc != V3_RUN && L.svc != V3_SVC_DONE) v3_service<PHASE, QUARTER>(L, c, env, col, P);
			__syncwarp();
			// (after the lane services: a lean bucket scan that met a repeat has just become a request)
			const unsigned coop = __ballot_sync(0xffffffffu, L.svc == V3_SVC_COOP || L.svc == V3_SVC_COOP2);
			if (coop) v3_coop_scan(L, c, coop);
		}
		if (__any_sync(0xffffffffu, L.npend > V3_PEND_SLOTS - 2u)) {
			// a queue is nearly full: the whole warp classifies what it has queued
			for (u32 k = 0; k < V3_PEND_SLOTS; k++) {
				if (!__any_sync(0xffffffffu, k < L.npend)) break;
				if (k < L.npend) v3_classify_entry(P, k, col);
				__syncwarp();
			}
			L.npend = 0;
		}
		// most of the warp inside long anchors: their windows back to back, without the trip around them
		if (BURST) {
			const u32 run_now = (u32)__popc(running);  // (as of the top of this trip: lanes served since then are not counted)
			if ((trip & (V3_BURST_EVERY - 1u)) == 0u &&
				v3_burst_now((u32)__popc(__ballot_sync(0xffffffffu, L.svc == V3_RUN && L.job == V3_EXT)), run_now)) {
				for (u32 r = 0; r < V3_BURST_ROUNDS; r++) {
					if (L.svc == V3_RUN && L.job == V3_EXT) v3_ext_round(L, c);
					if (!v3_burst_on((u32)__popc(__ballot_sync(0xffffffffu, L.svc == V3_RUN && L.job == V3_EXT)), run_now)) break;
				}
			}
		}
		if (L.svc == V3_RUN) v3_trip<PHASE, QUARTER>(L, c, col, P);
		__syncwarp();
	}
}

// Can this walk go through k_walk_v3? (else: k_walk_chunks_fast)
// (threshold <= K + 15: a tag-1 directory entry must be able to prove a match of threshold length)
static inline bool v3_applies(const SubjectIndex &S, u32 threshold) {
	return S.K > 0 && threshold <= V3_MAX_T && (u32)S.K <= threshold && threshold <= (u32)S.K + 15u && S.fdir != nullptr;
}
