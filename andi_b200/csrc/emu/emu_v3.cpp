// andi_b200/csrc/emu/emu_v3.cpp -- TEST INFRASTRUCTURE: serial host emulation of k_walk_v3.
//
// The per-lane logic is the very text the kernels compile (../walk_v3_lane.h); only the
// primitives (loads, bit scans, the unit counter), the generic slow step and the warp loop are
// host code. A "warp" here is 32 lanes run one after the other through the same sections the
// kernel runs (serve parked lanes when v3_serve_now() says so, then one trip of every running
// lane); several warps share the unit counter round-robin. tests/test_walk_v3_emulation.py feeds
// it an index built with numpy from the oracle's suffix array and compares the reduced records
// with the oracle's rows; it also returns the lane statistics DESIGN.md quotes.
//
//   g++ -O2 -std=c++17 -shared -fPIC -o libemu_v3.so emu_v3.cpp
#include <cstdint>
#include <cstring>
#include <vector>

typedef unsigned long long u64;
typedef uint32_t u32;

#define ANDI_UNIT_WORDS 38
#define V3_FN static inline
#define V3_CELL_STRIDE 1
#define V3_SERVE_BATCH 6u
#define V3_SERVE_EVERY 8u
#define V3_COUNT_STATS 1

enum { ST_trips, ST_ext_trips, ST_cand_trips, ST_steps, ST_lucky_hits, ST_lookups, ST_tag0, ST_tag1, ST_wide_gaps, ST_slow_steps, ST_slow_tail, ST_slow_tag3, ST_wide_pairs, ST_slow_long, ST_cols_trips, ST_pushes, ST_drained, ST_drains, ST_drain_rounds, ST_coop_scans, ST_ext_rounds, ST_bursts,
	   ST_warp_trips, ST_running_lanes, ST_services, ST_served_lanes, ST_N };
static u64 *g_stats = nullptr;
#define V3_STAT(name)                \
	do {                             \
		if (g_stats) g_stats[ST_##name]++; \
	} while (0)

static inline u32 v3_ctz64(u64 x) { return (u32)__builtin_ctzll(x); }
static inline u32 v3_ctz32(u32 x) { return x ? (u32)__builtin_ctz(x) : 32u; }
static inline u32 v3_popc32(u32 x) { return (u32)__builtin_popcount(x); }
static inline u64 v3_ld_fdir(const u64 *p) { return *p; }
static inline u32 v3_ld_sa(const u32 *p) { return *p; }
static inline void v3_window64(const u64 *w, u32 pos, u64 &lo, u64 &hi) {
	u32 i = pos >> 5, sh = (pos & 31u) * 2u;
	u64 a = w[i], b = w[i + 1], c = w[i + 2];
	lo = sh ? (a >> sh) | (b << (64 - sh)) : a;
	hi = sh ? (b >> sh) | (c << (64 - sh)) : b;
}
static inline u32 v3_kmer_key(u64 win, int k) {
	u32 key = 0;
	for (int c = 0; c < k; c++) key = (key << 2) | (u32)((win >> (2 * c)) & 3u);
	return key;
}

struct V3Lane;
struct V3Const;
static inline void v3_count_slice(const V3Lane &L, const V3Const &c, u32 *col, u32 sign);

#include "../walk_v3_lane.h"

// model.c:259-278 by counting the characters of the query slice (the kernel uses the pool's
// prefix-composition table; the emulation has no such table and does not need the speed)
static inline void v3_count_slice(const V3Lane &L, const V3Const &, u32 *col, u32 sign) {
	for (u32 x = 0; x < L.ll; x++) {
		const u32 q = (u32)(L.q_code[(L.lq + x) >> 5] >> (((L.lq + x) & 31u) * 2u)) & 3u;
		col[q * 5u] += sign;
	}
}

static inline u32 code_at(const u64 *w, u32 pos) { return (u32)(w[pos >> 5] >> ((pos & 31u) * 2u)) & 3u; }

struct Host {
	const u64 *s_code;
	const u32 *SA;
	u32 N, mid;
	u32 run(u32 p) const { return p < mid ? mid - p : (p == mid ? 0u : N - p); }
	u32 sym_s(u32 p) const { return p >= N ? 0u : (p == mid ? 2u : code_at(s_code, p) + 4u); }
	u32 match(const u64 *q, u32 qpos, u32 rem, u32 p) const {
		u32 lim = rem < run(p) ? rem : run(p), m = 0;
		while (m < lim && code_at(s_code, p + m) == code_at(q, qpos + m)) m++;
		return m;
	}
	// longest match of q[qpos..qpos+rem) anywhere in RS (non-SPEC semantics of walk_kernels.cuh)
	void lookup(const u64 *q, u32 qpos, u32 rem, u32 &len, bool &unique, u32 &at) const {
		u32 lo = 0, hi = N;
		while (lo < hi) {
			u32 m = lo + ((hi - lo) >> 1), p = SA[m], c = match(q, qpos, rem, p);
			bool less = c == rem ? false : sym_s(p + c) < code_at(q, qpos + c) + 4u;
			if (less)
				lo = m + 1;
			else
				hi = m;
		}
		int lm = lo > 0 ? (int)match(q, qpos, rem, SA[lo - 1]) : -1, rm = lo < N ? (int)match(q, qpos, rem, SA[lo]) : -1;
		if (lm <= 0 && rm <= 0) {
			len = 0, unique = false, at = 0;
			return;
		}
		if (rm >= lm) {
			len = (u32)rm, at = SA[lo];
			unique = rm > lm && !(lo + 1 < N && (int)match(q, qpos, rem, SA[lo + 1]) >= rm);
		} else {
			len = (u32)lm, at = SA[lo - 1];
			unique = !(lo >= 2 && (int)match(q, qpos, rem, SA[lo - 2]) >= lm);
		}
	}
};

struct Env {
	Host h;
	u64 total, counter;
	u32 *records;
	const u64 *pool_code;
	const u64 *q_off;
	const u32 *q_len;
	u32 self, nq;
	u64 next_unit() { return counter++; }
	template <int PHASE>
	bool open_unit(u64 unit, V3Lane &L, const V3Const &c, u32 *col) {
		u32 k, ch;
		v3_split_unit(unit, total, c.cpq, k, ch);
		if (k == self) return false;
		return v3_begin_unit<PHASE>(L, c, pool_code + q_off[k], q_len[k], ch, records + unit * ANDI_UNIT_WORDS, col);
	}
	// one iteration of src/process.c:153-197 in the generic form (walk_step of walk_kernels.cuh)
	template <bool QUARTER>
	void slow_step(V3Lane &L, u32 *col, u32 sign) {
		V3_STAT(slow_steps);
		V3_STAT(steps);
		const u64 *q = L.q_code;
		const u32 t = c_t, border = h.N / 2, rem = L.qlen - L.pos;
		u32 cur_s = 0, cur_len = 0;
		bool found = false;
		u32 advance = L.pos - L.lq, gap = advance - L.ll, guess = L.ls + advance;
		if (guess < h.N && gap <= t) {
			cur_s = guess, cur_len = h.match(q, L.pos, rem, guess);
			found = cur_len >= t;
		}
		if (!found) {
			bool unique;
			u32 at;
			h.lookup(q, L.pos, rem, cur_len, unique, at);
			found = unique && cur_len >= t;
			if (found) cur_s = at;
		}
		if (found) {
			u32 end_s = L.ls + L.ll, end_q = L.lq + L.ll;
			bool pairs = cur_s > end_s && (L.pos - end_q) == (cur_s - end_s) && ((cur_s < border) == (L.ls < border));
			if (pairs || L.paired || L.ll >= 2 * t) {
				if (QUARTER) {
					L.sumq += (L.ll >> 2) * sign, L.sumr += (L.ll & 3u) * sign;
				} else {
					V3Const none{};
					v3_count_slice(L, none, col, sign);
				}
			}
			if (pairs)
				for (u32 x = 0; x < L.pos - end_q; x++) {
					if (end_s + x == h.mid) continue;
					col[code_at(h.s_code, end_s + x) * 4 + code_at(q, end_q + x)] += sign;
				}
			L.paired = pairs ? 1u : 0u;
			L.ls = cur_s, L.lq = L.pos, L.ll = cur_len;
		}
		L.pos += cur_len + 1;
	}
	u32 c_t;
};

template <int PHASE, bool QUARTER>
static void run_phase(Env &env, const V3Const &c, u32 n_warps) {
	struct Warp {
		V3Lane lane[32];
		u32 cells[32][16];
		u32 pq[32][V3_PEND_SLOTS], ps[32][V3_PEND_SLOTS];
		unsigned char pg[32][V3_PEND_SLOTS];
		V3Pend pend(u32 x) { return V3Pend{pq[x], ps[x], pg[x]}; }
		u32 trip = 0;
		bool done = false;
	};
	std::vector<Warp> warps(n_warps);
	for (auto &w : warps) {
		memset(w.lane, 0, sizeof w.lane);
		for (auto &l : w.lane) l.svc = V3_SVC_FETCH;
	}
	env.counter = 0;
	for (u32 live = n_warps; live;) {
		for (auto &w : warps) {
			if (w.done) continue;
			u32 parked = 0, running = 0, idle = 0;
			for (auto &l : w.lane) parked += l.svc != V3_RUN && l.svc != V3_SVC_DONE, running += l.svc == V3_RUN, idle += l.svc == V3_SVC_DONE;
			if (idle == 32) {
				w.done = true, live--;
				continue;
			}
			if (v3_serve_now(parked, running, w.trip)) {
				if (g_stats) g_stats[ST_services]++, g_stats[ST_served_lanes] += parked;
				for (u32 x = 0; x < 32; x++)
					if (w.lane[x].svc != V3_RUN && w.lane[x].svc != V3_SVC_DONE) v3_service<PHASE, QUARTER>(w.lane[x], c, env, w.cells[x], w.pend(x));
			}
			u32 most = 0;
			for (auto &l : w.lane) most = l.npend > most ? l.npend : most;
			if (most > V3_PEND_SLOTS - 2u) {  // a queue is nearly full: the whole warp classifies what it has queued
				if (g_stats) g_stats[ST_drains]++, g_stats[ST_drain_rounds] += most;
				for (u32 x = 0; x < 32; x++) v3_drain_lane(w.lane[x], w.pend(x), w.cells[x]);
			}
			auto ext_lanes = [&w]() {
				u32 n = 0;
				for (auto &l : w.lane) n += l.svc == V3_RUN && l.job == V3_EXT;
				return n;
			};
			u32 run_now = 0;
			for (auto &l : w.lane) run_now += l.svc == V3_RUN;
			if ((w.trip & (V3_BURST_EVERY - 1u)) == 0u && v3_burst_now(ext_lanes(), run_now)) {  // most of the running lanes inside long anchors (walk_v3.cuh)
				if (g_stats) g_stats[ST_bursts]++;
				for (u32 r = 0; r < V3_BURST_ROUNDS; r++) {
					for (auto &l : w.lane)
						if (l.svc == V3_RUN && l.job == V3_EXT) v3_ext_round(l, c);
					if (!v3_burst_on(ext_lanes(), run_now)) break;
				}
			}
			running = 0;
			for (u32 x = 0; x < 32; x++)
				if (w.lane[x].svc == V3_RUN) running++, v3_trip<PHASE, QUARTER>(w.lane[x], c, w.cells[x], w.pend(x));
			if (g_stats && running) g_stats[ST_warp_trips]++, g_stats[ST_running_lanes] += running;
			w.trip++;
		}
	}
}

extern "C" int emu_v3_stats(void) { return ST_N; }

// Both launches (PHASE 1 then PHASE 2) for one subject. Returns 0, or -1 on bad arguments.
extern "C" long emu_walk_v3(const u64 *s_code, u32 N, u32 mid, const u32 *SA, const u64 *fdir, int K, u32 self, u32 threshold,
							const u64 *pool_code, const u64 *q_word_off, const u32 *q_len, u32 nq, u32 chunk, u32 cpq,
							u32 *records, u64 *stats, u32 n_warps, int quarter) {
	if (!s_code || !SA || !fdir || !pool_code || !records || threshold > V3_MAX_T || K > (int)threshold || (int)threshold > K + 15 || n_warps == 0) return -1;
	V3Const c;
	c.t = threshold, c.N = N, c.mid = mid, c.border = N / 2, c.chunk = chunk, c.cpq = cpq, c.K = K, c.s_code = s_code, c.fdir = fdir, c.SA = SA;
	Env env;
	env.h.s_code = s_code, env.h.SA = SA, env.h.N = N, env.h.mid = mid;
	env.total = (u64)nq * cpq, env.records = records, env.pool_code = pool_code, env.q_off = q_word_off, env.q_len = q_len;
	env.self = self, env.nq = nq, env.c_t = threshold;
	g_stats = stats;
	c.qcode_base = nullptr, c.qcomp_base = nullptr;
	if (quarter)
		run_phase<1, true>(env, c, n_warps);
	else
		run_phase<1, false>(env, c, n_warps);
	g_stats = stats ? stats + ST_N : nullptr;
	if (quarter)
		run_phase<2, true>(env, c, n_warps);
	else
		run_phase<2, false>(env, c, n_warps);
	g_stats = nullptr;
	return 0;
}
