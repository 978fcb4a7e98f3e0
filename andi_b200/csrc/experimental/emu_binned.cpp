// andi_b200/csrc/experimental/emu_binned.cpp -- EXPERIMENTAL, test infrastructure.
//
// Serial host emulation of k_walk_binned (walk_binned.cuh): the per-unit phase logic is the very
// text the kernel compiles (walk_binned_phases.h); only the primitives (loads, bit scans, the
// counters) and the scheduler are host code. One "CTA" of ANDI_BIN_SLOTS slots runs the same
// super-steps -- drain every queue of the current set in order, push into the next set, swap --
// just one unit at a time. tests/test_binned_emulation.py feeds it an index built with numpy
// from the oracle's suffix array and compares the reduced records with the oracle's rows, so the
// state machine of the round-2 kernel is checked on the CPU before it ever meets a GPU.
//
//   g++ -O2 -shared -fPIC -o libemu_binned.so emu_binned.cpp        (make -C andi_b200/csrc emu)
#include <cstdint>
#include <cstring>
#include <vector>

typedef unsigned long long u64;
typedef uint32_t u32;

// ---- the types of walk_kernels.cuh / text.cuh, reduced to the fields the phases touch
struct TextView {
	const u64 *code;
	const u64 *spec;
	u32 len, mid;
};
struct SubjectIndex {
	TextView rs;
	const u32 *SA;
	const u64 *fdir;
	int K;
	u32 self;
	const u64 *qcode_base;
};
struct QueryView {
	TextView t;
	u32 has_sep;
};
#define ANDI_UNIT_WORDS 38
#define ANDI_SCAN_MAX 8
#define ANDI_FDIR_TAG(e) ((u32)((e) >> 62))
#define BIN_FN static inline
#define BIN_OUTLINE_FN static

// ---- primitives
static inline u64 bin_ld64(const u64 *p) { return *p; }
static inline u32 bin_ld32(const u32 *p) { return *p; }
static inline u32 bin_ffs64(u64 x) { return (u32)__builtin_ctzll(x); }
static inline u32 bin_ffs32(u32 x) { return (u32)__builtin_ctz(x); }
static inline u32 bin_popc32(u32 x) { return (u32)__builtin_popcount(x); }
static inline u32 bin_atomic_inc(u32 *p) { return (*p)++; }
static inline u64 bin_next_unit(u64 *p) { return (*p)++; }
static inline u32 min(u32 a, u32 b) { return a < b ? a : b; }
static inline u64 min(u64 a, u64 b) { return a < b ? a : b; }

// text.cuh, restated without funnel shifts
static inline void window64(const u64 *w, u32 pos, u64 &lo, u64 &hi) {
	u32 i = pos >> 5, sh = (pos & 31u) * 2u;
	u64 a = w[i], b = w[i + 1], c = w[i + 2];
	lo = sh ? (a >> sh) | (b << (64 - sh)) : a;
	hi = sh ? (b >> sh) | (c << (64 - sh)) : b;
}
static inline u32 window16(const u64 *w, u32 pos) {
	const u32 *h = reinterpret_cast<const u32 *>(w) + (pos >> 4);
	u64 v = (u64)h[0] | ((u64)h[1] << 32);
	return (u32)(v >> ((pos & 15u) * 2u));
}
static inline u32 kmer_key(u64 win, int k) {
	u32 key = 0;
	for (int c = 0; c < k; c++) key = (key << 2) | (u32)((win >> (2 * c)) & 3u);
	return key;
}
static inline u32 code_at(const u64 *w, u32 pos) { return (u32)(w[pos >> 5] >> ((pos & 31u) * 2u)) & 3u; }

// The generic search, by brute force over every suffix (the emulation's inputs are small).
// Non-SPEC semantics of walk_kernels.cuh: a comparison never crosses '#' (position mid) nor the
// end of RS, and '#' itself matches nothing.
static inline void bin_slow_lookup(const SubjectIndex &S, const u64 *q_code, u32 qlen, u32 pos, u32 &len, bool &unique,
								   u32 &at) {
	const u32 N = S.rs.len, mid = S.rs.mid, rem = qlen - pos;
	u32 best = 0, cnt = 0, where = 0;
	for (u32 p = 0; p < N; p++) {
		u32 run = p < mid ? mid - p : (p == mid ? 0u : N - p), lim = min(rem, run), m = 0;
		while (m < lim && code_at(S.rs.code, p + m) == code_at(q_code, pos + m)) m++;
		if (m > best)
			best = m, cnt = 1, where = p;
		else if (m == best)
			cnt++;
	}
	len = best, unique = best > 0 && cnt == 1, at = where;
}

#include "walk_binned_phases.h"

extern "C" int emu_binned_slots(void) { return ANDI_BIN_SLOTS; }

static u64 emu_rand(u64 &x) {  // xorshift64*
	x ^= x >> 12, x ^= x << 25, x ^= x >> 27;
	return x * 0x2545F4914F6CDD1DULL;
}

// seed == 0: the kernel's lock-step super-steps; seed != 0: a random asynchronous order.
// Returns the number of super-steps (batches for seed != 0), or -1 on bad arguments. stats[q] += units processed in phase q,
// stats[BQ_N + q] += 32-lane batches a CTA would have issued for them (ceil per queue and super-step).
extern "C" long emu_walk_binned(const u64 *s_code, u32 N, u32 mid, const u32 *SA, const u64 *fdir, int K, u32 self,
								u32 threshold, const u64 *pool_code, const u64 *q_word_off, const u32 *q_len, u32 nq,
								u32 chunk, u32 cpq, u32 *records, u64 *stats, u64 seed) {
	if (!s_code || !SA || !fdir || !pool_code || !records || K <= 0 || (u64)nq * cpq >= 0xffffffffULL) return -1;
	SubjectIndex S;
	S.rs.code = s_code, S.rs.spec = nullptr, S.rs.len = N, S.rs.mid = mid;
	S.SA = SA, S.fdir = fdir, S.K = K, S.self = self, S.qcode_base = pool_code;
	std::vector<QueryView> queries(nq);
	for (u32 k = 0; k < nq; k++) {
		queries[k].t.code = pool_code + q_word_off[k], queries[k].t.spec = nullptr;
		queries[k].t.len = q_len[k], queries[k].t.mid = 0xffffffffu, queries[k].has_sep = 0;
	}
	BinConst c;
	c.t = threshold, c.N = N, c.mid = mid, c.border = N / 2, c.chunk = chunk, c.cpq = cpq, c.K = K;
	const u64 total = (u64)nq * cpq;
	u64 next_unit = 0;
	BinShared *shp = new BinShared();
	BinShared &sh = *shp;
	memset(shp, 0, sizeof(BinShared));
	for (u32 t = 0; t < ANDI_BIN_SLOTS; t++) sh.queue[0][BQ_FETCH][t] = (unsigned short)t;
	sh.count[0][BQ_FETCH] = ANDI_BIN_SLOTS;
	u32 cur = 0, waited[BQ_N] = {0};
	long steps = 0;
	if (seed) {
		// Asynchronous order: the phase logic must not depend on the lock-step schedule (a later
		// kernel lets warps pull batches independently). One queue set; repeatedly take a random
		// number of units off the front of a random non-empty queue and run them, pushing straight
		// back into the same set (as ring buffers would).
		u64 rng = seed;
		std::vector<std::vector<unsigned short>> ring(BQ_N);
		for (u32 t = 0; t < ANDI_BIN_SLOTS; t++) ring[BQ_FETCH].push_back((unsigned short)t);
		for (;;) {
			u32 nonempty = 0;
			for (u32 q = 0; q < BQ_N; q++) nonempty += !ring[q].empty();
			if (!nonempty) break;
			u32 q;
			do q = (u32)(emu_rand(rng) % BQ_N);
			while (ring[q].empty());
			u32 take = 1 + (u32)(emu_rand(rng) % 32);
			if (take > ring[q].size()) take = (u32)ring[q].size();
			std::vector<unsigned short> batch(ring[q].begin(), ring[q].begin() + take);
			ring[q].erase(ring[q].begin(), ring[q].begin() + take);
			for (u32 x = 0; x < BQ_N; x++) sh.count[0][x] = 0;
			for (unsigned short s : batch) bin_phase(q, s, sh, 0, S, queries.data(), nullptr, c, total, records, &next_unit);
			for (u32 x = 0; x < BQ_N; x++)
				for (u32 y = 0; y < sh.count[0][x]; y++) ring[x].push_back(sh.queue[0][x][y]);
			if (stats) stats[q] += take, stats[BQ_N + q] += 1;
			steps++;
		}
		delete shp;
		return steps;
	}
	for (;;) {
		u32 pending = 0;
		for (u32 q = 0; q < BQ_N; q++) pending += sh.count[cur][q];
		if (pending == 0) break;
		const u32 nxt = cur ^ 1u;
		for (u32 q = 0; q < BQ_N; q++) {
			const u32 have = sh.count[cur][q];
			if (bin_defer(have, waited[q], pending)) {	// same rule as the kernel: let a thin queue fill up
				for (u32 x = 0; x < have; x++) bin_push(sh, nxt, q, sh.queue[cur][q][x]);
				waited[q]++;
				continue;
			}
			waited[q] = 0;
			if (stats) stats[q] += have, stats[BQ_N + q] += (have + 31) / 32;
			for (u32 x = 0; x < have; x++)
				bin_phase(q, sh.queue[cur][q][x], sh, nxt, S, queries.data(), nullptr, c, total, records, &next_unit);
		}
		for (u32 q = 0; q < BQ_N; q++) sh.count[cur][q] = 0, sh.head[q] = 0;
		cur = nxt;
		steps++;
	}
	delete shp;
	return steps;
}
