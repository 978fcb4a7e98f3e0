// andi_b200/csrc/walk_v3_lane.h -- per-lane logic of the round-2 anchor-walk kernels
// (k_walk_v3, walk_v3.cuh). Compiles twice from this one text: into the CUDA kernels (nvcc) and
// into the serial warp emulation of emu/emu_v3.cpp (g++; tests/test_walk_v3_emulation.py), so the
// state machine is checked against the oracle on the CPU before it meets a GPU.
//
// Same units, same records, same results as k_walk_chunks (walk_kernels.cuh): one lane walks one
// (query, chunk) unit through the loop of src/process.c:153-197. What changes against the phase
// pipeline of round 1 (13.6 of 32 lanes active, ~440 thread instructions per walk step):
//
//  * ONE trip of the main loop = ONE window job for every running lane, the same straight-line
//    code for all of them: load 64 columns of query and subject, XOR, find the first mismatch,
//    act. A job is STEP (start of a walk step: the lucky diagonal of src/process.c:82-100), EXT
//    (an anchor longer than the window keeps growing) or CAND (the only suffix that carries the
//    query's k-mer, from the directory view fdir). 93 % of all walk steps end in their first trip.
//  * The lucky window starts at the END OF THE PREVIOUS ANCHOR, not at pos_Q: the gap columns that
//    src/model.c:309-337 classifies when the new anchor pairs with the last one are the low
//    columns of the very window the compare has just loaded. No COLS phase, no second memory
//    round, no remembered mismatch class.
//  * Anchor interiors (src/model.c:247-254: len/4 to each diagonal cell, remainder to TtoT) go
//    to two registers (sum of len>>2, sum of len&3), not to four shared-memory cells.
//  * The boundary replay (the true chain entering chunk c+1 against that chunk's amnesic walk) is
//    a SEPARATE LAUNCH (PHASE 2) over the same code: no lane of the main launch carries the
//    second chain or tests for it.
//  * Everything rare -- fetching a unit, writing its record, buckets with several suffixes,
//    query tails shorter than K, an ESA anchor that pairs (gap columns not in registers) -- is
//    a SERVICE request: the lane parks, and the warp serves parked lanes together once enough
//    of them wait (or every few trips), so the rare code runs with several lanes instead of one.
//    The slow step itself is walk_step<>() of walk_kernels.cuh, the proven generic form.
//
// Texts without separators only (RAW / JC / KIMURA, and LOGDET / ANI for pool queries: their
// anchor interiors need the pool's prefix-composition table); join mode stays with
// k_walk_chunks_fast.
#pragma once

#ifndef V3_FN
#error "define V3_FN and the v3_* primitives before including walk_v3_lane.h"
#endif

enum : u32 { V3_RUN = 0, V3_SVC_FETCH = 1, V3_SVC_FINISH = 2, V3_SVC_SLOW = 3, V3_SVC_DONE = 4, V3_SVC_COOP = 5, V3_SVC_COOP2 = 6 };
enum : u32 { V3_STEP = 0, V3_EXT = 1, V3_COLS = 2, V3_CAND1 = 3, V3_CAND2 = 4, V3_RESOLVED = 5 };
#define V3_COOP_MAX 4096u  // buckets up to this size are scanned by the whole warp (v3_scan_full), larger ones by the generic step

#define V3_EVEN 0x5555555555555555ULL
#define V3_MAX_T 31u	   // the gap columns of a lucky anchor (<= threshold of them) lie in the low window word
#ifndef V3_PEND_SLOTS
#define V3_PEND_SLOTS 6u   // pending-gap queue entries per lane (16 columns each)
#endif

struct V3Const {
	u32 t, N, mid, border, chunk, cpq;
	int K;
	const u64 *s_code;
	const u64 *fdir;
	const u32 *SA;
	// LOGDET / ANI only: base of the pool's code plane and of its prefix-composition table (same
	// word geometry, k_comp_prefix): comp = qcomp_base + (q_code - qcode_base)
	const u64 *qcode_base;
	const void *qcomp_base;
};

// Chain A is the one being advanced. PHASE 2 carries a second chain B and `a_true` (A is the
// true chain / the amnesic one); `sign` (+1 / -1 as u32) follows a_true.
struct V3Lane {
	u32 svc, job;
	u32 pos, ls, lq, ll, paired;
	u32 cand_p, cand2, len1;
	u32 sumq, sumr;
	u32 npend;
	const u64 *q_code;
	u32 qlen, c_end;
	u64 unit;
	u32 b_pos, b_ls, b_lq, b_ll, b_paired, a_true, flag;
};

// This lane's column of the pending-gap queue (entry k at q[k * V3_CELL_STRIDE] etc.): gap columns
// that src/model.c:309-337 has to classify, 16 per entry as 32-bit halves of the two windows.
// Classifying them on the spot would run with the three or four lanes of a warp that happen to have
// a wide gap in this trip (28 % of all warp instructions in the first form of this kernel); queued,
// the whole warp classifies together once any lane's queue is nearly full (v3_drain).
struct V3Pend {
	u32 *q, *s;
	unsigned char *g;  // number of columns (1..16) | 0x80 when the entry counts negatively (PHASE 2)
};

V3_FN void v3_push_gap(V3Lane &L, const V3Pend &P, u64 q0, u64 s0, u32 g, u32 sign) {
	const u32 neg = sign == 1u ? 0u : 0x80u;
	const u32 k = L.npend;
	P.q[k * V3_CELL_STRIDE] = (u32)q0, P.s[k * V3_CELL_STRIDE] = (u32)s0;
	P.g[k * V3_CELL_STRIDE] = (unsigned char)((g < 16u ? g : 16u) | neg);
	if (g > 16u) {
		P.q[(k + 1u) * V3_CELL_STRIDE] = (u32)(q0 >> 32), P.s[(k + 1u) * V3_CELL_STRIDE] = (u32)(s0 >> 32);
		P.g[(k + 1u) * V3_CELL_STRIDE] = (unsigned char)((g - 16u) | neg);
	}
	L.npend = k + (g > 16u ? 2u : 1u);
	V3_STAT(pushes);
}

// model.c:309-337 for queue entry k of this lane (the caller knows k < L.npend).
V3_FN void v3_classify_entry(const V3Pend &P, u32 k, u32 *col) {
	const u32 q = P.q[k * V3_CELL_STRIDE], s = P.s[k * V3_CELL_STRIDE], gb = P.g[k * V3_CELL_STRIDE];
	const u32 g = gb & 0x7fu, sign = (gb & 0x80u) ? 0xffffffffu : 1u;
	const u32 vm = g >= 16u ? 0x55555555u : (0x55555555u & ((1u << (2u * g)) - 1u));
	const u32 x = q ^ s;
	u32 neq = (x | (x >> 1)) & vm;
	const u32 eq = vm & ~neq, ql = q & eq, qh = (q >> 1) & eq;
	const u32 nt = v3_popc32(ql & qh), nc = v3_popc32(ql) - nt, ng = v3_popc32(qh) - nt, na = v3_popc32(eq) - nc - ng - nt;
	col[0 * V3_CELL_STRIDE] += na * sign;
	col[5 * V3_CELL_STRIDE] += nc * sign;
	col[10 * V3_CELL_STRIDE] += ng * sign;
	col[15 * V3_CELL_STRIDE] += nt * sign;
	while (neq) {
		const u32 b = v3_ctz32(neq);
		neq &= neq - 1u;
		col[((((s >> b) & 3u) << 2) | ((q >> b) & 3u)) * V3_CELL_STRIDE] += sign;
	}
	V3_STAT(drained);
}

// First differing column of a 64-column XOR window (64 when there is none), without branches.
V3_FN u32 v3_first_diff(u64 x0, u64 x1) {
	const u32 a = (u32)x0, b = (u32)(x0 >> 32), c = (u32)x1, d = (u32)(x1 >> 32);
	// v3_ctz32(0) == 32: a piece without a difference contributes its 16 columns
	const u32 da = v3_ctz32(a) >> 1, db = 16u + (v3_ctz32(b) >> 1), dc = 32u + (v3_ctz32(c) >> 1), dd = 48u + (v3_ctz32(d) >> 1);
	return a ? da : (b ? db : (c ? dc : dd));
}

// One trip of a running lane (L.svc == V3_RUN). `col` = this lane's column of the count cells
// (cell x at col[x * V3_CELL_STRIDE]). On entry L.npend <= V3_PEND_SLOTS - 2.
template <int PHASE, bool QUARTER>
V3_FN void v3_trip(V3Lane &L, const V3Const &c, u32 *col, const V3Pend &P) {
	const u32 t = c.t;
	u32 sign = 1u;
	if (L.job == V3_STEP) {
		if (PHASE == 1) {
			if (L.pos >= L.c_end) {
				L.svc = V3_SVC_FINISH;
				return;
			}
		} else {
			// the two chains from the start of chunk c+1 on (L.c_end = the end of the QUERY): stop when
			// they are in the same state, else advance the one that is behind -- for as long as it takes,
			// beyond chunk c+1 if need be (a boundary inside a repeat, an anchor-free stretch). The
			// correction stays exact wherever they meet: with X_j = the amnesic chain of chunk j carried on
			// to the end of the query and W_j its counts from the start of chunk j, W_j = U_j + D_j +
			// W_(j+1) holds for D_j = (X_j minus X_(j+1) up to their meeting point), whichever chunk that
			// point lies in. Chains that reach the end of the query apart (flag 0) have D_j just the same.
			bool same = L.pos == L.b_pos && L.ls == L.b_ls && L.lq == L.b_lq && L.ll == L.b_ll && L.paired == L.b_paired;
			u32 t_pos = L.a_true ? L.pos : L.b_pos, p_pos = L.a_true ? L.b_pos : L.pos;
			if (same || (t_pos >= L.c_end && p_pos >= L.c_end)) {
				L.flag = same ? 1u : 0u;
				L.svc = V3_SVC_FINISH;
				return;
			}
			bool step_true = t_pos <= p_pos;
			if (step_true != (L.a_true != 0u)) {
				u32 x;
				x = L.pos, L.pos = L.b_pos, L.b_pos = x;
				x = L.ls, L.ls = L.b_ls, L.b_ls = x;
				x = L.lq, L.lq = L.b_lq, L.b_lq = x;
				x = L.ll, L.ll = L.b_ll, L.b_ll = x;
				x = L.paired, L.paired = L.b_paired, L.b_paired = x;
				L.a_true ^= 1u;
			}
		}
	}
	if (PHASE == 2) sign = L.a_true ? 1u : 0xffffffffu;

	const u32 job = L.job;
	const u32 end_q = L.lq + L.ll, end_s = L.ls + L.ll;
	const bool is_step = job == V3_STEP, is_ext = job == V3_EXT, is_cols = job == V3_COLS, is_cand = job >= V3_CAND1;
	const u32 g = L.pos - end_q;	 // process.c:88-89, gap = advance - last.length (pos_Q is fixed while a step lasts)
	const u32 guess = end_s + g;	 // process.c:91: last.pos_S + advance
	// The window starts at the END OF THE LAST ANCHOR whenever the gap to pos_Q fits (g <= V3_MAX_T),
	// lucky attempt (process.c:93: g <= t) or not: the diagonal compare is then at hand for a
	// directory candidate that turns out to lie on this diagonal, together with its gap columns.
	const bool diag_ok = is_step && g <= V3_MAX_T && guess < c.N;
	const bool lucky = diag_ok && g <= t;
	// CAND1 / CAND2 (the two suffixes of a tag-2 directory entry): the one on the diagonal of the
	// last anchor would pair with it (process.c:167-169), so it is compared through the same window
	const bool diag = is_cand && g <= V3_MAX_T && L.cand_p - end_s == g;
	const bool from_end = is_ext || diag_ok || diag;
	// COLS: the gap columns of an anchor that paired over more than V3_MAX_T columns, 32 per trip;
	// they end where the anchor (already the "last" one) begins, L.len1 of them are left
	const u32 wq = from_end ? end_q : (is_cols ? L.lq - L.len1 : L.pos);
	const u32 ws = from_end ? end_s : (is_cols ? L.ls - L.len1 : (is_cand ? L.cand_p : 0u));
	const u32 gg = (diag_ok || diag) ? g : 0u;	// columns of the window in front of the compare
	const u32 cq = wq + gg, cs = ws + gg;
	const u32 run = cs < c.mid ? c.mid - cs : (cs == c.mid ? 0u : c.N - cs);
	const u32 rem = L.qlen - cq;
	const u32 clim = (is_step && !diag_ok) ? 0u : (rem < run ? rem : run);

	u64 q0, q1, s0, s1;
	v3_window64(L.q_code, wq, q0, q1);
	v3_window64(c.s_code, ws, s0, s1);
	const u64 x0 = ((q0 ^ s0) >> (2u * gg)) << (2u * gg);  // gg <= V3_MAX_T
	const u32 D = v3_first_diff(x0, q1 ^ s1);
	const u32 raw = D - gg;
	const bool complete = D < 64u || raw >= clim;
	u32 matched = raw < clim ? raw : clim;
	V3_STAT(trips);

	// ---- what the window decides. Deliberately written as selects, not as if-blocks: every branch
	// region runs with the few lanes that need it while the rest of the warp waits (the first form
	// of this kernel spent two thirds of its issue slots that way).
	const bool c1 = job == V3_CAND1, c2 = job == V3_CAND2, c3 = job == V3_RESOLVED;
	const u32 l1 = L.len1;
	// EXT: the anchor grows; done when the window saw its end
	L.ll += is_ext ? matched : 0u;
	const bool ext_done = is_ext && complete;
	// CAND1 (first of two suffixes that carry the k-mer): remember its length, compare the other one.
	// CAND2: the longer of the two is the match; equal lengths = not unique (process.c:122).
	// A candidate longer than the window is fine when it is the only / the longer one (EXT follows).
	// RESOLVED: the bucket scan is done (v3_scan_full: three or more suffixes, or two long ones): best
	// candidate in cand_p, its length in len1, cand2 = unique; this trip only accounts for it.
	const bool slow_long = (c1 && !complete) || (c2 && !complete && l1 >= matched);
	const bool tie = (c2 && l1 == matched) || (c3 && L.cand2 == 0u), first_better = c2 && l1 > matched;
	u32 cur_s = is_cand ? (first_better ? L.cand2 : L.cand_p) : guess;
	bool in_window = (diag_ok || diag) && !first_better;
	bool complete_a = complete || c3;  // the anchor (if any) ends inside what has been compared
	matched = (first_better || c3) ? l1 : matched;
	const bool cand_final = (c2 && !slow_long) || c3;
	bool anchor = is_step ? (lucky && matched >= t) : (cand_final && !tie && matched >= t);
	bool plain_end = cand_final && !anchor;	 // process.c:122: no anchor, pos_Q += length + 1
	const bool lookup = is_step && !anchor;
	{
		const u32 p = L.cand_p;
		L.cand_p = c1 ? L.cand2 : p, L.cand2 = c1 ? p : L.cand2, L.len1 = c1 ? matched : l1;
	}
	u32 next_job = c1 ? (u32)V3_CAND2 : (ext_done ? (u32)V3_STEP : job);
	if (ext_done) L.pos = L.lq + L.ll + 1u;
	// two suffixes that both run past the window (a repeat with two copies): the whole warp compares
	// them to their ends (v3_scan_full); both positions are in cand_p / cand2 after the swap above
	if (slow_long) next_job = V3_STEP, L.svc = V3_SVC_COOP2;
	L.job = next_job;
	// what goes to the pending-gap queue at the end of this trip: COLS trips fetch the gap columns of
	// an anchor that paired over more than V3_MAX_T columns, 32 per trip
	u32 push_n = is_cols ? (l1 < 32u ? l1 : 32u) : 0u;
	if (is_cols) {
		L.len1 = l1 - push_n;
		if (l1 == push_n) L.job = L.cand2 ? V3_EXT : V3_STEP;
	}
#ifdef V3_COUNT_STATS
	if (is_ext) V3_STAT(ext_trips);
	if (is_cols) V3_STAT(cols_trips);
	if (is_cand) V3_STAT(cand_trips);
	if (slow_long) V3_STAT(slow_long);
#endif

	if (lookup) {
		// process.c:117: the longest match anywhere in RS, through the directory view
		if (c.K <= 0 || L.qlen - L.pos < (u32)c.K) {
			V3_STAT(slow_tail);
			L.svc = V3_SVC_SLOW;
		} else {
			const u64 kw = gg ? ((q0 >> (2u * gg)) | (q1 << (64u - 2u * gg))) : q0;
			const u64 fe = v3_ld_fdir(c.fdir + v3_kmer_key(kw, c.K));
			const u32 tag = (u32)(fe >> 62);
			V3_STAT(lookups);
			// tag 0, absent k-mer: only the length matters (it is < K <= threshold).
			// tag 1, ONE suffix carries the k-mer: the entry holds its text position and the 15 bases
			// that follow the k-mer there, so the match length up to K + 15 (>= threshold) comes out
			// of the entry -- no candidate window, no second trip. A match that long is an anchor at
			// once (a single candidate is always unique, process.c:122) and grows in EXT trips.
			// (computed for every lane that looks up, selected by tag: no branch region of its own)
			const bool t0 = tag == 0u, t1 = tag == 1u;
			const u32 p1 = (u32)fe & 0x7fffffffu;
			const u32 x = ((u32)(kw >> (2 * c.K)) ^ (u32)(fe >> 31)) & 0x3fffffffu;
			const u32 d = v3_ctz32(x) >> 1;	 // 16 when all 15 agree
			const u32 prun = p1 < c.mid ? c.mid - p1 : (p1 == c.mid ? 0u : c.N - p1), prem = L.qlen - L.pos;
			const u32 plim = prem < prun ? prem : prun, have = (u32)c.K + (d < 15u ? d : 15u);
			const u32 len1 = have < plim ? have : plim;
			matched = t0 ? (u32)fe : (t1 ? len1 : matched);
			complete_a = t1 ? (d < 15u || have >= plim) : complete_a;
			anchor = t1 && len1 >= t;
			plain_end = t0 || (t1 && !anchor);
			cur_s = t1 ? p1 : cur_s;
			in_window = t1 ? (diag_ok && p1 == guess) : in_window;
#ifdef V3_COUNT_STATS
			if (t0) V3_STAT(tag0);
			if (t1) V3_STAT(tag1);
#endif
			// tag 2: two suffixes, both text positions in the entry -- the one that could pair goes
			// last so that its window is the one at hand when the anchor is accounted
			const u32 p2 = (u32)(fe >> 31) & 0x7fffffffu;
			const bool p1_diag = p1 - end_s == g;
			if (tag == 2u) L.cand_p = p1_diag ? p2 : p1, L.cand2 = p1_diag ? p1 : p2, L.job = V3_CAND1;
			if (tag == 3u) {  // three or more: the service routine scans the bucket
				V3_STAT(slow_tag3);
				L.cand_p = (u32)fe, L.cand2 = (u32)(fe >> 32) & 0x3fffffffu;
				L.svc = L.cand2 <= V3_COOP_MAX ? V3_SVC_COOP : V3_SVC_SLOW;
			}
		}
	}
	if (plain_end) {
		V3_STAT(steps);
		L.pos += matched + 1u, L.job = V3_STEP;
	}
	if (anchor) {
		// ---- process.c:160-196
		V3_STAT(steps);
		if (is_step) V3_STAT(lucky_hits);
		const bool pairs = cur_s > end_s && (L.pos - end_q) == (cur_s - end_s) && ((cur_s < c.border) == (L.ls < c.border));
		if (pairs || L.paired || L.ll >= 2u * t) {	// the previous anchor's interior
			if (QUARTER) {	// model.c:247-254 (RAW / JC / KIMURA): a quarter to each diagonal cell, remainder to TtoT
				L.sumq += (L.ll >> 2) * sign;
				L.sumr += (L.ll & 3u) * sign;
			} else {  // model.c:259-278 (LOGDET / ANI): the composition of the query slice
				v3_count_slice(L, c, col, sign);
			}
		}
		L.ls = cur_s, L.lq = L.pos, L.ll = matched, L.paired = pairs ? 1u : 0u;
		if (complete_a) L.pos += matched + 1u;
		L.job = complete_a ? V3_STEP : V3_EXT;
		if (pairs) {
			// model.c:309-337 on the g gap columns (g >= 1)
			if (!in_window) {
				// more than V3_MAX_T of them, or a window that does not hold them: COLS trips fetch
				// them before the walk goes on; L.cand2 remembers whether the anchor still grows
				V3_STAT(wide_pairs);
				L.len1 = g, L.cand2 = complete_a ? 0u : 1u, L.job = V3_COLS;
			} else if (g == 1u) {
				col[((((u32)s0 & 3u) << 2) | ((u32)q0 & 3u)) * V3_CELL_STRIDE] += sign;
			} else {
				V3_STAT(wide_gaps);
				push_n = g;
			}
		}
	}
	if (push_n) v3_push_gap(L, P, q0, s0, push_n, sign);
}

// One more window of a growing anchor (job EXT) outside the trips: what an EXT trip does, and nothing
// else. Near-identical genomes (outbreak isolates: a few SNPs per megabase) have anchors of tens of
// kilobases, every unit runs its first one to its end, and a whole trip per 64 columns made such a pool
// several times slower than a pool of ordinary divergence -- the reference's easiest case. When most of
// a warp is in EXT (v3_burst_now) the warp takes these rounds back to back instead.
V3_FN void v3_ext_round(V3Lane &L, const V3Const &c) {
	const u32 eq = L.lq + L.ll, es = L.ls + L.ll;
	const u32 erun = es < c.mid ? c.mid - es : (es == c.mid ? 0u : c.N - es), erem = L.qlen - eq;
	const u32 elim = erem < erun ? erem : erun;
	u64 a0, a1, b0, b1;
	v3_window64(L.q_code, eq, a0, a1);
	v3_window64(c.s_code, es, b0, b1);
	const u32 E = v3_first_diff(a0 ^ b0, a1 ^ b1);
	L.ll += E < elim ? E : elim;
	V3_STAT(ext_rounds);
	if (E < 64u || E >= elim) L.pos = L.lq + L.ll + 1u, L.job = V3_STEP;
}
#define V3_BURST_EVERY 8u	 // a warp looks for a burst every so many trips (a power of two): the look itself costs the ordinary walk
#define V3_BURST_ROUNDS 64u	 // windows per lane and burst at most
// a burst starts when three in five of the running lanes are in EXT and goes on while three in ten (of those that ran then) are
V3_FN bool v3_burst_now(u32 ext_lanes, u32 running) { return ext_lanes != 0u && 5u * ext_lanes >= 3u * running; }
V3_FN bool v3_burst_on(u32 ext_lanes, u32 running) { return ext_lanes != 0u && 10u * ext_lanes >= 3u * running; }

// Classify everything this lane has queued.
V3_FN void v3_drain_lane(V3Lane &L, const V3Pend &P, u32 *col) {
	for (u32 k = 0; k < L.npend; k++) v3_classify_entry(P, k, col);
	L.npend = 0;
}

// Unit number -> (query index k, chunk c). PHASE 2 units are the boundaries: unit (k, c) replays
// the entry into chunk c + 1.
V3_FN void v3_split_unit(u64 unit, u64 total, u32 cpq, u32 &k, u32 &c) {
	if (total <= 0xffffffffULL) {
		k = (u32)unit / cpq, c = (u32)unit - k * cpq;
	} else {
		k = (u32)(unit / cpq), c = (u32)(unit % cpq);
	}
}

// Start unit (qid's planes at q_code, length qlen, chunk c). Returns false when the unit holds no work.
// PHASE 1: records of a last chunk also get D = 0, flag = 1 here (nothing follows it).
template <int PHASE>
V3_FN bool v3_begin_unit(V3Lane &L, const V3Const &c, const u64 *q_code, u32 qlen, u32 chunk_no, u32 *rec, u32 *col) {
	const u64 start = (u64)chunk_no * c.chunk;
	if (start >= qlen) return false;
	const u64 end1 = start + c.chunk;
	const u32 c_end = (u32)(end1 < qlen ? end1 : qlen);
	if (PHASE == 2 && c_end >= qlen) return false;	// last chunk: no boundary
	L.q_code = q_code, L.qlen = qlen;
	L.job = V3_STEP, L.cand_p = 0, L.cand2 = 0, L.len1 = 0, L.sumq = 0, L.sumr = 0, L.flag = 1, L.npend = 0;
#pragma unroll
	for (int x = 0; x < 16; x++) col[x * V3_CELL_STRIDE] = 0;
	if (PHASE == 1) {
		L.pos = (u32)start, L.ls = L.lq = L.ll = L.paired = 0;
		L.c_end = c_end;
		if (c_end >= qlen) {
#pragma unroll
			for (int x = 0; x < 16; x++) rec[16 + x] = 0;
			rec[37] = 1;
		}
	} else {
		// chain A = the true chain leaving chunk c (E_c, written by the PHASE 1 launch), chain B =
		// the amnesic walk of chunk c + 1
		L.pos = rec[32], L.ls = rec[33], L.lq = rec[34], L.ll = rec[35], L.paired = rec[36];
		L.b_pos = c_end, L.b_ls = L.b_lq = L.b_ll = L.b_paired = 0;
		L.a_true = 1;
		L.c_end = qlen;
	}
	return true;
}

// Write what this launch owes the record of the finished unit.
template <int PHASE>
V3_FN void v3_finish_unit(const V3Lane &L, u32 *rec, const u32 *col) {
	const u32 base = PHASE == 1 ? 0u : 16u;
#pragma unroll
	for (int x = 0; x < 16; x++) {
		u32 v = col[x * V3_CELL_STRIDE];
		if (x == 0 || x == 5 || x == 10 || x == 15) v += L.sumq;
		if (x == 15) v += L.sumr;
		rec[base + x] = v;
	}
	if (PHASE == 1) {
		rec[32] = L.pos, rec[33] = L.ls, rec[34] = L.lq, rec[35] = L.ll, rec[36] = L.paired;
	} else {
		rec[37] = L.flag;
		if (!L.flag) {
			// the chains reached the end of the query apart: the state the chain of chunk c ends in
			// replaces E_c (this unit was its only reader); the lowest such chunk of a pair carries
			// the pair's true final state (src/process.c:199-211 needs it)
			const bool a = L.a_true != 0u;
			rec[32] = a ? L.pos : L.b_pos, rec[33] = a ? L.ls : L.b_ls, rec[34] = a ? L.lq : L.b_lq, rec[35] = a ? L.ll : L.b_ll;
			rec[36] = a ? L.paired : L.b_paired;
		}
	}
}

// Requests V3_SVC_COOP (a bucket: L.cand_p = first SA index, L.cand2 = number of suffixes) and
// V3_SVC_COOP2 (the two suffixes L.cand_p, L.cand2 of a tag-2 entry): every candidate is compared with
// the query at pos_Q TO THE END of its match, however long; the longest is the match, unique if no
// other is as long (process.c:117-122). Result: job RESOLVED (accounted in the lane's next trip), cand_p = a
// best candidate, len1 = its length, cand2 = unique. Inside repeats (IS elements, rRNA operons) every
// directory lookup lands here; the generic step (one lane, binary search with kilobase compares) took
// about 10 us for each. (A per-lane scan of small buckets, one candidate after the other, was slower than
// this even on repeat-free pools: two dependent loads per candidate while the whole warp waits.)
// This is the serial form (the emulation, and the specification of the kernel's warp-wide form
// v3_coop_scan in walk_v3.cuh: one candidate per lane, all growing together).
V3_FN void v3_scan_full(V3Lane &L, const V3Const &c) {
	const bool two = L.svc == V3_SVC_COOP2;
	const u32 a = L.cand_p, b = L.cand2, count = two ? 2u : b, rem = L.qlen - L.pos;
	u32 best = 0, best_p = 0, best_n = 0;
	for (u32 k = 0; k < count; k++) {
		const u32 p = two ? (k ? b : a) : v3_ld_sa(c.SA + a + k);
		const u32 run = p < c.mid ? c.mid - p : (p == c.mid ? 0u : c.N - p), lim = rem < run ? rem : run;
		u32 len = 0;
		while (len < lim) {
			u64 q0, q1, s0, s1;
			v3_window64(L.q_code, L.pos + len, q0, q1);
			v3_window64(c.s_code, p + len, s0, s1);
			const u32 D = v3_first_diff(q0 ^ s0, q1 ^ s1);
			len += D < lim - len ? D : lim - len;
			if (D < 64u) break;
		}
		best_n = len > best ? 1u : (len == best ? best_n + 1u : best_n);
		best_p = len > best ? p : best_p;
		best = len > best ? len : best;
	}
	V3_STAT(coop_scans);
	L.cand_p = best_p, L.len1 = best, L.cand2 = best_n == 1u ? 1u : 0u, L.job = V3_RESOLVED, L.svc = V3_RUN;
}

// Serve one parked lane. Env supplies what differs between the kernel and the emulation:
//   u64 total; u32 *records; u64 next_unit(); bool open_unit(u64 unit, V3Lane &, u32 *&rec)  (query lookup +
//   v3_begin_unit; false = no work in this unit); void slow_step(V3Lane &, u32 *col, u32 sign).
template <int PHASE, bool QUARTER, class Env>
V3_FN void v3_service(V3Lane &L, const V3Const &c, Env &env, u32 *col, const V3Pend &P) {
#ifndef V3_COOP_IN_WARP_LOOP
	if (L.svc == V3_SVC_COOP || L.svc == V3_SVC_COOP2) {
		v3_scan_full(L, c);
		return;
	}
#endif
	if (L.svc == V3_SVC_SLOW) {
		env.template slow_step<QUARTER>(L, col, (PHASE == 2 && !L.a_true) ? 0xffffffffu : 1u);
		L.svc = V3_RUN;
		return;
	}
	if (L.svc == V3_SVC_FINISH) {
		v3_drain_lane(L, P, col);
		v3_finish_unit<PHASE>(L, env.records + L.unit * ANDI_UNIT_WORDS, col);
		L.svc = V3_SVC_FETCH;
	}
	if (L.svc == V3_SVC_FETCH) {
		L.svc = V3_SVC_DONE;
		for (;;) {
			const u64 unit = env.next_unit();
			if (unit >= env.total) break;
			if (env.template open_unit<PHASE>(unit, L, c, col)) {
				L.unit = unit, L.svc = V3_RUN;
				break;
			}
		}
	}
}

// When does a warp stop to serve its parked lanes? `parked` = lanes with a request, `running` =
// lanes with a window job.
V3_FN bool v3_serve_now(u32 parked, u32 running, u32 trip) {
	return parked != 0u && (running == 0u || parked >= V3_SERVE_BATCH || (trip & (V3_SERVE_EVERY - 1u)) == V3_SERVE_EVERY - 1u);
}
