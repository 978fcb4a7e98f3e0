// andi_b200/csrc/walk_fast.cuh -- k_walk_chunks_fast: the chunked anchor walk of
// walk_kernels.cuh as a phase pipeline with carry-over.
//
// Same units, same records, same results as k_walk_chunks<QUARTER, SPEC>. QUARTER = RAW / JC /
// KIMURA counting without separators is the headline configuration; !QUARTER (LOGDET / ANI)
// needs the pool's prefix-composition table; SPEC (join mode, '!' in subject or queries) reads
// the spec planes next to the code planes where a hint byte says a separator is near.
//
// What changes is the execution shape. In the straightforward kernel every lane runs nested while-loops
// (window compares of different lengths, bucket scans, bitmap probes) and the warp waits for
// its slowest lane at every level: ncu showed 6 of 32 lanes active. A pure lane-level state
// machine (one op per lane per trip) does not fix that either: lanes drift apart and every
// state's code runs with the few lanes that happen to be in it (measured: 7 of 32).
//
// Here one trip of the main loop is a FIXED sequence of phases, each executed at most once:
//
//   BEGIN -> CMP (window 1) -> DIR -> CAND -> CMP (window 2 / candidate) -> SLOW -> DECIDE
//         -> COLS
//
// A lane flows through as many consecutive phases as its walk step needs -- the common steps
// (lucky anchor within two 64-base windows; lucky miss -> directory -> one candidate) complete in ONE
// trip -- and only carries over into the next trip when it needs a phase again (a long
// compare, a second candidate, more gap columns). Nobody waits for a loop, and lanes stay
// aligned at step boundaries, so BEGIN / CMP / DECIDE run with most of the warp.
//
//   BEGIN   chunk/phase bookkeeping, then set up the lucky compare (process.c:86-99)
//   CMP     one 64-base window of a compare (lucky diagonal or directory candidate)
//   DIR     one load from the directory view fdir: absent k-mer -> its prefix length, done;
//           one suffix -> its text position, compare it in this trip; several -> CAND
//   CAND    fetch SA[candidate] (buckets with several suffixes, and their later candidates)
//   SLOW    anything unusual -> longest_match<SPEC>() of walk_kernels.cuh
//   DECIDE  process.c:160-196: pairing, accounting, advance
//   COLS    classify up to 16 gap columns (model.c:309-337)
//
// The single-column gap -- by far the most common one, a lone substitution between two
// anchors -- costs no memory op at all: the class of the column that ended a compare is
// remembered with the anchor (`mm`).
#pragma once
#include "walk_kernels.cuh"

enum : u32 { OP_FETCH = 0, OP_BEGIN, OP_CMP, OP_DIR, OP_CAND, OP_COLS, OP_SLOW, OP_DECIDE, OP_IDLE };

#define ANDI_MM_VALID 0x10u

struct LaneCompare {
	u32 cs, ck, clim, is_cand;	// subject start, matched so far, limit, lucky(0)/candidate(1)
};

struct LaneLookup {
	u32 key, cand, hi, best, best_p, best_cnt, best_mm;
};

struct LaneResult {
	u32 s, len, mm, found;
};

#define ANDI_KEY_SEP 0x80000000u  // L.key: a separator among the first K query characters (SPEC)

// One 64-base window of the current compare; sets `op` when the compare has ended.
// SPEC: bytes are equal when code pair and spec pair are equal (text.cuh); a column with a
// separator on either side never becomes a remembered `mm` class (COLS skips such columns).
template <bool SPEC>
__device__ __forceinline__ void cmp_window(u32 &op, LaneCompare &C, LaneLookup &L, LaneResult &R,
										   const u64 *__restrict__ q_code, const u64 *__restrict__ q_spec,
										   const u64 *__restrict__ s_code, const u64 *__restrict__ s_spec,
										   const unsigned char *__restrict__ q_sep3, const unsigned char *__restrict__ s_sep3,
										   u32 a_pos, u32 qlen, u32 t, int K) {
	// plain window loads (three words = 64 bases per text): they hit L1/L2, and a register cache
	// of the last words costs more instructions and registers than it saves (measured: +14 %
	// throughput without it). 64 bases end 96 % of all compares in one pass.
	// SPEC: one byte per text says whether any of the three words holds a separator (k_sep3);
	// the spec planes are only read where it does
	u32 sep_near = 0;
	if (SPEC) sep_near = __ldg(q_sep3 + ((a_pos + C.ck) >> 5)) | __ldg(s_sep3 + ((C.cs + C.ck) >> 5));
	u64 q0, q1, s0, s1;
	window64(q_code, a_pos + C.ck, q0, q1);
	window64(s_code, C.cs + C.ck, s0, s1);
	if (C.ck == 0 && !C.is_cand) L.key = K > 0 ? kmer_key(q0, K) : 0u;
	// first differing base of the window (a differing 2-bit code has its lowest set bit at 2d or 2d+1)
	u64 x0 = q0 ^ s0, x1 = q1 ^ s1;
	u64 sep0 = 0, sep1 = 0;
	if (SPEC && sep_near) {
		u64 qs0, qs1, ss0, ss1;
		window64(q_spec, a_pos + C.ck, qs0, qs1);
		window64(s_spec, C.cs + C.ck, ss0, ss1);
		x0 |= qs0 ^ ss0, x1 |= qs1 ^ ss1;
		sep0 = qs0 | ss0, sep1 = qs1 | ss1;
		if (C.ck == 0 && !C.is_cand && K > 0 && (qs0 & ((1ULL << (2 * K)) - 1ULL))) L.key |= ANDI_KEY_SEP;
	}
	u32 left = C.clim - C.ck;
	u32 d = x0 ? (u32)(__ffsll((long long)x0) - 1) >> 1 : (x1 ? 32u + ((u32)(__ffsll((long long)x1) - 1) >> 1) : 64u);
	u32 len, mm = 0;
	if (d >= left) {
		len = C.clim;  // the limit (end of query / '#' / end of RS) ends the match
	} else if (d < 64u) {
		len = C.ck + d;
		u64 sw = d < 32u ? s0 : s1, qw = d < 32u ? q0 : q1;
		u32 sh = 2u * (d & 31u);
		mm = ANDI_MM_VALID | ((((u32)(sw >> sh)) & 3u) << 2) | (((u32)(qw >> sh)) & 3u);
		if (SPEC && (((d < 32u ? sep0 : sep1) >> sh) & 1ULL)) mm = 0;
	} else {
		C.ck += 64u;
		return;	 // compare continues
	}
	if (!C.is_cand) {
		if (len >= t) {
			R.found = 1, R.s = C.cs, R.len = len, R.mm = mm;
			op = OP_DECIDE;
		} else if (K > 0 && qlen - a_pos >= (u32)K && !(SPEC && (L.key & ANDI_KEY_SEP))) {
			op = OP_DIR;  // process.c:117: longest match anywhere in RS
		} else {
			op = OP_SLOW;
		}
	} else {
		if (len > L.best)
			L.best = len, L.best_cnt = 1, L.best_p = C.cs, L.best_mm = mm;
		else if (len == L.best)
			L.best_cnt++;
		L.cand++;
		if (L.cand < L.hi) {
			op = OP_CAND;
		} else if (L.best >= (u32)K) {
			R.found = (L.best_cnt == 1 && L.best >= t) ? 1u : 0u;
			R.s = L.best_p, R.len = L.best, R.mm = L.best_mm;
			op = OP_DECIDE;
		} else {
			// cannot happen (the directory counts only suffixes that carry the whole k-mer):
			// leave it to the generic search
			op = OP_SLOW;
		}
	}
}

#ifndef ANDI_FAST_BLOCKS_PER_SM
#define ANDI_FAST_BLOCKS_PER_SM 4
#endif
template <bool QUARTER, bool SPEC>
__global__ void __launch_bounds__(ANDI_WALK_THREADS, (QUARTER && !SPEC) ? ANDI_FAST_BLOCKS_PER_SM : ANDI_FAST_BLOCKS_PER_SM - 1)
k_walk_chunks_fast(const SubjectIndex S, const QueryView *__restrict__ queries, const u32 *__restrict__ query_ids,
				   u32 nq, u32 chunk, u32 cpq, u32 threshold, u32 *__restrict__ records,
				   unsigned long long *__restrict__ next_unit) {
	__shared__ u32 cells[2][16][ANDI_WALK_THREADS];
	const u32 tid = threadIdx.x;
	const unsigned long long total = (unsigned long long)nq * cpq;
	// units are handed out dynamically: their cost varies with the divergence of the pair
	unsigned long long unit = atomicAdd(next_unit, 1ULL);
	const u32 t = threshold, N = S.rs.len, mid = S.rs.mid, border = N / 2;
	const int K = S.K;
	const u64 *__restrict__ s_code = S.rs.code;
	const u64 *__restrict__ s_spec = S.rs.spec;

	// ---- lane state
	u32 op = OP_FETCH;
	const u64 *q_code = nullptr;
	u32 qlen = 0, c_end = 0, c2_end = 0, phase = 1, a_true = 1;
	// chain A (the one being advanced) and chain B (phase 2 only); *_pm = paired | mm << 1
	u32 a_pos = 0, a_ls = 0, a_lq = 0, a_ll = 0, a_pm = 0;
	u32 b_pos = 0, b_ls = 0, b_lq = 0, b_ll = 0, b_pm = 0;
	u32 set = 0, sign = 1;
	LaneCompare C = {0, 0, 0, 0};
	LaneLookup L = {0, 0, 0, 0, 0, 0, 0};
	LaneResult R = {0, 0, 0, 0};
	u32 cols_s = 0, cols_q = 0, cols_left = 0;

	for (;;) {
		// ------------------------------------------------------------ FETCH
		if (op == OP_FETCH) {
			op = OP_IDLE;
			while (unit < total) {
				u32 k, c;
				if (total <= 0xffffffffULL) {  // 32-bit division: a fifth of the instructions
					k = (u32)unit / cpq, c = (u32)unit - k * cpq;
				} else {
					k = (u32)(unit / cpq), c = (u32)(unit % cpq);
				}
				u32 qid = query_ids ? query_ids[k] : k;
				u32 ql = queries[qid].t.len;
				unsigned long long start = (unsigned long long)c * chunk;
				if (qid != S.self && start < ql) {
					q_code = queries[qid].t.code;
					qlen = ql;
					a_pos = (u32)start, a_ls = a_lq = a_ll = a_pm = 0;
					c_end = (u32)min((unsigned long long)ql, start + chunk);
					c2_end = (u32)min((unsigned long long)ql, start + 2ULL * chunk);
					phase = 1, set = 0, sign = 1;
#pragma unroll
					for (int x = 0; x < 16; x++) cells[0][x][tid] = 0, cells[1][x][tid] = 0;
					op = OP_BEGIN;
					break;
				}
				unit = atomicAdd(next_unit, 1ULL);
			}
		}
		if (__all_sync(0xffffffffu, op == OP_IDLE)) break;

		__syncwarp();  // phase boundary: every lane of the warp re-converges here
		// ------------------------------------------------------------ BEGIN
		if (op == OP_BEGIN) {
			bool finished = false;
			u32 flag = 1;
			u32 *rec = records + unit * ANDI_UNIT_WORDS;
			if (phase == 1 && a_pos >= c_end) {
				rec[32] = a_pos, rec[33] = a_ls, rec[34] = a_lq, rec[35] = a_ll, rec[36] = a_pm & 1u;
				if (c_end >= qlen) {
					finished = true;
				} else {
					phase = 2, set = 1, a_true = 1;
					b_pos = c_end, b_ls = b_lq = b_ll = b_pm = 0;
				}
			}
			if (phase == 2 && !finished) {
				u32 t_pos = a_true ? a_pos : b_pos, p_pos = a_true ? b_pos : a_pos;
				bool same = a_pos == b_pos && a_ls == b_ls && a_lq == b_lq && a_ll == b_ll && ((a_pm ^ b_pm) & 1u) == 0;
				if (same) {
					finished = true;
				} else if (t_pos >= c2_end || p_pos >= c2_end) {
					finished = true, flag = 0;
				} else {
					bool step_true = t_pos <= p_pos;
					if (step_true != (a_true != 0)) {
						u32 x;
						x = a_pos, a_pos = b_pos, b_pos = x;
						x = a_ls, a_ls = b_ls, b_ls = x;
						x = a_lq, a_lq = b_lq, b_lq = x;
						x = a_ll, a_ll = b_ll, b_ll = x;
						x = a_pm, a_pm = b_pm, b_pm = x;
						a_true ^= 1u;
					}
					sign = a_true ? 1u : 0xffffffffu;
				}
			}
			if (finished) {
#pragma unroll
				for (int x = 0; x < 16; x++) rec[x] = cells[0][x][tid];
#pragma unroll
				for (int x = 0; x < 16; x++) rec[16 + x] = flag ? cells[1][x][tid] : 0u;
				rec[37] = flag;
				unit = atomicAdd(next_unit, 1ULL);
				op = OP_FETCH;
			} else {
				// process.c:86-99: the diagonal of the previous anchor, if close enough
				u32 rem = qlen - a_pos;
				u32 advance = a_pos - a_lq;
				u32 gap = advance - a_ll;
				u32 guess = a_ls + advance;
				if (guess < N && gap <= t) {
					C.cs = guess;
					u32 run = SPEC ? N - guess : (guess < mid ? mid - guess : (guess == mid ? 0u : N - guess));
					C.clim = min(rem, run);
				} else {
					C.cs = 0, C.clim = 0;  // no lucky attempt: the window op only fetches the query k-mer
				}
				C.ck = 0, C.is_cand = 0;
				op = OP_CMP;
			}
		}

		__syncwarp();  // phase boundary: every lane of the warp re-converges here
		// ------------------------------------------------------------ CMP, first window of the trip
		if (op == OP_CMP) cmp_window<SPEC>(op, C, L, R, q_code, q_code + S.qspec_delta, s_code, s_spec,
											   S.qsep3_base + (q_code - S.qcode_base), S.s_sep3, a_pos, qlen, t, K);

		__syncwarp();  // phase boundary: every lane of the warp re-converges here
		// ------------------------------------------------------------ DIR
		if (op == OP_DIR) {
			u64 fe = __ldg(S.fdir + L.key);
			u32 tag = ANDI_FDIR_TAG(fe);
			L.best = 0, L.best_cnt = 0, L.best_p = 0, L.best_mm = 0;
			if (tag == 0u) {  // absent k-mer: only the length of the match matters
				R.found = 0, R.len = (u32)fe, R.s = 0, R.mm = 0;
				op = OP_DECIDE;
			} else if (tag == 1u) {	 // one suffix starts with this k-mer: compare it right away
				u32 p = (u32)fe & 0x7fffffffu, rem = qlen - a_pos;
				u32 run = SPEC ? N - p : (p < mid ? mid - p : (p == mid ? 0u : N - p));
				L.cand = 0, L.hi = 1;
				C.cs = p, C.ck = 0, C.clim = min(rem, run), C.is_cand = 1;
				op = OP_CMP;
			} else {
				// two or more suffixes: this kernel walks them through the suffix array
				u64 de = __ldg(S.dir + L.key);
				u32 t0 = (u32)de, cnt = (u32)(de >> 32);
				if (cnt <= ANDI_SCAN_MAX) {
					L.cand = t0, L.hi = t0 + cnt;
					op = OP_CAND;
				} else {
					op = OP_SLOW;
				}
			}
		}

		__syncwarp();  // phase boundary: every lane of the warp re-converges here
		// ------------------------------------------------------------ CAND
		if (op == OP_CAND) {
			u32 p = __ldg(S.SA + L.cand), rem = qlen - a_pos;
			u32 run = SPEC ? N - p : (p < mid ? mid - p : (p == mid ? 0u : N - p));
			C.cs = p, C.ck = 0, C.clim = min(rem, run), C.is_cand = 1;
			op = OP_CMP;
		}

		__syncwarp();  // phase boundary: every lane of the warp re-converges here
		// ------------------------------------------------------------ CMP, second window / candidate
		if (op == OP_CMP) cmp_window<SPEC>(op, C, L, R, q_code, q_code + S.qspec_delta, s_code, s_spec,
											   S.qsep3_base + (q_code - S.qcode_base), S.s_sep3, a_pos, qlen, t, K);

		__syncwarp();  // phase boundary: every lane of the warp re-converges here
		// ------------------------------------------------------------ SLOW (rare)
		if (op == OP_SLOW) {
			TextView qv;
			qv.code = q_code, qv.spec = SPEC ? q_code + S.qspec_delta : nullptr, qv.len = qlen, qv.mid = 0xffffffffu;
			MatchResult m = longest_match<SPEC>(S, qv, a_pos, qlen - a_pos);
			R.found = (m.unique && m.len >= t) ? 1u : 0u;
			R.len = m.len, R.mm = 0;
			R.s = R.found ? __ldg(S.SA + m.at) : 0u;
			op = OP_DECIDE;
		}

		__syncwarp();  // phase boundary: every lane of the warp re-converges here
		// ------------------------------------------------------------ DECIDE: process.c:160-196
		if (op == OP_DECIDE) {
			bool need_cols = false;
			if (R.found) {
				u32 *col = &cells[set][0][tid];
				u32 end_s = a_ls + a_ll, end_q = a_lq + a_ll;
				bool pairs = R.s > end_s && (a_pos - end_q) == (R.s - end_s) && ((R.s < border) == (a_ls < border));
				bool count_last = pairs || (a_pm & 1u) || a_ll >= 2u * t;
				if (count_last && QUARTER) {  // model.c:247-254
					u32 f = (a_ll >> 2) * sign;
					col[0 * ANDI_WALK_THREADS] += f;
					col[5 * ANDI_WALK_THREADS] += f;
					col[10 * ANDI_WALK_THREADS] += f;
					col[15 * ANDI_WALK_THREADS] += f + (a_ll & 3u) * sign;
				}
				if (count_last && !QUARTER) {
					// model.c:259-278: composition of the query slice [a_lq, a_lq + a_ll), in O(1)
					// from the per-word prefix composition of the pool (k_comp_prefix): two table
					// entries and the two partial words at the slice ends, one memory round.
					const uint4 *cp = S.qcomp_base + (q_code - S.qcode_base);
					u32 b0 = a_lq, b1 = a_lq + a_ll;
					uint4 p0 = __ldg(cp + (b0 >> 5)), p1 = __ldg(cp + (b1 >> 5));
					u64 w0 = __ldg(q_code + (b0 >> 5)), w1 = __ldg(q_code + (b1 >> 5));
					u64 m0 = ANDI_EVEN_BITS & ((1ULL << (2u * (b0 & 31u))) - 1ULL);
					u64 m1 = ANDI_EVEN_BITS & ((1ULL << (2u * (b1 & 31u))) - 1ULL);
					if (SPEC) {	 // separators are not counted
						m0 &= ~__ldg(q_code + S.qspec_delta + (b0 >> 5));
						m1 &= ~__ldg(q_code + S.qspec_delta + (b1 >> 5));
					}
					u64 l0 = w0 & m0, h0 = (w0 >> 1) & m0, l1 = w1 & m1, h1 = (w1 >> 1) & m1;
					u32 na = p1.x - p0.x + (u32)__popcll(m1 & ~l1 & ~h1) - (u32)__popcll(m0 & ~l0 & ~h0);
					u32 nc = p1.y - p0.y + (u32)__popcll(l1 & ~h1) - (u32)__popcll(l0 & ~h0);
					u32 ng = p1.z - p0.z + (u32)__popcll(h1 & ~l1) - (u32)__popcll(h0 & ~l0);
					u32 nt = p1.w - p0.w + (u32)__popcll(h1 & l1) - (u32)__popcll(h0 & l0);
					col[0 * ANDI_WALK_THREADS] += na * sign;
					col[5 * ANDI_WALK_THREADS] += nc * sign;
					col[10 * ANDI_WALK_THREADS] += ng * sign;
					col[15 * ANDI_WALK_THREADS] += nt * sign;
				}
				if (pairs) {
					u32 g = a_pos - end_q;
					u32 lm = a_pm >> 1;
					if (g == 1u && (lm & ANDI_MM_VALID)) {
						col[(lm & 15u) * ANDI_WALK_THREADS] += sign;
					} else {
						cols_s = end_s, cols_q = end_q, cols_left = g;
						need_cols = true;
					}
				}
				a_ls = R.s, a_lq = a_pos, a_ll = R.len;
				a_pm = (pairs ? 1u : 0u) | (R.mm << 1);
			}
			a_pos += R.len + 1u;
			op = need_cols ? OP_COLS : OP_BEGIN;
		}

		__syncwarp();  // phase boundary: every lane of the warp re-converges here
		// ------------------------------------------------------------ COLS: model.c:309-337
		if (op == OP_COLS) {
			// 16 columns per pass in 32-bit arithmetic: a gap behind a lucky anchor is at most
			// `threshold` columns wide, and half-width masks / popcounts halve the phase
			u32 span = min(16u, cols_left);
			u32 qw = window16(q_code, cols_q), sw = window16(s_code, cols_s);
			u32 valid = span == 16u ? 0x55555555u : (0x55555555u & ((1u << (2u * span)) - 1u));
			if (cols_s <= mid && mid - cols_s < span) valid &= ~(1u << (2u * (mid - cols_s)));	// '#' column
			if (SPEC && (__ldg(S.s_sep3 + (cols_s >> 5)) | __ldg(S.qsep3_base + (q_code - S.qcode_base) + (cols_q >> 5))))
				valid &= ~(window16(s_spec, cols_s) | window16(q_code + S.qspec_delta, cols_q));
			u32 x = qw ^ sw;
			u32 neq = (x | (x >> 1)) & valid, eq = valid & ~neq;
			u32 lo = qw & 0x55555555u, hb = (qw >> 1) & 0x55555555u;
			u32 *col = &cells[set][0][tid];
			col[0 * ANDI_WALK_THREADS] += (u32)__popc(eq & ~hb & ~lo) * sign;
			col[5 * ANDI_WALK_THREADS] += (u32)__popc(eq & ~hb & lo) * sign;
			col[10 * ANDI_WALK_THREADS] += (u32)__popc(eq & hb & ~lo) * sign;
			col[15 * ANDI_WALK_THREADS] += (u32)__popc(eq & hb & lo) * sign;
			while (neq) {
				u32 d2 = (u32)(__ffs((int)neq) - 1);
				neq &= neq - 1;
				u32 cls = (((sw >> d2) & 3u) << 2) | ((qw >> d2) & 3u);
				col[cls * ANDI_WALK_THREADS] += sign;
			}
			cols_s += span, cols_q += span, cols_left -= span;
			if (cols_left == 0) op = OP_BEGIN;
		}
	}
}
