// andi_b200/csrc/walk_fast.cuh -- k_walk_chunks_fast: the chunked anchor walk of
// walk_kernels.cuh as a lane-level micro-op machine.
//
// Same units, same records, same results as k_walk_chunks<QUARTER=true, SPEC=false> (RAW / JC /
// KIMURA counting, no '!' in subject or queries -- the headline configuration). What changes
// is the execution shape. In the straightforward kernel every lane runs nested loops (window
// compares of different lengths, bucket scans, bitmap probes) and the warp waits for its
// slowest lane at every level: ncu showed 6 of 32 lanes active. Here every lane carries a tiny
// program counter `op`; one trip of the single main loop performs for each lane exactly ONE
// memory round -- a 32-base window op on two streams, or a directory / suffix-array / bitmap
// probe -- followed by register-only transitions. Lanes never wait for another lane's loop.
//
//   OP_BEGIN  chunk/phase bookkeeping, then set up the lucky compare (process.c:86-99)
//   OP_CMP    one 32-base window of a compare (lucky diagonal or directory candidate)
//   OP_DIR    k-mer directory probe        OP_CAND  fetch SA[candidate]
//   OP_BITS   presence bitmaps (match shorter than K)
//   OP_COLS   classify up to 32 gap columns (model.c:309-337)
//   OP_SLOW   anything unusual -> longest_match<false>() of walk_kernels.cuh
//   DECIDE    (not a memory op) process.c:160-196: pairing, accounting, advance
//
// The single-column gap -- by far the most common one, a lone substitution between two
// anchors -- costs no memory op at all: the class of the column that ended a compare is
// remembered with the anchor (`mm`).
#pragma once
#include "walk_kernels.cuh"

enum : u32 { OP_FETCH = 0, OP_BEGIN, OP_CMP, OP_DIR, OP_CAND, OP_BITS, OP_COLS, OP_SLOW, OP_DECIDE, OP_IDLE };

#define ANDI_MM_VALID 0x10u

struct WordCache {
	u32 idx;	 // word index of w0 (0xfffffff0 = empty; idx + 1 must not wrap to a valid index)
	u64 w0, w1;	 // words idx and idx+1
};

__global__ void __launch_bounds__(ANDI_WALK_THREADS, 3)
k_walk_chunks_fast(const SubjectIndex S, const QueryView *__restrict__ queries, const u32 *__restrict__ query_ids,
				   u32 nq, u32 chunk, u32 cpq, u32 threshold, u32 *__restrict__ records) {
	__shared__ u32 cells[2][16][ANDI_WALK_THREADS];
	const u32 tid = threadIdx.x;
	const unsigned long long total = (unsigned long long)nq * cpq;
	const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
	unsigned long long unit = (unsigned long long)blockIdx.x * blockDim.x + tid;
	const u32 t = threshold, N = S.rs.len, mid = S.rs.mid, border = N / 2;
	const int K = S.K;
	const u64 *__restrict__ s_code = S.rs.code;

	// ---- lane state
	u32 op = OP_FETCH;
	const u64 *q_code = nullptr;
	u32 qlen = 0, c_end = 0, c2_end = 0, phase = 1, a_true = 1;
	// chain A (the one being advanced) and chain B (phase 2 only); *_pm = paired | mm << 1
	u32 a_pos = 0, a_ls = 0, a_lq = 0, a_ll = 0, a_pm = 0;
	u32 b_pos = 0, b_ls = 0, b_lq = 0, b_ll = 0, b_pm = 0;
	u32 set = 0, sign = 1;
	// current compare
	u32 cs = 0, ck = 0, clim = 0, cand_cmp = 0;
	// current lookup
	u32 key = 0, cand = 0, hi = 0, best = 0, best_p = 0, best_cnt = 0, best_mm = 0, bits_m = 0;
	// result handed to DECIDE
	u32 cur_s = 0, cur_len = 0, cur_mm = 0, found = 0;
	// gap columns
	u32 cols_s = 0, cols_q = 0, cols_left = 0;
	WordCache qc, sc;
	qc.idx = sc.idx = 0xfffffff0u, qc.w0 = qc.w1 = sc.w0 = sc.w1 = 0;

	for (;;) {
		// ------------------------------------------------------------ (0) unit fetch
		if (op == OP_FETCH) {
			op = OP_IDLE;
			while (unit < total) {
				u32 k = (u32)(unit / cpq), c = (u32)(unit % cpq);
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
					qc.idx = 0xfffffff0u;
#pragma unroll
					for (int x = 0; x < 16; x++) cells[0][x][tid] = 0, cells[1][x][tid] = 0;
					op = OP_BEGIN;
					break;
				}
				unit += stride;
			}
		}
		if (__all_sync(0xffffffffu, op == OP_IDLE)) break;

		// ------------------------------------------------------------ (1) step set-up
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
				unit += stride;
				op = OP_FETCH;
			} else {
				// process.c:86-99: the diagonal of the previous anchor, if close enough
				u32 rem = qlen - a_pos;
				u32 advance = a_pos - a_lq;
				u32 gap = advance - a_ll;
				u32 guess = a_ls + advance;
				if (guess < N && gap <= t) {
					cs = guess;
					u32 run = guess < mid ? mid - guess : (guess == mid ? 0u : N - guess);
					clim = min(rem, run);
				} else {
					cs = 0, clim = 0;  // no lucky attempt: the window op only fetches the query k-mer
				}
				ck = 0, cand_cmp = 0;
				op = OP_CMP;
			}
		}

		// ------------------------------------------------------------ (2) one memory round
		const bool text_op = (op == OP_CMP) | (op == OP_COLS);
		u32 ta = 0, tb = 0;	 // query / subject position of the window
		if (text_op) {
			ta = op == OP_CMP ? a_pos + ck : cols_q;
			tb = op == OP_CMP ? cs + ck : cols_s;
			u32 iq = ta >> 5, is = tb >> 5;
			bool qh0 = iq == qc.idx, qh1 = iq == qc.idx + 1u;
			bool sh0 = is == sc.idx, sh1 = is == sc.idx + 1u;
			if (qh1) qc.w0 = qc.w1;
			if (sh1) sc.w0 = sc.w1;
			if (!(qh0 | qh1)) qc.w0 = __ldg(q_code + iq);
			if (!qh0) qc.w1 = __ldg(q_code + iq + 1);
			if (!(sh0 | sh1)) sc.w0 = __ldg(s_code + is);
			if (!sh0) sc.w1 = __ldg(s_code + is + 1);
			qc.idx = iq, sc.idx = is;
		}
		u32 t0 = 0, t1 = 0, t2 = 0;
		if (op == OP_DIR) {
			t0 = __ldg(S.dir + key);
			t1 = __ldg(S.dir + key + 1);
		} else if (op == OP_CAND) {
			t0 = __ldg(S.SA + cand);
		} else if (op == OP_BITS) {
			u32 x0 = key >> (2 * (K - (int)bits_m));
			t0 = __ldg(S.present.bits + S.present.offset[bits_m] + (x0 >> 5)) >> (x0 & 31u);
			if (bits_m >= 2) {
				u32 x1 = x0 >> 2;
				t1 = __ldg(S.present.bits + S.present.offset[bits_m - 1] + (x1 >> 5)) >> (x1 & 31u);
			}
			if (bits_m >= 3) {
				u32 x2 = x0 >> 4;
				t2 = __ldg(S.present.bits + S.present.offset[bits_m - 2] + (x2 >> 5)) >> (x2 & 31u);
			}
		}

		// ------------------------------------------------------------ (3) consume
		u32 nop = op;
		if (op == OP_CMP) {
			u32 shq = (ta & 31u) * 2u, shs = (tb & 31u) * 2u;
			u64 qw = shq ? (qc.w0 >> shq) | (qc.w1 << (64u - shq)) : qc.w0;
			u64 sw = shs ? (sc.w0 >> shs) | (sc.w1 << (64u - shs)) : sc.w0;
			if (ck == 0 && !cand_cmp) key = K > 0 ? kmer_key(qw, K) : 0u;
			u32 left = clim - ck;
			u64 x = qw ^ sw;
			x = (x | (x >> 1)) & ANDI_EVEN_BITS;
			if (left < 32u) x |= 1ULL << (2u * left);
			bool ended = false;
			u32 len = 0, mm = 0;
			if (x) {
				u32 d = (u32)(__ffsll((long long)x) - 1) >> 1;
				len = ck + d;
				ended = true;
				if (d < left) mm = ANDI_MM_VALID | ((((u32)(sw >> (2u * d))) & 3u) << 2) | (((u32)(qw >> (2u * d))) & 3u);
			} else {
				ck += 32u;
				if (ck == clim) ended = true, len = clim;
			}
			if (ended) {
				if (!cand_cmp) {
					if (len >= t) {
						found = 1, cur_s = cs, cur_len = len, cur_mm = mm;
						nop = OP_DECIDE;
					} else if (K > 0 && qlen - a_pos >= (u32)K) {
						nop = OP_DIR;  // process.c:117: longest match anywhere in RS
					} else {
						nop = OP_SLOW;
					}
				} else {
					if (len > best)
						best = len, best_cnt = 1, best_p = cs, best_mm = mm;
					else if (len == best)
						best_cnt++;
					cand++;
					if (cand < hi) {
						nop = OP_CAND;
					} else if (best >= (u32)K) {
						found = (best_cnt == 1 && best >= t) ? 1u : 0u;
						cur_s = best_p, cur_len = best, cur_mm = best_mm;
						nop = OP_DECIDE;
					} else {
						bits_m = (u32)(K - 1);
						nop = OP_BITS;
					}
				}
			}
		} else if (op == OP_DIR) {
			if (t1 > t0) {
				if (t1 - t0 <= ANDI_SCAN_MAX) {
					cand = t0, hi = t1, best = 0, best_cnt = 0, best_p = 0, best_mm = 0;
					nop = OP_CAND;
				} else {
					nop = OP_SLOW;
				}
			} else {
				bits_m = (u32)(K - 1);
				nop = OP_BITS;
			}
		} else if (op == OP_CAND) {
			u32 p = t0, rem = qlen - a_pos;
			u32 run = p < mid ? mid - p : (p == mid ? 0u : N - p);
			cs = p, ck = 0, clim = min(rem, run), cand_cmp = 1;
			nop = OP_CMP;
		} else if (op == OP_BITS) {
			u32 l = 0;
			bool done = true;
			if (t0 & 1u)
				l = bits_m;
			else if (bits_m >= 2 && (t1 & 1u))
				l = bits_m - 1;
			else if (bits_m >= 3 && (t2 & 1u))
				l = bits_m - 2;
			else if (bits_m > 3)
				bits_m -= 3, done = false;
			if (done) {
				found = 0, cur_len = l, cur_s = 0, cur_mm = 0;
				nop = OP_DECIDE;
			}
		} else if (op == OP_COLS) {
			u32 span = min(32u, cols_left);
			u32 shq = (ta & 31u) * 2u, shs = (tb & 31u) * 2u;
			u64 qw = shq ? (qc.w0 >> shq) | (qc.w1 << (64u - shq)) : qc.w0;
			u64 sw = shs ? (sc.w0 >> shs) | (sc.w1 << (64u - shs)) : sc.w0;
			u64 valid = span == 32u ? ANDI_EVEN_BITS : (ANDI_EVEN_BITS & ((1ULL << (2u * span)) - 1ULL));
			if (cols_s <= mid && mid - cols_s < span) valid &= ~(1ULL << (2u * (mid - cols_s)));  // '#' column
			u64 x = qw ^ sw;
			u64 neq = (x | (x >> 1)) & valid, eq = valid & ~neq;
			u64 lo = qw & ANDI_EVEN_BITS, hb = (qw >> 1) & ANDI_EVEN_BITS;
			u32 *col = &cells[set][0][tid];
			col[0 * ANDI_WALK_THREADS] += (u32)__popcll(eq & ~hb & ~lo) * sign;
			col[5 * ANDI_WALK_THREADS] += (u32)__popcll(eq & ~hb & lo) * sign;
			col[10 * ANDI_WALK_THREADS] += (u32)__popcll(eq & hb & ~lo) * sign;
			col[15 * ANDI_WALK_THREADS] += (u32)__popcll(eq & hb & lo) * sign;
			while (neq) {
				u32 d2 = (u32)(__ffsll((long long)neq) - 1);
				neq &= neq - 1;
				u32 cls = ((((u32)(sw >> d2)) & 3u) << 2) | (((u32)(qw >> d2)) & 3u);
				col[cls * ANDI_WALK_THREADS] += sign;
			}
			cols_s += span, cols_q += span, cols_left -= span;
			if (cols_left == 0) nop = OP_BEGIN;
		} else if (op == OP_SLOW) {
			TextView qv;
			qv.code = q_code, qv.spec = nullptr, qv.len = qlen, qv.mid = 0xffffffffu;
			MatchResult m = longest_match<false>(S, qv, a_pos, qlen - a_pos);
			found = (m.unique && m.len >= t) ? 1u : 0u;
			cur_len = m.len, cur_mm = 0;
			cur_s = found ? __ldg(S.SA + m.at) : 0u;
			nop = OP_DECIDE;
		}

		// ------------------------------------------------------------ (4) process.c:160-196
		if (nop == OP_DECIDE) {
			bool need_cols = false;
			if (found) {
				u32 *col = &cells[set][0][tid];
				u32 end_s = a_ls + a_ll, end_q = a_lq + a_ll;
				bool pairs = cur_s > end_s && (a_pos - end_q) == (cur_s - end_s) && ((cur_s < border) == (a_ls < border));
				bool count_last = pairs || (a_pm & 1u) || a_ll >= 2u * t;
				if (count_last) {  // model.c:247-254
					u32 f = (a_ll >> 2) * sign;
					col[0 * ANDI_WALK_THREADS] += f;
					col[5 * ANDI_WALK_THREADS] += f;
					col[10 * ANDI_WALK_THREADS] += f;
					col[15 * ANDI_WALK_THREADS] += f + (a_ll & 3u) * sign;
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
				a_ls = cur_s, a_lq = a_pos, a_ll = cur_len;
				a_pm = (pairs ? 1u : 0u) | (cur_mm << 1);
			}
			a_pos += cur_len + 1u;
			nop = need_cols ? OP_COLS : OP_BEGIN;
		}
		op = nop;
	}
}
