// andi_b200/csrc/experimental/walk_binned_phases.h -- EXPERIMENTAL, see walk_binned.cuh.
//
// The per-unit phase logic of the binned walk, written against a small set of primitives so that
// the SAME text compiles into the CUDA kernel (walk_binned.cuh) and into a serial host emulation
// (emu_binned.cpp, driven by tests/test_binned_emulation.py against the oracle). Whoever includes
// this file provides:
//   BIN_FN, BIN_OUTLINE_FN                  function qualifiers (inlined / one out-of-line copy)
//   SubjectIndex / QueryView / TextView      with the field names of walk_kernels.cuh
//   window64, window16, kmer_key             as in text.cuh
//   bin_ld64, bin_ld32                       read-only loads
//   bin_ffs64, bin_ffs32, bin_popc32         index of the lowest set bit / population count
//   bin_atomic_inc(u32 *)                    fetch-and-increment of a queue counter
//   bin_next_unit(unsigned long long *)      fetch-and-increment of the global unit dispenser
//   bin_slow_lookup(S, q_code, qlen, pos, len, unique, at)   the generic longest-match search
//   ANDI_UNIT_WORDS, ANDI_SCAN_MAX, ANDI_FDIR_TAG            as in walk_kernels.cuh / sa_bucket.cuh
#pragma once

#ifndef ANDI_BIN_SLOTS
#define ANDI_BIN_SLOTS 320	// units (and threads) per CTA; 248 bytes of shared memory per unit
#endif
#define ANDI_BIN_MM_VALID 0x10u

enum : u32 { BQ_FETCH = 0, BQ_CMP, BQ_DIR, BQ_CAND, BQ_SLOW, BQ_DECIDE, BQ_COLS, BQ_N };

// flags word of a slot
#define BF_PHASE2 1u	  // boundary replay (accumulator set 1)
#define BF_A_TRUE 2u	  // chain A is the true chain (phase 2)
#define BF_IS_CAND 4u	  // the running compare is a directory candidate, not the lucky diagonal
#define BF_FOUND 8u		  // DECIDE: an anchor was found
#define BF_CNT_SHIFT 4u	  // bits 4..5: number of candidates that reached `best` (0, 1, 2 = several)

struct BinShared {
	// walk state, one column per slot (names as in walk_fast.cuh)
	u32 qoff[ANDI_BIN_SLOTS];  // word offset of the query's code plane from S.qcode_base
	u32 qlen[ANDI_BIN_SLOTS], c_end[ANDI_BIN_SLOTS], flags[ANDI_BIN_SLOTS], unit[ANDI_BIN_SLOTS];
	u32 a_pos[ANDI_BIN_SLOTS], a_ls[ANDI_BIN_SLOTS], a_lq[ANDI_BIN_SLOTS], a_ll[ANDI_BIN_SLOTS], a_pm[ANDI_BIN_SLOTS];
	u32 b_pos[ANDI_BIN_SLOTS], b_ls[ANDI_BIN_SLOTS], b_lq[ANDI_BIN_SLOTS], b_ll[ANDI_BIN_SLOTS], b_pm[ANDI_BIN_SLOTS];
	u32 cs[ANDI_BIN_SLOTS], ck[ANDI_BIN_SLOTS], clim[ANDI_BIN_SLOTS];  // compare; COLS reuses cs = cols_s, ck = cols_left
	u32 key[ANDI_BIN_SLOTS];										   // query k-mer, then the candidate cursor
	u32 hi[ANDI_BIN_SLOTS], best[ANDI_BIN_SLOTS], best_p[ANDI_BIN_SLOTS], best_mm[ANDI_BIN_SLOTS];
	u32 cells[2][16][ANDI_BIN_SLOTS];
	// the two queue sets
	unsigned short queue[2][BQ_N][ANDI_BIN_SLOTS];
	u32 count[2][BQ_N];
	u32 head[BQ_N];
};

BIN_FN void bin_push(BinShared &sh, u32 nxt, u32 q, u32 slot) {
	u32 at = bin_atomic_inc(&sh.count[nxt][q]);
	sh.queue[nxt][q][at] = (unsigned short)slot;
}

// Scheduling rule shared by the kernel and the emulation: a queue holding less than a full batch
// is carried over unprocessed (its units wait, the other queues keep the warps busy) for at most
// ANDI_BIN_MAX_WAIT super-steps, unless it holds everything that is left.
#ifndef ANDI_BIN_MAX_WAIT
#define ANDI_BIN_MAX_WAIT 3
#endif
BIN_FN bool bin_defer(u32 have, u32 waited, u32 pending) {
	return have > 0 && have < 32u && waited < ANDI_BIN_MAX_WAIT && have < pending;
}

struct BinConst {
	u32 t, N, mid, border, chunk, cpq;
	int K;
};

// src/process.c:86-99 and the chunk / boundary-replay bookkeeping of walk_fast.cuh's BEGIN.
// Returns the queue the slot goes to (BQ_CMP, or BQ_FETCH once the unit's record is written).
// (one out-of-line copy: FETCH, DECIDE and COLS all end in it, and it holds the 38-word record write)
BIN_OUTLINE_FN u32 bin_begin(BinShared &sh, u32 s, const BinConst &c, u32 *__restrict__ records) {
	u32 fl = sh.flags[s];
	u32 a_pos = sh.a_pos[s], a_ls = sh.a_ls[s], a_lq = sh.a_lq[s], a_ll = sh.a_ll[s], a_pm = sh.a_pm[s];
	const u32 qlen = sh.qlen[s], c_end = sh.c_end[s];
	const u32 c2_end = (u32)min((unsigned long long)qlen, (unsigned long long)c_end + c.chunk);
	bool finished = false;
	u32 flag = 1;
	u32 *rec = records + (unsigned long long)sh.unit[s] * ANDI_UNIT_WORDS;
	if (!(fl & BF_PHASE2) && a_pos >= c_end) {
		rec[32] = a_pos, rec[33] = a_ls, rec[34] = a_lq, rec[35] = a_ll, rec[36] = a_pm & 1u;
		if (c_end >= qlen) {
			finished = true;
		} else {
			fl |= BF_PHASE2 | BF_A_TRUE;
			sh.b_pos[s] = c_end, sh.b_ls[s] = 0, sh.b_lq[s] = 0, sh.b_ll[s] = 0, sh.b_pm[s] = 0;
		}
	}
	if ((fl & BF_PHASE2) && !finished) {
		u32 b_pos = sh.b_pos[s], b_ls = sh.b_ls[s], b_lq = sh.b_lq[s], b_ll = sh.b_ll[s], b_pm = sh.b_pm[s];
		bool a_true = (fl & BF_A_TRUE) != 0;
		u32 t_pos = a_true ? a_pos : b_pos, p_pos = a_true ? b_pos : a_pos;
		bool same = a_pos == b_pos && a_ls == b_ls && a_lq == b_lq && a_ll == b_ll && ((a_pm ^ b_pm) & 1u) == 0;
		if (same) {
			finished = true;
		} else if (t_pos >= c2_end || p_pos >= c2_end) {
			finished = true, flag = 0;
		} else if ((t_pos <= p_pos) != a_true) {
			// advance the chain that is behind: it becomes chain A
			sh.b_pos[s] = a_pos, sh.b_ls[s] = a_ls, sh.b_lq[s] = a_lq, sh.b_ll[s] = a_ll, sh.b_pm[s] = a_pm;
			a_pos = b_pos, a_ls = b_ls, a_lq = b_lq, a_ll = b_ll, a_pm = b_pm;
			sh.a_pos[s] = a_pos, sh.a_ls[s] = a_ls, sh.a_lq[s] = a_lq, sh.a_ll[s] = a_ll, sh.a_pm[s] = a_pm;
			fl ^= BF_A_TRUE;
		}
	}
	if (finished) {
#pragma unroll
		for (int x = 0; x < 16; x++) rec[x] = sh.cells[0][x][s];
#pragma unroll
		for (int x = 0; x < 16; x++) rec[16 + x] = flag ? sh.cells[1][x][s] : 0u;
		rec[37] = flag;
		return BQ_FETCH;
	}
	u32 rem = qlen - a_pos, advance = a_pos - a_lq, gap = advance - a_ll, guess = a_ls + advance;
	if (guess < c.N && gap <= c.t) {
		u32 run = guess < c.mid ? c.mid - guess : (guess == c.mid ? 0u : c.N - guess);
		sh.cs[s] = guess, sh.clim[s] = min(rem, run);
	} else {
		sh.cs[s] = 0, sh.clim[s] = 0;  // no lucky attempt: the compare only fetches the query k-mer
	}
	sh.ck[s] = 0;
	sh.flags[s] = fl & ~(BF_IS_CAND | BF_FOUND);
	return BQ_CMP;
}

BIN_FN u32 bin_run_limit(const BinConst &c, u32 p) {
	return p < c.mid ? c.mid - p : (p == c.mid ? 0u : c.N - p);
}


// One unit `s` taken off queue `q` of the current super-step: run that phase, append the unit to
// the queue of its next phase in queue set `nxt` (or retire it when the dispenser is empty).
BIN_FN void bin_phase(u32 q, u32 s, BinShared &sh, u32 nxt, const SubjectIndex &S, const QueryView *queries,
					  const u32 *query_ids, const BinConst &c, unsigned long long total, u32 *records,
					  unsigned long long *next_unit) {
	const u64 *s_code = S.rs.code;
	const u64 *q_code = S.qcode_base + sh.qoff[s];
	const u32 chunk = c.chunk, cpq = c.cpq;
	if (q == BQ_FETCH) {
		// ------------------------------------------------ FETCH
		bool got = false;
		for (;;) {
			unsigned long long unit = bin_next_unit(next_unit);
			if (unit >= total) break;
			u32 k = (u32)unit / cpq, ch = (u32)unit - k * cpq;
			u32 qid = query_ids ? query_ids[k] : k;
			u32 ql = queries[qid].t.len;
			unsigned long long start = (unsigned long long)ch * chunk;
			if (qid == S.self || start >= ql) continue;
			sh.qoff[s] = (u32)(queries[qid].t.code - S.qcode_base);
			sh.qlen[s] = ql, sh.unit[s] = (u32)unit, sh.flags[s] = 0;
			sh.c_end[s] = (u32)min((unsigned long long)ql, start + chunk);
			sh.a_pos[s] = (u32)start, sh.a_ls[s] = 0, sh.a_lq[s] = 0, sh.a_ll[s] = 0, sh.a_pm[s] = 0;
#pragma unroll
			for (int x = 0; x < 16; x++) sh.cells[0][x][s] = 0, sh.cells[1][x][s] = 0;
			got = true;
			break;
		}
		if (got) bin_push(sh, nxt, bin_begin(sh, s, c, records), s);  // else: the slot retires
	} else if (q == BQ_CMP) {
		// ------------------------------------------------ CMP: one 64-base window
		const u32 a_pos = sh.a_pos[s], ck = sh.ck[s], cs = sh.cs[s], clim = sh.clim[s];
		u32 fl = sh.flags[s];
		const bool is_cand = (fl & BF_IS_CAND) != 0;
		u64 q0, q1, s0, s1;
		window64(q_code, a_pos + ck, q0, q1);
		window64(s_code, cs + ck, s0, s1);
		if (ck == 0 && !is_cand) sh.key[s] = c.K > 0 ? kmer_key(q0, c.K) : 0u;
		u64 x0 = q0 ^ s0, x1 = q1 ^ s1;
		u32 left = clim - ck;
		u32 d = x0 ? bin_ffs64(x0) >> 1
				   : (x1 ? 32u + (bin_ffs64(x1) >> 1) : 64u);
		u32 len, mm = 0, to;
		if (d >= left) {
			len = clim;
		} else if (d < 64u) {
			len = ck + d;
			u64 sw = d < 32u ? s0 : s1, qw = d < 32u ? q0 : q1;
			u32 shf = 2u * (d & 31u);
			mm = ANDI_BIN_MM_VALID | ((((u32)(sw >> shf)) & 3u) << 2) | (((u32)(qw >> shf)) & 3u);
		} else {
			sh.ck[s] = ck + 64u;
			bin_push(sh, nxt, BQ_CMP, s);
			return;
		}
		if (!is_cand) {
			if (len >= c.t) {
				sh.best_p[s] = cs, sh.best[s] = len, sh.best_mm[s] = mm;
				fl |= BF_FOUND;
				to = BQ_DECIDE;
			} else {
				to = (c.K > 0 && sh.qlen[s] - a_pos >= (u32)c.K) ? BQ_DIR : BQ_SLOW;
			}
		} else {
			u32 best = sh.best[s], cnt = (fl >> BF_CNT_SHIFT) & 3u;
			if (len > best)
				sh.best[s] = best = len, sh.best_p[s] = cs, sh.best_mm[s] = mm, cnt = 1;
			else if (len == best)
				cnt = min(cnt + 1u, 2u);
			fl = (fl & ~(3u << BF_CNT_SHIFT)) | (cnt << BF_CNT_SHIFT);
			u32 cand = sh.key[s] + 1u;
			sh.key[s] = cand;
			if (cand < sh.hi[s]) {
				to = BQ_CAND;
			} else if (best >= (u32)c.K) {
				if (cnt == 1u && best >= c.t) fl |= BF_FOUND;
				to = BQ_DECIDE;
			} else {
				to = BQ_SLOW;  // cannot happen: the directory counts only whole k-mers
			}
		}
		sh.flags[s] = fl;
		bin_push(sh, nxt, to, s);
	} else if (q == BQ_DIR) {
		// ------------------------------------------------ DIR
		u64 fe = bin_ld64(S.fdir + sh.key[s]);
		u32 tag = ANDI_FDIR_TAG(fe), to;
		u32 fl = sh.flags[s] & ~(3u << BF_CNT_SHIFT);
		sh.best[s] = 0, sh.best_p[s] = 0, sh.best_mm[s] = 0;
		if (tag == 0u) {
			sh.best[s] = (u32)fe;  // the length of the match; no anchor
			to = BQ_DECIDE;
		} else if (tag == 1u) {
			u32 p = (u32)fe;
			sh.key[s] = 0, sh.hi[s] = 1;
			sh.cs[s] = p, sh.ck[s] = 0, sh.clim[s] = min(sh.qlen[s] - sh.a_pos[s], bin_run_limit(c, p));
			fl |= BF_IS_CAND;
			to = BQ_CMP;
		} else {
			u32 t0 = (u32)fe, cnt = (u32)(fe >> 32) & 0x3fffffffu;
			sh.key[s] = t0, sh.hi[s] = t0 + cnt;
			to = cnt <= ANDI_SCAN_MAX ? BQ_CAND : BQ_SLOW;
		}
		sh.flags[s] = fl;
		bin_push(sh, nxt, to, s);
	} else if (q == BQ_CAND) {
		// ------------------------------------------------ CAND
		u32 p = bin_ld32(S.SA + sh.key[s]);
		sh.cs[s] = p, sh.ck[s] = 0, sh.clim[s] = min(sh.qlen[s] - sh.a_pos[s], bin_run_limit(c, p));
		sh.flags[s] |= BF_IS_CAND;
		bin_push(sh, nxt, BQ_CMP, s);
	} else if (q == BQ_SLOW) {
		// ------------------------------------------------ SLOW (rare)
		u32 a_pos = sh.a_pos[s], m_len = 0, m_pos = 0;
		bool m_unique = false;
		bin_slow_lookup(S, q_code, sh.qlen[s], a_pos, m_len, m_unique, m_pos);
		bool found = m_unique && m_len >= c.t;
		sh.best[s] = m_len, sh.best_mm[s] = 0, sh.best_p[s] = found ? m_pos : 0u;
		if (found) sh.flags[s] |= BF_FOUND;
		bin_push(sh, nxt, BQ_DECIDE, s);
	} else if (q == BQ_DECIDE) {
		// ------------------------------------------------ DECIDE: src/process.c:160-196
		u32 fl = sh.flags[s];
		u32 a_pos = sh.a_pos[s];
		const u32 r_len = sh.best[s];
		bool need_cols = false;
		if (fl & BF_FOUND) {
			const u32 r_s = sh.best_p[s], r_mm = sh.best_mm[s];
			const u32 a_ls = sh.a_ls[s], a_lq = sh.a_lq[s], a_ll = sh.a_ll[s], a_pm = sh.a_pm[s];
			const u32 set = (fl & BF_PHASE2) ? 1u : 0u;
			const u32 sign = ((fl & BF_PHASE2) && !(fl & BF_A_TRUE)) ? 0xffffffffu : 1u;
			u32 *col = &sh.cells[set][0][s];
			u32 end_s = a_ls + a_ll, end_q = a_lq + a_ll;
			bool pairs = r_s > end_s && (a_pos - end_q) == (r_s - end_s) && ((r_s < c.border) == (a_ls < c.border));
			if (pairs || (a_pm & 1u) || a_ll >= 2u * c.t) {	 // src/model.c:247-254
				u32 f = (a_ll >> 2) * sign;
				col[0 * ANDI_BIN_SLOTS] += f;
				col[5 * ANDI_BIN_SLOTS] += f;
				col[10 * ANDI_BIN_SLOTS] += f;
				col[15 * ANDI_BIN_SLOTS] += f + (a_ll & 3u) * sign;
			}
			if (pairs) {
				u32 g = a_pos - end_q, lm = a_pm >> 1;
				if (g == 1u && (lm & ANDI_BIN_MM_VALID)) {
					col[(lm & 15u) * ANDI_BIN_SLOTS] += sign;
				} else {
					sh.cs[s] = end_s, sh.ck[s] = g;	 // cols_s, cols_left; cols_q follows from the diagonal
					need_cols = true;
				}
			}
			sh.a_ls[s] = r_s, sh.a_lq[s] = a_pos, sh.a_ll[s] = r_len;
			sh.a_pm[s] = (pairs ? 1u : 0u) | (r_mm << 1);
		}
		sh.a_pos[s] = a_pos + r_len + 1u;
		bin_push(sh, nxt, need_cols ? (u32)BQ_COLS : bin_begin(sh, s, c, records), s);
	} else {
		// ------------------------------------------------ COLS: src/model.c:309-337
		u32 fl = sh.flags[s];
		u32 cols_s = sh.cs[s], cols_left = sh.ck[s];
		// the gap lies on the diagonal of the anchor just stored: q = s + (a_lq - a_ls)
		u32 cols_q = cols_s + (sh.a_lq[s] - sh.a_ls[s]);
		u32 span = min(16u, cols_left);
		u32 qw = window16(q_code, cols_q), sw = window16(s_code, cols_s);
		u32 valid = span == 16u ? 0x55555555u : (0x55555555u & ((1u << (2u * span)) - 1u));
		if (cols_s <= c.mid && c.mid - cols_s < span) valid &= ~(1u << (2u * (c.mid - cols_s)));  // '#' column
		u32 x = qw ^ sw;
		u32 neq = (x | (x >> 1)) & valid, eq = valid & ~neq;
		u32 lo = qw & 0x55555555u, hb = (qw >> 1) & 0x55555555u;
		const u32 set = (fl & BF_PHASE2) ? 1u : 0u;
		const u32 sign = ((fl & BF_PHASE2) && !(fl & BF_A_TRUE)) ? 0xffffffffu : 1u;
		u32 *col = &sh.cells[set][0][s];
		col[0 * ANDI_BIN_SLOTS] += bin_popc32(eq & ~hb & ~lo) * sign;
		col[5 * ANDI_BIN_SLOTS] += bin_popc32(eq & ~hb & lo) * sign;
		col[10 * ANDI_BIN_SLOTS] += bin_popc32(eq & hb & ~lo) * sign;
		col[15 * ANDI_BIN_SLOTS] += bin_popc32(eq & hb & lo) * sign;
		while (neq) {
			u32 d2 = bin_ffs32(neq);
			neq &= neq - 1;
			col[((((sw >> d2) & 3u) << 2) | ((qw >> d2) & 3u)) * ANDI_BIN_SLOTS] += sign;
		}
		cols_left -= span;
		if (cols_left) {
			sh.cs[s] = cols_s + span, sh.ck[s] = cols_left;
			bin_push(sh, nxt, BQ_COLS, s);
		} else {
			bin_push(sh, nxt, bin_begin(sh, s, c, records), s);
		}
	}

}
