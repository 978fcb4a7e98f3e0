// andi_b200/csrc/esa_kernels.cuh -- index construction kernels.
#pragma once
#include "text.cuh"

// ------------------------------------------------------------------ P0: pack + GC
// src/sequence.c:196-207 (calc_gc) and the 2-bit coding of src/model.c:295-299. One thread
// packs 32 input bytes into one code word and one spec word. counters[0] += #G + #C,
// counters[1] += #separators. Anything that is not A/C/G/T is stored as the separator '!'
// (inputs are normalized, src/sequence.c:260-282, so this is the only other byte).
__global__ void k_pack(const unsigned char *__restrict__ chars, u32 n, u64 *__restrict__ code,
					   u64 *__restrict__ spec, u32 nwords, unsigned long long *counters) {
	u32 w = blockIdx.x * blockDim.x + threadIdx.x;
	u32 gc = 0, sep = 0;
	if (w < nwords) {
		u64 cw = 0, sw = 0;
		u32 base = w * 32u;
		if (base < n) {
			u32 cnt = n - base < 32u ? n - base : 32u;
			for (u32 d = 0; d < cnt; d++) {
				u32 c = chars[base + d];
				bool nuc = (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T');
				u32 v = c & 6u;
				v ^= v >> 1;
				v >>= 1;
				if (!nuc) v = 0;
				cw |= (u64)v << (2 * d);
				sw |= (u64)(!nuc) << (2 * d);
				gc += nuc & ((v == 1) | (v == 2));
				sep += !nuc;
			}
		}
		code[w] = cw;
		spec[w] = sw;
	}
	// block reduction, one atomic per block and counter
	__shared__ u32 s_gc, s_sep;
	if (threadIdx.x == 0) s_gc = 0, s_sep = 0;
	__syncthreads();
	for (int o = 16; o; o >>= 1) {
		gc += __shfl_down_sync(0xffffffffu, gc, o);
		sep += __shfl_down_sync(0xffffffffu, sep, o);
	}
	if ((threadIdx.x & 31) == 0) {
		if (gc) atomicAdd(&s_gc, gc);
		if (sep) atomicAdd(&s_sep, sep);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		if (s_gc) atomicAdd(&counters[0], (unsigned long long)s_gc);
		if (s_sep) atomicAdd(&counters[1], (unsigned long long)s_sep);
	}
}

// Pack an RS string given as bytes (andi_esa_build_rs): same coding, '#' -> (spec, 1),
// ';' -> (spec, 2), '!' -> (spec, 0). counters[0] += number of '#', counters[1] += other separators.
__global__ void k_pack_rs_bytes(const unsigned char *__restrict__ chars, u32 n, u64 *__restrict__ code,
								u64 *__restrict__ spec, u32 nwords, unsigned long long *counters) {
	u32 w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= nwords) return;
	u64 cw = 0, sw = 0;
	u32 base = w * 32u, sep = 0, hashes = 0;
	if (base < n) {
		u32 cnt = n - base < 32u ? n - base : 32u;
		for (u32 d = 0; d < cnt; d++) {
			u32 c = chars[base + d];
			bool nuc = (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T');
			u32 v = c & 6u;
			v ^= v >> 1;
			v >>= 1;
			if (!nuc) {
				v = c == '#' ? 1u : c == ';' ? 2u : 0u;
				sep += (c != '#');
				hashes += (c == '#');
			}
			cw |= (u64)v << (2 * d);
			sw |= (u64)(!nuc) << (2 * d);
		}
	}
	code[w] = cw;
	spec[w] = sw;
	if (sep) atomicAdd(&counters[1], (unsigned long long)sep);
	if (hashes) atomicAdd(&counters[0], (unsigned long long)hashes);  // counters[0] = number of '#'
}

// src/sequence.c:143-189 (revcomp + catcomp) on packed planes:
// RS[i] = complement(fwd[n-1-i]) for i < n, '#' at n, fwd[i-n-1] above. The complement of a
// nucleotide code is code ^ 3; a separator ('!') becomes ';' in the reverse half
// (src/sequence.c:157-158, pinned by test/test_seq.c:69).
__global__ void k_build_rs(const u64 *__restrict__ fcode, const u64 *__restrict__ fspec, u32 n,
						   u64 *__restrict__ rcode, u64 *__restrict__ rspec, u32 nwords) {
	u32 w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= nwords) return;
	u32 N = 2 * n + 1;
	u64 cw = 0, sw = 0;
	u32 base = w * 32u;
	for (u32 d = 0; d < 32; d++) {
		u32 i = base + d;
		if (i >= N) break;
		u32 c, s;
		if (i < n) {
			u32 src = n - 1 - i;
			s = code_at(fspec, src) & 1u;
			c = s ? 2u : (code_at(fcode, src) ^ 3u);
		} else if (i == n) {
			s = 1, c = 1;
		} else {
			u32 src = i - n - 1;
			s = code_at(fspec, src) & 1u;
			c = s ? 0u : code_at(fcode, src);
		}
		cw |= (u64)c << (2 * d);
		sw |= (u64)s << (2 * d);
	}
	rcode[w] = cw;
	rspec[w] = sw;
}

// ------------------------------------------------------------------ E1: suffix array
// Stands where src/esa.c:303 calls divsufsort. Radix-sorted prefix doubling:
// round 0 orders all suffixes by their first 16 characters (3 bits per character in the
// reference's byte order, 48-bit keys), later rounds order the still-ambiguous suffixes by
// (group, rank[i + h]) with h = 16, 32, ...

__global__ void k_suffix_keys(TextView rs, u64 *__restrict__ keys, u32 *__restrict__ idx) {
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= rs.len) return;
	u64 cw = window32(rs.code, i), sw = window32(rs.spec, i);
	u64 key = 0;
#pragma unroll
	for (u32 d = 0; d < 16; d++) {
		u32 c = (u32)(cw >> (2 * d)) & 3u, s = (u32)(sw >> (2 * d)) & 1u;
		u32 sym = (i + d < rs.len) ? (s ? c + 1 : c + 4) : 0u;
		key = (key << 3) | sym;
	}
	keys[i] = key;
	idx[i] = i;
}

// v[j] = j where a new group starts, else 0 (max-scan turns this into group heads).
__global__ void k_head_values(const u64 *__restrict__ keys, u32 m, const u32 *__restrict__ pos,
							  u32 *__restrict__ v) {
	u32 k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= m) return;
	bool head = k == 0 || keys[k] != keys[k - 1];
	u32 where = pos ? pos[k] : k;
	v[k] = head ? where : 0u;
}

// After the scan: grp_sorted[k] is the SA index of the head of k's group.
// Writes rank[suffix] and the "still ambiguous" flag.
__global__ void k_apply_groups(const u32 *__restrict__ grp_sorted, u32 m, const u32 *__restrict__ pos,
							   const u32 *__restrict__ suffix, u32 *__restrict__ SA,
							   u32 *__restrict__ grp, u32 *__restrict__ rank,
							   unsigned char *__restrict__ ambiguous) {
	u32 k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= m) return;
	u32 where = pos ? pos[k] : k;
	u32 g = grp_sorted[k];
	u32 sfx = suffix[k];
	if (pos) SA[where] = sfx;  // round 0 sorts straight into SA
	grp[where] = g;
	rank[sfx] = g;
	bool head = g == where;
	bool next_head = (k + 1 == m) || (grp_sorted[k + 1] == (pos ? pos[k + 1] : k + 1));
	ambiguous[k] = !(head && next_head);
}

__global__ void k_round_keys(const u32 *__restrict__ pos, u32 m, const u32 *__restrict__ SA,
							 const u32 *__restrict__ grp, const u32 *__restrict__ rank, u32 h, u32 N,
							 u64 *__restrict__ keys, u32 *__restrict__ vals) {
	u32 k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= m) return;
	u32 j = pos[k];
	u32 sfx = SA[j];
	u32 second = (sfx + h < N) ? rank[sfx + h] + 1u : 0u;
	keys[k] = ((u64)grp[j] << 32) | second;
	vals[k] = sfx;
}

// ------------------------------------------------------------------ E2: LCP
// src/esa.c:373-426. phi[SA[j]] = SA[j-1]; PLCP in text order with the l-1 carry-over, each
// thread owning a slice of 32 consecutive text positions (the carry restarts at 0 at a slice
// start, which only costs comparisons, never correctness); LCP[j] = PLCP[SA[j]].

__global__ void k_phi(const u32 *__restrict__ SA, u32 N, int32_t *__restrict__ phi) {
	u32 j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= N) return;
	phi[SA[j]] = j ? (int32_t)SA[j - 1] : -1;
}

template <bool SPEC>
__global__ void k_plcp(TextView rs, int32_t *__restrict__ phi_plcp) {
	u32 t = blockIdx.x * blockDim.x + threadIdx.x;
	u32 begin = t * 32u;
	if (begin >= rs.len) return;
	u32 end = min(begin + 32u, rs.len);
	u32 l = 0;
	for (u32 i = begin; i < end; i++) {
		int32_t k = phi_plcp[i];
		if (k < 0) {
			phi_plcp[i] = -1;
			continue;
		}
		u32 lim = SPEC ? rs.len - max(i, (u32)k) : pair_limit_fast(rs, i, (u32)k);
		if (l < lim) l += match_len<SPEC>(rs, i + l, rs, (u32)k + l, lim - l);
		phi_plcp[i] = (int32_t)l;
		l = l ? l - 1 : 0;
	}
}

__global__ void k_lcp_from_plcp(const u32 *__restrict__ SA, const int32_t *__restrict__ plcp, u32 N,
								int32_t *__restrict__ LCP) {
	u32 j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j > N) return;
	LCP[j] = (j == 0 || j == N) ? -1 : plcp[SA[j]];
}

// ------------------------------------------------------------------ k-mer directory
// Not in the reference: the device-side replacement for its 4^10 prefix cache + child-table
// descent. dir[x] (written by k_bucket_sort, sa_bucket.cuh) = first SA index of the suffixes
// that start with the k-mer x (k nucleotides, no separator, inside the text) | their number
// << 32; they are contiguous in SA.
//
// Presence bitmaps: level m (1 <= m < K) has bit x set iff the m-mer x occurs in RS. Level K-1
// is left by k_bucket_sort (a folded warp ballot of its counts), lower levels come by OR-ing groups of four bits, and suffixes
// that hit a separator / the end before K characters are patched in. They only serve to
// build the prefix-length table (k_prefix_len) and are freed afterwards.
__global__ void k_presence_down(const u32 *__restrict__ upper, u32 nbits, u32 *__restrict__ lower) {
	u32 w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w * 32u >= nbits) return;
	u32 out = 0;
	for (u32 b = 0; b < 32; b++) {
		u32 x = w * 32u + b;
		if (x >= nbits) break;
		u32 up = 4u * x;  // bits up..up+3 of the upper level (never straddle a word)
		if ((upper[up >> 5] >> (up & 31u)) & 0xfu) out |= 1u << b;
	}
	lower[w] = out;
}

struct PresenceLevels {
	u32 *bits;		 // all levels in one buffer
	u32 offset[16];	 // word offset of level m
};

// All levels below `top` from level `top` in ONE launch of a single CTA (each level is a quarter of
// the one above: from level K-2 down they are small enough for one CTA, and ten launches of a
// microsecond each cost more than the work).
__global__ void __launch_bounds__(1024) k_presence_down_all(PresenceLevels lv, int top) {
	for (int m = top - 1; m >= 1; m--) {
		const u32 nbits = 1u << (2 * m);
		const u32 *upper = lv.bits + lv.offset[m + 1];
		u32 *lower = lv.bits + lv.offset[m];
		for (u32 w = threadIdx.x; w * 32u < nbits; w += blockDim.x) {
			u32 out = 0;
			for (u32 b = 0; b < 32; b++) {
				u32 x = w * 32u + b;
				if (x >= nbits) break;
				u32 up = 4u * x;
				if ((upper[up >> 5] >> (up & 31u)) & 0xfu) out |= 1u << b;
			}
			lower[w] = out;
		}
		__syncthreads();  // the next level reads what this one wrote (same CTA: visible after the barrier)
	}
}

// Positions [first, first + count) are examined (all of the text when separators may be anywhere,
// just the K positions in front of '#' and of the text end otherwise).
__global__ void k_presence_patch(TextView rs, int K, PresenceLevels lv, u32 first, u32 count) {
	u32 p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= count) return;
	p += first;
	if (p >= rs.len) return;
	// run = number of nucleotides starting at p before a separator / the end (capped at K)
	u64 sw = window32(rs.spec, p);
	u32 run = sw ? (u32)(__ffsll((long long)sw) - 1) >> 1 : 32u;
	run = min(run, rs.len - p);
	if (run >= (u32)K) return;	// covered by the directory
	u64 cw = window32(rs.code, p);
	for (u32 m = 1; m <= run && m < (u32)K; m++) {
		u32 x = kmer_key(cw, (int)m);
		atomicOr(&lv.bits[lv.offset[m] + (x >> 5)], 1u << (x & 31u));
	}
}

// plen[x] for every k-mer x: length of the longest prefix of x (0 .. K-1) that occurs in RS --
// what a lookup of an ABSENT k-mer returns. It does not depend on the last character, so one
// thread serves the four k-mers 4y..4y+3 of a (K-1)-mer y: one probe sequence, one 32-byte
// directory read, and the four entries of the walk's own directory view (fdir, see
// sa_bucket.cuh: tag 0 + plen / tag 1 + text position of the only suffix and the 15 bases that follow the k-mer
// there / tag 2 + the text positions of
// both suffixes / tag 3 + first index and count) in the same pass. Runs once the suffix array is final.
__global__ void k_prefix_len(PresenceLevels lv, int K, const u64 *__restrict__ dir, const u32 *__restrict__ SA,
							 const u64 *__restrict__ code, unsigned char *__restrict__ plen, u64 *__restrict__ fdir) {
	u32 y = blockIdx.x * blockDim.x + threadIdx.x;
	if (y >= (1u << (2 * (K - 1)))) return;
	u32 l = 0;
	for (int m = K - 1; m >= 1; m--) {
		u32 z = y >> (2 * (K - 1 - m));
		if ((lv.bits[lv.offset[m] + (z >> 5)] >> (z & 31u)) & 1u) {
			l = (u32)m;
			break;
		}
	}
	reinterpret_cast<u32 *>(plen)[y] = l * 0x01010101u;
	const ulonglong2 *din = reinterpret_cast<const ulonglong2 *>(dir) + 2 * (size_t)y;
	ulonglong2 d01 = din[0], d23 = din[1];
	u64 de[4] = {d01.x, d01.y, d23.x, d23.y}, out[4];
#pragma unroll
	for (int c = 0; c < 4; c++) {
		u32 first = (u32)de[c], count = (u32)(de[c] >> 32);
		if (count == 0)
			out[c] = l;
		else if (count == 1) {
			// the only suffix: its position and the 15 bases behind the k-mer there (whatever the
			// plane holds: the walk cuts every compare at '#' / the text end by position)
			const u32 p = SA[first];
			out[c] = (1ULL << 62) | ((window32(code, p + (u32)K) & 0x3fffffffULL) << 31) | p;
		}
		else if (count == 2)
			out[c] = (2ULL << 62) | ((u64)SA[first + 1] << 31) | SA[first];
		else
			out[c] = (3ULL << 62) | ((u64)count << 32) | first;
	}
	ulonglong2 *dout = reinterpret_cast<ulonglong2 *>(fdir) + 2 * (size_t)y;
	dout[0] = make_ulonglong2(out[0], out[1]);
	dout[1] = make_ulonglong2(out[2], out[3]);
}

// ------------------------------------------------------------------ E3: FVC
// src/esa.c:229-245: FVC[i] = S[SA[i] + LCP[i]] as the original byte.
__device__ __forceinline__ char byte_at(const TextView &rs, u32 pos) {
	if (pos >= rs.len) return '\0';
	u32 c = code_at(rs.code, pos), s = code_at(rs.spec, pos) & 1u;
	return s ? (c == 0 ? '!' : c == 1 ? '#' : ';') : (c == 0 ? 'A' : c == 1 ? 'C' : c == 2 ? 'G' : 'T');
}

__global__ void k_fvc(TextView rs, const u32 *__restrict__ SA, const int32_t *__restrict__ LCP,
					  char *__restrict__ FVC) {
	u32 j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= rs.len) return;
	FVC[j] = byte_at(rs, (u32)((int32_t)SA[j] + LCP[j]));
}

// ------------------------------------------------------------------ E4: CLD
// src/esa.c:312-363 through its nearest-smaller-value form:
// for x in [1, N-1]: y = nearest index < x with LCP[y] <= LCP[x], k = nearest index > x with
// LCP[k] < LCP[x]; LCP[y]==LCP[x] or LCP[k] < LCP[y] -> CLD[y] = x, else CLD[k-1] = x.
// Nearest smaller values are found with a min-pyramid over LCP (fan-out 32 per level).

struct MinPyramid {
	const int32_t *level[8];  // level[0] = LCP itself (N+1 entries)
	u32 size[8];
	int levels;
};

__global__ void k_min_reduce32(const int32_t *__restrict__ in, u32 n_in, int32_t *__restrict__ out,
							   u32 n_out) {
	u32 o = blockIdx.x * blockDim.x + threadIdx.x;
	if (o >= n_out) return;
	u32 b = o * 32u, e = min(b + 32u, n_in);
	int32_t m = in[b];
	for (u32 i = b + 1; i < e; i++) m = min(m, in[i]);
	out[o] = m;
}

// nearest index < x whose value is <= v (exists: LCP[0] = -1)
__device__ u32 prev_le(const MinPyramid &P, u32 x, int32_t v) {
	// scan inside x's own block of 32 at level 0, then climb
	u32 i = x;
	int lvl = 0;
	for (;;) {
		// scan leftwards inside the current block at this level
		u32 block_start = i & ~31u;
		while (i > block_start) {
			i--;
			if (P.level[lvl][i] <= v) goto descend;
		}
		// nothing in this block left of i: move to the parent level, position = this block
		i = block_start >> 5;
		lvl++;
	}
descend:
	while (lvl > 0) {
		// block i at level lvl contains a value <= v: take the rightmost child that does
		lvl--;
		u32 b = i << 5, e = min(b + 32u, P.size[lvl]);
		u32 c = e;
		while (c > b) {
			c--;
			if (P.level[lvl][c] <= v) break;
		}
		i = c;
	}
	return i;
}

// nearest index > x whose value is < v (exists: LCP[N] = -1)
__device__ u32 next_lt(const MinPyramid &P, u32 x, int32_t v) {
	u32 i = x;
	int lvl = 0;
	for (;;) {
		u32 block_end = min((i | 31u) + 1u, P.size[lvl]);
		while (i + 1 < block_end) {
			i++;
			if (P.level[lvl][i] < v) goto descend;
		}
		i = i >> 5;
		lvl++;
	}
descend:
	while (lvl > 0) {
		lvl--;
		u32 b = i << 5, e = min(b + 32u, P.size[lvl]);
		u32 c = b;
		while (c < e && !(P.level[lvl][c] < v)) c++;
		i = c;
	}
	return i;
}

__global__ void k_cld(MinPyramid P, u32 N, int32_t *__restrict__ CLD) {
	u32 x = blockIdx.x * blockDim.x + threadIdx.x;
	if (x == 0) {
		CLD[0] = (int32_t)N;
		CLD[N] = 0;	 // never written by the reference
		return;
	}
	if (x >= N) return;
	const int32_t *LCP = P.level[0];
	int32_t v = LCP[x];
	u32 y = prev_le(P, x, v);
	u32 k = next_lt(P, x, v);
	if (LCP[y] == v || LCP[k] < LCP[y])
		CLD[y] = (int32_t)x;
	else
		CLD[k - 1] = (int32_t)x;
}

// ------------------------------------------------------------------ E5: prefix cache
// src/esa.c:73-215 evaluated independently for each of the 4^10 slots: one thread follows the
// reference's depth-first filling order along its own 10-mer (child-table steps of
// src/esa.c:441-511) and stores the interval that ends up in its slot.

struct EsaView {
	TextView rs;
	const u32 *SA;
	const int32_t *LCP;
	const int32_t *CLD;
	const char *FVC;
};

struct Inter {
	int32_t l, i, j, m;
};

__device__ __forceinline__ bool inter_empty(const Inter &v) { return v.i == -1 && v.j == -1; }

__device__ Inter child_interval(const EsaView &E, Inter p, char a) {
	Inter none = p;
	none.i = none.j = -1;
	if (p.i == p.j) return byte_at(E.rs, E.SA[p.i] + (u32)p.l) == a ? p : none;
	int32_t start = p.i, split = p.m, depth = p.l;
	char c = byte_at(E.rs, E.SA[start] + (u32)depth);
	for (;;) {
		if (c == a) {
			Inter r;
			r.i = start;
			if (start == split - 1) {
				r.j = start, r.m = -1, r.l = E.LCP[start];
			} else {
				r.j = split - 1, r.m = E.CLD[split - 1], r.l = E.LCP[r.m];
			}
			return r;
		}
		if (c > a) return none;
		start = split;
		if (start != p.j) {
			split = E.CLD[split];
			if (E.LCP[split] == depth) {
				c = E.FVC[start];
				continue;
			}
		}
		if (E.FVC[start] != a) return none;
		Inter r = {E.LCP[split], start, p.j, split};
		return r;
	}
}

__global__ void k_prefix_cache(EsaView E, Inter *__restrict__ cache) {
	const int CL = 10;
	u32 slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= (1u << (2 * CL))) return;
	int32_t N = (int32_t)E.rs.len;
	int32_t m0 = E.CLD[N - 1];
	Inter at = {E.LCP[m0], 0, N - 1, m0};
	int pos = 0;
	while (pos < CL) {
		if (inter_empty(at)) break;
		char a = "ACGT"[(slot >> (2 * (CL - 1 - pos))) & 3u];
		Inter sub = child_interval(E, at, a);
		if (inter_empty(sub)) break;
		if (sub.i == sub.j) {
			sub.l = pos + 1;
			at = sub;
			break;
		}
		if (sub.l <= pos + 1) {
			at = sub;
			pos++;
			continue;
		}
		if (sub.l >= CL) break;
		int k = pos + 1;
		bool special = false, differs = false;
		for (; k < sub.l; k++) {
			char c = byte_at(E.rs, E.SA[sub.i] + (u32)k);
			if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T')) {
				special = true;
				break;
			}
			if ("ACGT"[(slot >> (2 * (CL - 1 - k))) & 3u] != c) {
				differs = true;
				break;
			}
		}
		if (differs) break;
		at = sub;
		if (special) break;
		pos = k;
	}
	cache[slot] = at;
}
