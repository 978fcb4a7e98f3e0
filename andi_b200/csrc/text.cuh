// andi_b200/csrc/text.cuh -- packed-text device helpers shared by every kernel.
//
// Text layout in HBM (both for pool sequences and for RS = revcomp '#' forward):
//   code plane  u64 words, 32 characters per word, 2 bits each, character c of a word at
//               bits [2*(c&31), 2*(c&31)+2).  A=0 C=1 G=2 T=3 (nucl2bit, src/model.c:295-299).
//   spec plane  same geometry; 01 where the character is not a nucleotide. For those the code
//               plane holds 0 for '!', 1 for '#', 2 for ';'.
// Byte equality of the reference (`S[a] == Q[b]`, src/process.c:59-65, src/esa.c:408,546,592)
// is therefore "code pair equal and spec pair equal", and the byte order the suffix array
// is built under ('\0' < '!' < '#' < ';' < A < C < G < T) is sym3() below.
// Every plane is 16-byte aligned and allocated with zero guard words (plane_words()) so a
// 32-character window may start at any position up to and including the text length and a
// 64-character window (two aligned word pairs) at any position below it.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef uint32_t u32;

#define ANDI_EVEN_BITS 0x5555555555555555ULL

struct TextView {
	const u64 *code;
	const u64 *spec;
	u32 len;  // characters
	u32 mid;  // RS only: index of '#' (== forward length n); 0xffffffff for plain sequences
};

// 32 characters starting at pos (character 0 in the low bits): bits [2*(pos&31), +64) of the
// 128-bit pair (w[i+1] : w[i]), taken with two 32-bit funnel shifts. The second word is always
// loaded (the guard words make that safe); that is cheaper than branching on an aligned pos.
__device__ __forceinline__ u64 window32(const u64 *__restrict__ w, u32 pos) {
	u32 i = pos >> 5, sh = (pos & 31u) * 2u;
	u64 a = __ldg(w + i), b = __ldg(w + i + 1);
	u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
	bool upper = sh >= 32u;
	u32 x0 = upper ? a1 : a0, x1 = upper ? b0 : a1, x2 = upper ? b1 : b0;
	u32 s = sh & 31u;
	u32 lo = __funnelshift_r(x0, x1, s), hi = __funnelshift_r(x1, x2, s);
	return ((u64)hi << 32) | lo;
}

// 64 characters starting at pos: lo = characters 0..31, hi = 32..63. The three words i..i+2
// that hold them are fetched as the two ALIGNED 16-byte pairs that cover them (planes are 16-byte
// aligned, plane_words() leaves room for the last pair): two 128-bit requests instead of three
// 64-bit ones -- the walk is bound by L1/TEX request throughput as much as by issue slots.
// The caller guarantees pos < text length.
__device__ __forceinline__ void window64(const u64 *__restrict__ w, u32 pos, u64 &lo, u64 &hi) {
	u32 i = pos >> 5, sh = (pos & 31u) * 2u;
	const uint4 *pair = reinterpret_cast<const uint4 *>(w) + (i >> 1);
	uint4 a = __ldg(pair), b = __ldg(pair + 1);
	// eight 32-bit pieces a.x a.y a.z a.w b.x b.y b.z b.w; the window starts at piece 2*(i&1) + (sh>=32)
	bool odd = i & 1u, upper = sh >= 32u;
	u32 y0 = odd ? a.z : a.x, y1 = odd ? a.w : a.y, y2 = odd ? b.x : a.z, y3 = odd ? b.y : a.w, y4 = odd ? b.z : b.x,
		y5 = odd ? b.w : b.y;
	u32 x0 = upper ? y1 : y0, x1 = upper ? y2 : y1, x2 = upper ? y3 : y2, x3 = upper ? y4 : y3, x4 = upper ? y5 : y4;
	u32 s = sh & 31u;
	lo = ((u64)__funnelshift_r(x1, x2, s) << 32) | __funnelshift_r(x0, x1, s);
	hi = ((u64)__funnelshift_r(x3, x4, s) << 32) | __funnelshift_r(x2, x3, s);
}

// 16 characters starting at pos, from the 32-bit halves of the plane (little endian: the low
// half of a word holds its characters 0..15). pos < text length.
__device__ __forceinline__ u32 window16(const u64 *__restrict__ w, u32 pos) {
	const u32 *h = reinterpret_cast<const u32 *>(w) + (pos >> 4);
	return __funnelshift_r(__ldg(h), __ldg(h + 1), (pos & 15u) * 2u);
}

__device__ __forceinline__ u32 code_at(const u64 *__restrict__ w, u32 pos) {
	return (u32)(__ldg(w + (pos >> 5)) >> ((pos & 31u) * 2u)) & 3u;
}

// Length of the common prefix of a[pa..] and b[pb..], at most `limit` characters.
// SPEC=false ignores the spec planes: the caller guarantees that no separator lies inside
// either range (it cuts `limit` at '#', and neither text contains '!' / ';').
template <bool SPEC>
__device__ __forceinline__ u32 match_len(const TextView &a, u32 pa, const TextView &b, u32 pb,
										 u32 limit) {
	u32 k = 0;
	while (k < limit) {
		u64 x = window32(a.code, pa + k) ^ window32(b.code, pb + k);
		if (SPEC) x |= window32(a.spec, pa + k) ^ window32(b.spec, pb + k);
		if (x) {  // lowest set bit sits in the 2-bit pair of the first differing character
			k += (u32)(__ffsll((long long)x) - 1) >> 1;
			return k < limit ? k : limit;
		}
		k += 32;
	}
	return limit;
}

// Rank of the character at pos in the reference's byte order; 0 past the end.
// For RS in SPEC=false mode the only separator is '#' at t.mid.
template <bool SPEC>
__device__ __forceinline__ u32 sym3(const TextView &t, u32 pos) {
	if (pos >= t.len) return 0;
	u32 c = code_at(t.code, pos);
	if (SPEC) {
		u32 s = code_at(t.spec, pos) & 1u;
		return s ? c + 1 : c + 4;
	}
	return pos == t.mid ? 2u : c + 4;
}

// How far a comparison that starts at RS position p may run before it hits '#' or the end.
// SPEC=true lets the planes decide, so only the end counts.
template <bool SPEC>
__device__ __forceinline__ u32 rs_run(const TextView &rs, u32 p) {
	if (SPEC) return rs.len - p;
	if (p < rs.mid) return rs.mid - p;
	if (p == rs.mid) return 0;
	return rs.len - p;
}

// Longest possible common prefix of two different suffixes a, b of RS when the planes are
// not consulted for '#': the shorter suffix ends, or '#' faces another character.
__device__ __forceinline__ u32 pair_limit_fast(const TextView &rs, u32 a, u32 b) {
	u32 lo = a < b ? a : b, hi = a < b ? b : a;
	u32 lim = rs.len - hi;
	if (hi <= rs.mid)
		lim = min(lim, rs.mid - hi);
	else if (lo <= rs.mid)
		lim = min(lim, rs.mid - lo);
	return lim;
}

// First k characters (k <= 16) of a window as an integer with the FIRST character most
// significant, i.e. monotone in lexicographic order. Used as k-mer directory key.
__device__ __forceinline__ u32 kmer_key(u64 win, int k) {
	u32 w = (u32)win;							  // 16 characters
	w = __brev(w);								  // character order reversed, bits inside pairs swapped
	w = ((w >> 1) & 0x55555555u) | ((w & 0x55555555u) << 1);
	return k >= 16 ? w : (w >> (32 - 2 * k));
}
