/* andi_b200/csrc/host_pack.c -- 2-bit packing of the sequence pool ON THE HOST, before the upload.
 *
 * andi_pool_set_host used to copy the normalized characters (one byte per base) over PCIe and pack
 * them on the GPU. At 3085 x 2.1 Mbp that is 6.5 GB over a ~25 GB/s link every time a pool is set:
 * a third of bench.py's end-to-end step. Packed on the host the same pool is 1.6 GB on the wire.
 * This file is the host side of that: 32 characters -> one code word + one spec word of the layout
 * in text.cuh (A0 C1 G2 T3 = nucl2bit of src/model.c:295-299, base d of a word at bits 2d; anything
 * that is not A/C/G/T is the contig separator '!' of src/sequence.c:78-125: code 0, spec 01), the
 * G+C count of src/sequence.c:196-207 and the separator count on the way. AVX2 where the CPU has it
 * (checked at run time), plain C otherwise; OpenMP over sequences. Plain C, no CUDA.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#define EVEN 0x5555555555555555ULL

static inline uint64_t spread32(uint32_t m) { /* bit d -> bit 2d */
	uint64_t x = m;
	x = (x | (x << 16)) & 0x0000FFFF0000FFFFULL;
	x = (x | (x << 8)) & 0x00FF00FF00FF00FFULL;
	x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0FULL;
	x = (x | (x << 2)) & 0x3333333333333333ULL;
	x = (x | (x << 1)) & EVEN;
	return x;
}

static void pack_scalar(const unsigned char *src, size_t n, uint64_t *code, uint64_t *spec, size_t w0, size_t w1, uint64_t *gc,
						uint64_t *sep) {
	uint64_t g = 0, s = 0;
	for (size_t w = w0; w < w1; w++) {
		uint64_t cw = 0, sw = 0;
		const size_t base = w * 32;
		for (size_t d = 0; d < 32 && base + d < n; d++) {
			const unsigned c = src[base + d];
			const int nuc = c == 'A' || c == 'C' || c == 'G' || c == 'T';
			const uint64_t v = nuc ? (((c >> 1) & 3u) ^ ((c >> 2) & 1u)) : 0u;
			cw |= v << (2 * d);
			sw |= (uint64_t)(!nuc) << (2 * d);
		}
		code[w] = cw, spec[w] = sw;
		const uint64_t lo = cw & EVEN, hi = (cw >> 1) & EVEN;
		g += (uint64_t)__builtin_popcountll(lo ^ hi); /* C = 01, G = 10 */
		s += (uint64_t)__builtin_popcountll(sw);
	}
	*gc += g, *sep += s;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void pack_avx2(const unsigned char *src, size_t n, uint64_t *code, uint64_t *spec, size_t w0,
													   size_t w1, uint64_t *gc, uint64_t *sep) {
	const size_t full = n / 32; /* words made of 32 real characters */
	const size_t wv = w1 < full ? w1 : full;
	uint64_t g = 0, s = 0;
	const __m256i A = _mm256_set1_epi8('A'), C = _mm256_set1_epi8('C'), G = _mm256_set1_epi8('G'), T = _mm256_set1_epi8('T');
	const __m256i three = _mm256_set1_epi8(3), one = _mm256_set1_epi8(1);
	const __m256i w14 = _mm256_set1_epi16(0x0401);		/* bytes (1, 4): c0 + 4 c1 */
	const __m256i w116 = _mm256_set1_epi32(0x00100001); /* words (1, 16): + 16 (c2 + 4 c3) */
	const __m256i pick = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 4, 8, 12, -1, -1, -1, -1, -1,
										  -1, -1, -1, -1, -1, -1, -1);
	size_t w = w0;
	for (; w < wv; w++) {
		const __m256i c = _mm256_loadu_si256((const __m256i *)(src + w * 32));
		const __m256i valid = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(c, A), _mm256_cmpeq_epi8(c, C)),
											  _mm256_or_si256(_mm256_cmpeq_epi8(c, G), _mm256_cmpeq_epi8(c, T)));
		const uint32_t bad = ~(uint32_t)_mm256_movemask_epi8(valid);
		/* code = ((c >> 1) & 3) ^ ((c >> 2) & 1), zero for a separator */
		__m256i v = _mm256_xor_si256(_mm256_and_si256(_mm256_srli_epi16(c, 1), three), _mm256_and_si256(_mm256_srli_epi16(c, 2), one));
		v = _mm256_and_si256(v, valid);
		/* four 2-bit codes per byte: pairs, then pairs of pairs, then one byte out of every dword */
		v = _mm256_madd_epi16(_mm256_maddubs_epi16(v, w14), w116);
		v = _mm256_shuffle_epi8(v, pick);
		const uint64_t cw = (uint64_t)(uint32_t)_mm256_extract_epi32(v, 0) | ((uint64_t)(uint32_t)_mm256_extract_epi32(v, 4) << 32);
		const uint64_t sw = bad ? spread32(bad) : 0;
		code[w] = cw, spec[w] = sw;
		const uint64_t lo = cw & EVEN, hi = (cw >> 1) & EVEN;
		g += (uint64_t)__builtin_popcountll(lo ^ hi);
		s += (uint64_t)__builtin_popcount(bad);
	}
	*gc += g, *sep += s;
	if (w < w1) pack_scalar(src, n, code, spec, w, w1, gc, sep); /* the last, partial word and the guard words */
}
#endif

/* Pack characters [0, n) of one sequence into code[0, nwords) / spec[0, nwords) (nwords covers the
 * guard words of text.cuh: everything past the text is zero). *gc / *sep are added to. */
void andi_host_pack(const char *chars, size_t n, uint64_t *code, uint64_t *spec, size_t nwords, uint64_t *gc, uint64_t *sep) {
	const unsigned char *src = (const unsigned char *)chars;
#if defined(__x86_64__)
	static int have_avx2 = -1;
	if (have_avx2 < 0) have_avx2 = __builtin_cpu_supports("avx2") ? 1 : 0;
	if (have_avx2) {
		pack_avx2(src, n, code, spec, 0, nwords, gc, sep);
		return;
	}
#endif
	pack_scalar(src, n, code, spec, 0, nwords, gc, sep);
}

/* A whole pool: sequence k goes to code + word_off[k]; gc[k] / sep[k] are set. Threads over
 * sequences [k0, k1) (a chunk of the pool, so that the caller can upload chunk by chunk). */
void andi_host_pack_pool(const char *const *seqs, const size_t *lens, const size_t *word_off, const size_t *nwords, size_t k0, size_t k1,
						 uint64_t *code, uint64_t *spec, uint64_t *gc, uint64_t *sep) {
#pragma omp parallel for schedule(dynamic, 1)
	for (size_t k = k0; k < k1; k++) {
		uint64_t g = 0, s = 0;
		andi_host_pack(seqs[k], lens[k], code + word_off[k], spec + word_off[k], nwords[k], &g, &s);
		gc[k] = g, sep[k] = s;
	}
}
