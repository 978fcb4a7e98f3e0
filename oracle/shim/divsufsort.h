/* oracle/shim/divsufsort.h -- TEST INFRASTRUCTURE ONLY.
 * libdivsufsort (Y. Mori; unpinned in the reference's configure.ac:32-38) is not
 * installed in this image. The reference calls exactly one entry point,
 * divsufsort(T, SA, n) at src/esa.c:303. The suffix array of a string is unique, so any
 * correct sorter (unsigned byte order, a proper prefix sorts first) yields the same SA. */
#ifndef ANDI_ORACLE_SHIM_DIVSUFSORT_H
#define ANDI_ORACLE_SHIM_DIVSUFSORT_H
#include <stdint.h>
typedef int32_t saidx_t;
typedef uint8_t sauchar_t;
#ifdef __cplusplus
extern "C" {
#endif
/* returns 0 on success, -1 on bad arguments, -2 on allocation failure (libdivsufsort's codes) */
int divsufsort(const unsigned char *T, saidx_t *SA, saidx_t n);
#ifdef __cplusplus
}
#endif
#endif
