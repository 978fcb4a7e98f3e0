/* oracle/shim/config.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 * Stands in for the autoconf-generated config.h the reference includes at
 * src/global.h:12 and src/esa.h:9 (autotools are not installed here). */
#ifndef ANDI_ORACLE_SHIM_CONFIG_H
#define ANDI_ORACLE_SHIM_CONFIG_H
#define VERSION "1.15-oracle-shim"
#define HAVE_STRCHRNUL 1
#define HAVE_REALLOCARRAY 1
#endif
