/* Zeroing allocator for the oracle build (TEST INFRASTRUCTURE).
 *
 * The reference reads memory it never wrote (extract_sarea leaves the last row and
 * column of the search area unwritten, MIMC_module.c:869-871; GMA_float_conv2 leaves
 * the border of `out` unwritten but scans all of it, :2539-2567).  Forcing every
 * malloc to return zeroed memory makes those reads deterministic without touching
 * the reference sources.  Injected with `gcc -include zalloc.h`.
 */
#ifndef MIMC3_ZALLOC_H
#define MIMC3_ZALLOC_H
#include <stdlib.h>
#define malloc(n) calloc(1, (n))
#endif
