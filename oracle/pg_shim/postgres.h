/* TEST INFRASTRUCTURE ONLY -- a stand-in for <postgres.h> that lets the reference's k-means
 * (/root/reference/lantern_hnsw/src/hnsw/product_quantization.c, compiled UNMODIFIED from where it lies by
 * oracle/Makefile target `refpq`) build without a Postgres tree.  Everything product_quantization.c takes from
 * Postgres is listed here; nothing else is stubbed:
 *
 *   uint8 / uint32 / float4 / bool   c.h typedefs                          -> stdint / float / stdbool
 *   palloc, palloc0                  memory-context allocators             -> malloc / calloc (leaked: the oracle process is short-lived,
 *                                                                             exactly like a transaction-scoped context that is never reset here)
 *   CHECK_FOR_INTERRUPTS()           miscadmin.h, CTRL-C polling           -> no-op
 *   PG_VERSION_NUM                   140000: selects the `random()` branch of get_random_tid (product_quantization.c:28-32)
 *   random()                         libc PRNG behind the initial centres  -> oracle_scripted_random(): the TEST feeds the sequence,
 *                                                                             so both sides start from the same rows (the reference's
 *                                                                             results are otherwise non-deterministic by design,
 *                                                                             test/sql/hnsw_pq_index.sql:85-86)
 */
#ifndef ORACLE_PG_SHIM_POSTGRES_H
#define ORACLE_PG_SHIM_POSTGRES_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

typedef uint8_t uint8;
typedef uint32_t uint32;
typedef float float4;

#define PG_VERSION_NUM 140000
#define palloc(sz) malloc(sz)
#define palloc0(sz) calloc(1, (sz))
#define CHECK_FOR_INTERRUPTS() ((void)0)

long oracle_scripted_random(void);
#define random() oracle_scripted_random()
#endif
