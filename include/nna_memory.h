/*
 * nna_memory.h -- the allocator seam (reference include/nna_memory.h:28-119,
 * src/memory.c:76-274).  The ORAM/DDR ioctl allocator becomes: pinned host memory
 * registered with the CUDA device ("DDR"), and a bump allocator over a small
 * device scratch region ("ORAM").
 */
#ifndef THINGINO_ACCEL_NNA_MEMORY_H
#define THINGINO_ACCEL_NNA_MEMORY_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
void *nna_malloc(size_t size);                      /* reference src/memory.c:76 */
void *nna_memalign(size_t alignment, size_t size);  /* :126 */
void *nna_calloc(size_t nmemb, size_t size);        /* :160 */
void nna_free(void *ptr);                           /* :174 */
void *nna_oram_malloc(size_t size);                 /* :198 bump allocator (host-visible scratch) */
void nna_oram_free(void *ptr);                      /* :230 */
int nna_oram_get_stats(size_t *total, size_t *used, size_t *free_bytes); /* :236 */
void nna_cache_flush(void *ptr, size_t size);       /* :248 stub in the reference; here: no-op (pinned memory is coherent) */
void nna_cache_invalidate(void *ptr, size_t size);  /* :262 */
#ifdef __cplusplus
}
#endif
#endif
