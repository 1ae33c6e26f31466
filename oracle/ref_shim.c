/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE, never linked into the product.
 *
 * Host stand-in for the five device functions the reference's mars runtime needs
 * (reference src/device_internal.h:12-31, implemented for the SoC in src/device.c
 * which opens /dev/mem and /dev/soc-nna and therefore cannot run here), plus a few
 * helpers that drive the *unmodified* reference objects:
 *   - oracle_ref_arena_config(): size of the zeroed "DDR" arena (8 MiB = reference
 *     default, src/mars/mars_runtime.c:209; larger values need the one-line arena
 *     patch applied by build_ref.sh),
 *   - oracle_ref_run_layer(): run exactly one layer of a loaded model through the
 *     reference's own mars_run(),
 *   - oracle_ref_parse_output()/oracle_ref_nms(): the file-static post-process of
 *     reference src/mars/mars_yolo_test.c:80-130, reached by including that TU.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load
 * the resulting oracle/_ref/libmars_ref.so.
 */
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void *g_arena = NULL;
static size_t g_arena_bytes = 0;
static int g_verbose = 0;

/* stdout/stderr chatter of the reference (one printf per conv per run,
 * src/mars/mars_runtime.c:604-624) is routed here by -Dprintf=/-Dfprintf=. */
int oracle_ref_printf(const char *fmt, ...) {
    if (!g_verbose) return 0;
    va_list ap; va_start(ap, fmt);
    int r = vfprintf(stdout, fmt, ap);
    va_end(ap);
    return r;
}
int oracle_ref_fprintf(FILE *f, const char *fmt, ...) {
    if (!g_verbose) return 0;
    va_list ap; va_start(ap, fmt);
    int r = vfprintf(f, fmt, ap);
    va_end(ap);
    return r;
}
void oracle_ref_set_verbose(int v) { g_verbose = v; }

int oracle_ref_arena_config(size_t bytes) {
    free(g_arena);
    g_arena = NULL;
    g_arena_bytes = 0;
    if (bytes == 0) return 0;
    /* slack after the arena: the reference may read a few hundred bytes past
     * short bias blobs -- inside the arena in every shipped
     * model, the slack only keeps a hypothetical overrun off the heap metadata. */
    if (posix_memalign(&g_arena, 4096, bytes + 4096) != 0) return -1;
    memset(g_arena, 0, bytes + 4096);
    g_arena_bytes = bytes;
    return 0;
}
void *oracle_ref_arena(void) { return g_arena; }
size_t oracle_ref_arena_size(void) { return g_arena_bytes; }
/* target of the scripted patch of src/mars/mars_runtime.c:209 */
size_t oracle_ref_ddr_size(void) { return g_arena_bytes ? g_arena_bytes : (size_t)8 * 1024 * 1024; }

/* --- the device seam (reference src/device_internal.h, include/nna.h) --- */
int nna_init(void) {
    if (!g_arena) return oracle_ref_arena_config((size_t)8 * 1024 * 1024);
    return 0;
}
void nna_deinit(void) {}
void *nna_device_get_ddr(void) {
    if (!g_arena) nna_init();
    return g_arena;
}
uint32_t nna_device_get_ddr_pbase(void) { return 0; }
void *nna_device_get_oram(void) { return NULL; }
