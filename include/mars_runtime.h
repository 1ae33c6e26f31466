/*
 * mars_runtime.h -- the C runtime API applications link against (B200 build).
 *
 * Drop-in for the reference include/mars_runtime.h:19-138: same ten entry points,
 * same error codes, same public struct layout (callers dereference
 * mars_runtime_tensor_t.{desc,vaddr,paddr,alloc_size} and mars_model_t.header
 * directly -- reference src/mars/mars_test.c:62-141, src/mars/mars_yolo_test.c:157-189).
 * Underneath, mars_run() replays CUDA kernels over an HBM arena instead of
 * dispatching to the MIPS MXU/NNA; `vaddr` stays a CPU-dereferenceable (pinned)
 * mirror, `paddr` carries the device address.
 */
#ifndef MARS_RUNTIME_H
#define MARS_RUNTIME_H

#include "mars.h"
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference include/mars_runtime.h:19-29 */
typedef enum {
    MARS_OK = 0,
    MARS_ERR_INVALID_MAGIC = -1,
    MARS_ERR_VERSION_MISMATCH = -2,
    MARS_ERR_ALLOC_FAILED = -3,
    MARS_ERR_INVALID_FILE = -4,
    MARS_ERR_NNA_INIT_FAILED = -5,
    MARS_ERR_LAYER_FAILED = -6,
    MARS_ERR_INVALID_TENSOR = -7,
    MARS_ERR_INVALID_LAYER = -8
} mars_error_t;

/* reference include/mars_runtime.h:32-38 */
typedef struct {
    mars_tensor_t desc; /* descriptor as read from the file */
    void *vaddr;        /* host-visible address (pinned mirror of the arena slot) */
    void *paddr;        /* device address inside the HBM arena */
    size_t alloc_size;  /* bytes the caller may fill: work-buffer size / weight size */
    bool is_external;
} mars_runtime_tensor_t;

/* reference include/mars_runtime.h:41-44 */
typedef struct {
    mars_layer_t desc;
    bool is_executed;
} mars_runtime_layer_t;

/* reference include/mars_runtime.h:47-67.  The B200 library allocates a larger
 * private object and hands out a pointer to this public prefix. */
typedef struct {
    mars_header_t header;
    mars_runtime_tensor_t *tensors;
    mars_runtime_layer_t *layers;
    void *ddr_base;  /* host mirror of the arena: [weights | buf0 | buf1 | (buf2)] */
    void *ddr_paddr; /* device base of image slot 0 (64-bit, not the reference's 32-bit pbase) */
    size_t ddr_size;
    void *oram_base;
    void *oram_paddr;
    size_t oram_size;
    void *weights;
    size_t weights_size;
    uint64_t total_inference_us;
    uint32_t inference_count;
} mars_model_t;

/* reference include/mars_runtime.h:79 / src/mars/mars_runtime.c:351-386 */
mars_error_t mars_load_file(const char *path, mars_model_t **model);
/* reference include/mars_runtime.h:88 / src/mars/mars_runtime.c:126-349 */
mars_error_t mars_load_memory(const void *data, size_t size, mars_model_t **model);
/* reference include/mars_runtime.h:94 / src/mars/mars_runtime.c:388-393 */
void mars_free(mars_model_t *model);
/* reference include/mars_runtime.h:102,110 / src/mars/mars_runtime.c:395-411 */
mars_runtime_tensor_t *mars_get_input(mars_model_t *model, int index);
mars_runtime_tensor_t *mars_get_output(mars_model_t *model, int index);
/* reference include/mars_runtime.h:117 / src/mars/mars_runtime.c:439-459 */
mars_error_t mars_run(mars_model_t *model);
/* reference include/mars_runtime.h:123 / src/mars/mars_runtime.c:58-76 */
const char *mars_get_error_string(mars_error_t err);
/* reference include/mars_runtime.h:128,133 / src/mars/mars_runtime.c:413-419 */
int mars_get_num_inputs(mars_model_t *model);
int mars_get_num_outputs(mars_model_t *model);
/* reference include/mars_runtime.h:138 / src/mars/mars_runtime.c:421-434 */
void mars_print_summary(mars_model_t *model);

#ifdef __cplusplus
}
#endif
#endif /* MARS_RUNTIME_H */
