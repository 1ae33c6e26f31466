/*
 * nna.h -- device bring-up calls applications make around mars_* (B200 build).
 * Same names/returns as the reference include/nna.h:26-80 and
 * include/nna_types.h:18-23,66-71; "the NNA" is now a CUDA device.
 */
#ifndef THINGINO_ACCEL_NNA_H
#define THINGINO_ACCEL_NNA_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#define NNA_SUCCESS 0
#define NNA_ERROR_INIT (-1)
#define NNA_ERROR_DEVICE (-2)
#define NNA_ERROR_MEMORY (-3)
#define NNA_ERROR_INVALID (-4)
#define NNA_ERROR_TIMEOUT (-5)

/* reference include/nna_types.h:66-71 */
typedef struct {
    uint32_t oram_vbase;
    uint32_t oram_pbase;
    uint32_t oram_size;
    uint32_t version;
} nna_hw_info_t;

int nna_init(void);                       /* reference src/device.c:133: here cudaSetDevice + context */
void nna_deinit(void);
int nna_get_hw_info(nna_hw_info_t *info); /* oram_size = shared memory per SM, version = 100 (sm_100a) */
int nna_is_ready(void);
const char *nna_get_version(void);
int nna_lock(void);                       /* stubs returning success in the reference (src/device.c:435-443) */
int nna_unlock(void);

/* inner seam, reference src/device_internal.h:12-31 */
void *nna_device_get_ddr(void);        /* host mirror of the most recently loaded model's arena */
uint32_t nna_device_get_ddr_pbase(void); /* low 32 bits of the device arena address */
void *nna_device_get_oram(void);       /* NULL: no ORAM; shared memory/TMEM are not host-mappable */
int nna_device_get_fd(void);           /* -1 */
int nna_device_get_memfd(void);        /* -1 */
void *nna_device_get_nndma_io(void);   /* NULL */
void *nna_device_get_nndma_desram(void); /* NULL */
#ifdef __cplusplus
}
#endif
#endif
