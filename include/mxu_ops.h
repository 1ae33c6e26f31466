/*
 * mxu_ops.h -- element-wise fp32 helpers and the per-layer convolution entry
 * points of the reference's inner seam (reference include/mxu_ops.h:22-75,
 * declared again at src/mars/mars_runtime.c:22-47).  Kept for source
 * compatibility: host pointers, synchronous, one H2D/D2H round trip per call --
 * the efficient drop-in boundary is mars_run()/mars_b200_* one level up.
 */
#ifndef MXU_OPS_H
#define MXU_OPS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
void mxu_init(void *nna_mem);      /* reference src/mars/mxu_ops.c:145 (no-op off-MIPS) */
int mxu_is_initialized(void);      /* reference src/mars/mxu_ops.c:146: 0 off-MIPS; here 1 once a CUDA context exists */
void mxu_mul_f32(float *out, const float *a, const float *b, size_t count);  /* :148-150 */
void mxu_add_f32(float *out, const float *a, const float *b, size_t count);  /* :152-154 */
void mxu_sub_f32(float *out, const float *a, const float *b, size_t count);  /* :156-158 */
void mxu_relu_f32(float *out, const float *in, size_t count);                /* :160-162 */

/* reference src/mars/mxu_conv.c:630-670: int8 NCHW x OIHW */
void conv2d_int8_mxu(const signed char *input, int in_h, int in_w, int in_c,
                     const signed char *weight, int out_c, int kh, int kw, const int *bias,
                     signed char *output, int out_h, int out_w, int stride_h, int stride_w,
                     int pad_top, int pad_left, float in_scale, float w_scale, float out_scale);
/* reference src/mars/mxu_conv.c:713-757: int8 NHWC x OHWI */
void conv2d_int8_nhwc_mxu(const signed char *input, int in_h, int in_w, int in_c,
                          const signed char *weight, int out_c, int kh, int kw, const int *bias,
                          signed char *output, int out_h, int out_w, int stride_h, int stride_w,
                          int pad_top, int pad_left, float in_scale, float w_scale, float out_scale);
/* reference src/mars/mxu_conv.c:673-710: fp32 NCHW x OIHW, sequential accumulation */
void conv2d_float32_mxu(const float *input, int in_h, int in_w, int in_c,
                        const float *weight, int out_c, int kh, int kw, const float *bias,
                        float *output, int out_h, int out_w, int stride_h, int stride_w,
                        int pad_top, int pad_left, float *scratch);
#ifdef __cplusplus
}
#endif
#endif
