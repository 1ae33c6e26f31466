/*
 * mars_math.h -- fp32 vector/matrix helpers (B200 build).
 * Same three entry points as the reference include/mars_math.h:17-31; host
 * pointers in, host pointers out, synchronous.  The dot product and the GEMM keep
 * the reference's strictly sequential accumulation order (src/mars/mars_math.c:26-55)
 * so results are bit-identical, not merely close.
 */
#pragma once
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
/* dst[i] = a[i] + b[i]            (reference src/mars/mars_math.c:14-29) */
void mars_vec_add_f32(float *dst, const float *a, const float *b, size_t n);
/* sum_i a[i]*b[i], i ascending     (reference src/mars/mars_math.c:31-39) */
float mars_vec_dot_f32(const float *a, const float *b, size_t n);
/* C[MxN] = A[MxK] * B[KxN], k ascending (reference src/mars/mars_math.c:41-55) */
void mars_matmul_f32(float *C, const float *A, const float *B, size_t M, size_t K, size_t N);
#ifdef __cplusplus
}
#endif
