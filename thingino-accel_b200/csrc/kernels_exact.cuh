/*
 * kernels_exact.cuh -- "reference-semantics" CUDA kernels for every layer type.
 *
 * One device "point function" per layer type computes ONE output element exactly as
 * the reference's portable C does (same integer accumulation, same fp32 operation
 * sequence with explicit round-to-nearest intrinsics so nvcc cannot contract to FMA,
 * the x86 float->int overflow rule).  The same point function is driven three ways:
 *   - parallel kernels (one thread per output element) for hazard-free layers,
 *   - pass kernels (one launch per output channel) for in-place NCHW convs,
 *   - a literal single-thread kernel that walks the reference's loop nest in order
 *     for anything whose aliasing makes the loop order observable.
 * Faster specialised kernels (conv_tc.cu, fused epilogues) are checked against these.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mars_internal.h"

namespace marsb200 {

/* plain-old-data copy of an Op for kernel arguments */
struct KOp {
    int kind, mode;
    int64_t in0, in1, in2, out, w, bias;
    int ic, ih, iw, oc, oh, ow, kh, kw, sh, sw, pt, pl;
    int coff;
    float f0, f1, f2;
    unsigned long long n;
    int lut, post_relu, post_lut;
    /* in-place 1x1 conv with the following SIGMOID and MUL folded in (k_conv1x1_inplace_reg): S = table lut_s[y] stored at out_s,
     * Z = table lut_z[y] at out_z (-1 = not stored), store_y = 0: the pre-activation plane itself is dead */
    int64_t out_s, out_z;
    int lut_s, lut_z, store_y;
    int fast_bin;    /* int8 mul / add: the host proved |result| < 2^31 and finite scales, the conversion-light sequence is exact */
    int pass_oc;     /* EXEC_OC_PASSES: the output channel of this launch */
    int use_scratch; /* write the pass to scratch instead of the output plane */
};

/* per-image addressing: arena offset -> device pointer */
struct Img {
    const uint8_t *w;   /* weights: valid for offsets [0, W) */
    uint8_t *s_minus_W; /* slot base minus W: valid for offsets [W, arena) */
    int64_t W;
    uint8_t *scratch;
};

__device__ __forceinline__ Img make_img(const ArenaView &v, int img) {
    Img im;
    im.w = v.wbase;
    im.s_minus_W = v.sbase + (uint64_t)img * v.slot_stride - v.W;
    im.W = (int64_t)v.W;
    im.scratch = v.scratch ? v.scratch + (uint64_t)img * v.scratch_stride : nullptr;
    return im;
}

/* read accessor over an operand that starts at arena offset `off`.  XL=false: the operand
 * lies entirely in one region (one base pointer); XL=true: per-byte translation, used when an
 * operand straddles the weights/slot boundary (e.g. the bias over-read of SURVEY C.3). */
template <bool XL>
struct Rd {
    const uint8_t *p;
    const uint8_t *w;
    const uint8_t *s_minus_W;
    int64_t off, W;
    __device__ __forceinline__ Rd(const Img &im, int64_t o) : w(im.w), s_minus_W(im.s_minus_W), off(o), W(im.W) {
        p = (o < im.W ? im.w : (const uint8_t *)im.s_minus_W) + o;
    }
    __device__ __forceinline__ uint8_t u8(int64_t i) const {
        if (XL) { int64_t a = off + i; return a < W ? w[a] : s_minus_W[a]; }
        return p[i];
    }
    __device__ __forceinline__ int8_t i8(int64_t i) const { return (int8_t)u8(i); }
    /* 4-byte little-endian load at BYTE index i, alignment-agnostic */
    __device__ __forceinline__ uint32_t u32(int64_t i) const {
        if (!XL && ((reinterpret_cast<uintptr_t>(p + i) & 3) == 0)) return *reinterpret_cast<const uint32_t *>(p + i);
        return (uint32_t)u8(i) | ((uint32_t)u8(i + 1) << 8) | ((uint32_t)u8(i + 2) << 16) | ((uint32_t)u8(i + 3) << 24);
    }
    __device__ __forceinline__ float f32(int64_t elem) const { return __uint_as_float(u32(elem * 4)); }
};

__device__ __forceinline__ uint8_t *wr_ptr(const Img &im, int64_t off) { return im.s_minus_W + off; }

/* ---- arithmetic contracts ---------------------------- */
/* (int32_t)v as x86 cvttss2si: NaN and |v| >= 2^31 give INT_MIN (CUDA's cvt.rzi saturates instead) */
__device__ __forceinline__ int32_t f2i_x86(float v) {
    if (!(v > -2147483904.0f && v < 2147483648.0f)) return (int32_t)0x80000000;
    return __float2int_rz(v);
}
__device__ __forceinline__ int8_t clamp_i8(int32_t r) { return (int8_t)(r > 127 ? 127 : (r < -128 ? -128 : r)); }
/* reference src/mars/mxu_conv.c:663-666 */
__device__ __forceinline__ int8_t requant_conv(int32_t acc, float cs) {
    float scaled = __fmul_rn(__int2float_rn(acc), cs);
    float biased = __fadd_rn(scaled, scaled >= 0.0f ? 0.5f : -0.5f);
    return clamp_i8(f2i_x86(biased));
}
/* reference src/mars/mars_runtime.c:831,898 */
__device__ __forceinline__ int8_t requant_mul_inv(float y, float inv) {
    return clamp_i8(f2i_x86(__fadd_rn(__fmul_rn(y, inv), 0.5f)));
}
/* reference src/mars/mars_runtime.c:764,1147 */
__device__ __forceinline__ int8_t requant_div(float y, float scale) {
    return clamp_i8(f2i_x86(__fadd_rn(__fdiv_rn(y, scale), 0.5f)));
}

/* ---- point functions ------------------------------------------------------ */
template <bool XL>
__device__ __forceinline__ int32_t bias_i32(const Img &im, const KOp &o, int oc) {
    if (o.bias < 0) return 0;
    Rd<XL> b(im, o.bias);
    return (int32_t)b.u32(4 * (int64_t)oc); /* raw 4-byte load whatever the bias dtype */
}

/* reference src/mars/mxu_conv.c:646-666 */
template <bool XL>
__device__ __forceinline__ int8_t conv_i8_nchw_point(const Img &im, const KOp &o, int oc, int oh, int ow) {
    Rd<XL> in(im, o.in0), w(im, o.w);
    uint32_t acc = (uint32_t)bias_i32<XL>(im, o, oc); /* wrap-around like x86 */
    const int64_t wbase = (int64_t)oc * o.ic * o.kh * o.kw;
    const int ih0 = oh * o.sh - o.pt, iw0 = ow * o.sw - o.pl;
    for (int ic = 0; ic < o.ic; ic++) {
        const int64_t cb = (int64_t)ic * o.ih * o.iw;
        for (int y = 0; y < o.kh; y++) {
            int ih = ih0 + y;
            if (ih < 0 || ih >= o.ih) continue;
            for (int x = 0; x < o.kw; x++) {
                int iw = iw0 + x;
                if (iw < 0 || iw >= o.iw) continue;
                acc += (uint32_t)((int32_t)in.i8(cb + (int64_t)ih * o.iw + iw) *
                                  (int32_t)w.i8(wbase + ((int64_t)ic * o.kh + y) * o.kw + x));
            }
        }
    }
    return requant_conv((int32_t)acc, o.f0);
}

/* reference src/mars/mxu_conv.c:730-753 */
template <bool XL>
__device__ __forceinline__ int8_t conv_i8_nhwc_point(const Img &im, const KOp &o, int oc, int oh, int ow) {
    Rd<XL> in(im, o.in0), w(im, o.w);
    uint32_t acc = (uint32_t)bias_i32<XL>(im, o, oc);
    const int64_t wbase = (int64_t)oc * o.kh * o.kw * o.ic;
    const int ih0 = oh * o.sh - o.pt, iw0 = ow * o.sw - o.pl;
    for (int y = 0; y < o.kh; y++) {
        int ih = ih0 + y;
        if (ih < 0 || ih >= o.ih) continue;
        for (int x = 0; x < o.kw; x++) {
            int iw = iw0 + x;
            if (iw < 0 || iw >= o.iw) continue;
            const int64_t ib = ((int64_t)ih * o.iw + iw) * o.ic, wb = wbase + ((int64_t)y * o.kw + x) * o.ic;
            for (int ic = 0; ic < o.ic; ic++) acc += (uint32_t)((int32_t)in.i8(ib + ic) * (int32_t)w.i8(wb + ic));
        }
    }
    return requant_conv((int32_t)acc, o.f0);
}

/* restated depthwise (parity unpinned): one input channel per output channel */
template <bool XL>
__device__ __forceinline__ int8_t dw_i8_point(const Img &im, const KOp &o, int nhwc, int c, int oh, int ow) {
    Rd<XL> in(im, o.in0), w(im, o.w);
    uint32_t acc = (uint32_t)bias_i32<XL>(im, o, c);
    for (int y = 0; y < o.kh; y++) {
        int ih = oh * o.sh - o.pt + y;
        if (ih < 0 || ih >= o.ih) continue;
        for (int x = 0; x < o.kw; x++) {
            int iw = ow * o.sw - o.pl + x;
            if (iw < 0 || iw >= o.iw) continue;
            int64_t ii = nhwc ? ((int64_t)ih * o.iw + iw) * o.ic + c : ((int64_t)c * o.ih + ih) * o.iw + iw;
            acc += (uint32_t)((int32_t)in.i8(ii) * (int32_t)w.i8(((int64_t)c * o.kh + y) * o.kw + x));
        }
    }
    return requant_conv((int32_t)acc, o.f0);
}

/* reference src/mars/mxu_conv.c:689-706: sum = bias; sum += in*w sequentially, no FMA */
template <bool XL>
__device__ __forceinline__ float conv_f32_nchw_point(const Img &im, const KOp &o, int oc, int oh, int ow) {
    Rd<XL> in(im, o.in0), w(im, o.w);
    float sum = 0.0f;
    if (o.bias >= 0) { Rd<XL> b(im, o.bias); sum = b.f32(oc); }
    const int64_t wbase = (int64_t)oc * o.ic * o.kh * o.kw;
    const int ih0 = oh * o.sh - o.pt, iw0 = ow * o.sw - o.pl;
    for (int ic = 0; ic < o.ic; ic++) {
        const int64_t cb = (int64_t)ic * o.ih * o.iw;
        for (int y = 0; y < o.kh; y++) {
            int ih = ih0 + y;
            if (ih < 0 || ih >= o.ih) continue;
            for (int x = 0; x < o.kw; x++) {
                int iw = iw0 + x;
                if (iw < 0 || iw >= o.iw) continue;
                sum = __fadd_rn(sum, __fmul_rn(in.f32(cb + (int64_t)ih * o.iw + iw),
                                               w.f32(wbase + ((int64_t)ic * o.kh + y) * o.kw + x)));
            }
        }
    }
    return sum;
}

/* flat element i of a unary/binary/bn layer; writes the result itself (1 or 4 bytes) */
template <bool XL>
__device__ __forceinline__ void flat_point(const Img &im, const KOp &o, const uint8_t *cpool, int64_t i) {
    uint8_t *out = wr_ptr(im, o.out);
    switch (o.kind) {
        case OP_BYTE_RELU: { /* reference src/mars/mars_runtime.c:700-707 */
            int8_t v = (int8_t)out[i];
            if (v < 0) out[i] = 0;
            break;
        }
        case OP_SIGMOID_I8: case OP_RELU_I8: case OP_LUT_I8: {
            Rd<XL> a(im, o.in0);
            out[i] = cpool[o.lut + (int)a.i8(i) + 128];
            break;
        }
        case OP_SIGMOID_F32: { /* :747; expf is CUDA's -> tolerance path */
            Rd<XL> a(im, o.in0);
            float x = a.f32(i);
            reinterpret_cast<float *>(out)[i] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
            break;
        }
        case OP_RELU_F32: { /* :1070 */
            Rd<XL> a(im, o.in0);
            float x = a.f32(i);
            reinterpret_cast<float *>(out)[i] = x > 0.0f ? x : __fmul_rn(x, o.f0);
            break;
        }
        case OP_MUL_I8: case OP_ADD_I8: { /* :827-834, :894-901 */
            Rd<XL> a(im, o.in0), b(im, o.in1);
            float va = __fmul_rn((float)a.i8(i), o.f0), vb = __fmul_rn((float)b.i8(i), o.f1);
            float y = o.kind == OP_MUL_I8 ? __fmul_rn(va, vb) : __fadd_rn(va, vb);
            out[i] = (uint8_t)requant_mul_inv(y, o.f2);
            break;
        }
        case OP_MUL_F32: case OP_ADD_F32: { /* :812-814, :879-881 */
            Rd<XL> a(im, o.in0), b(im, o.in1);
            float x = a.f32(i), y = b.f32(i);
            reinterpret_cast<float *>(out)[i] = o.kind == OP_MUL_F32 ? __fmul_rn(x, y) : __fadd_rn(x, y);
            break;
        }
        case OP_BN_I8: case OP_BN_F32: { /* :1119-1153, NCHW indexing; o.ic = C, plane = ih*iw */
            int64_t plane = (int64_t)o.ih * o.iw;
            int ci = (int)((i / plane) % o.ic);
            float sc = 1.0f, bi = 0.0f;
            if (o.in1 >= 0) { Rd<XL> s(im, o.in1); sc = s.f32(ci); }
            if (o.in2 >= 0) { Rd<XL> b(im, o.in2); bi = b.f32(ci); }
            Rd<XL> a(im, o.in0);
            if (o.kind == OP_BN_F32) {
                reinterpret_cast<float *>(out)[i] = __fadd_rn(__fmul_rn(a.f32(i), sc), bi);
            } else {
                float x = __fmul_rn((float)a.i8(i), o.f0);
                float y = __fadd_rn(__fmul_rn(x, sc), bi);
                out[i] = (uint8_t)requant_div(y, o.f1);
            }
            break;
        }
        default: break;
    }
}

/* reference src/mars/mars_runtime.c:934-956 (NHWC indexing, pads ignored, init -128) */
template <bool XL>
__device__ __forceinline__ int8_t maxpool_point(const Img &im, const KOp &o, int c, int oh, int ow) {
    Rd<XL> in(im, o.in0);
    int8_t mx = -128;
    for (int y = 0; y < o.kh; y++)
        for (int x = 0; x < o.kw; x++) {
            int ih = oh * o.sh + y, iw = ow * o.sw + x;
            if (ih < o.ih && iw < o.iw) {
                int8_t v = in.i8(((int64_t)ih * o.iw + iw) * o.ic + c);
                if (v > mx) mx = v;
            }
        }
    return mx;
}

/* reference src/mars/mars_runtime.c:1027-1040 */
template <bool XL>
__device__ __forceinline__ int8_t upsample_point(const Img &im, const KOp &o, int c, int oh, int ow) {
    Rd<XL> in(im, o.in0);
    int ih = oh / o.sh; if (ih >= o.ih) ih = o.ih - 1;
    int iw = ow / o.sw; if (iw >= o.iw) iw = o.iw - 1;
    return in.i8(((int64_t)ih * o.iw + iw) * o.ic + c);
}

/* ---- parallel kernels ----------------------------------------------------- */
/* grid = (ceil(elements/256), images) */
template <bool XL>
__global__ void __launch_bounds__(256) k_conv_point(ArenaView v, KOp o) {
    const Img im = make_img(v, blockIdx.y);
    const int64_t P = (int64_t)o.oh * o.ow;
    const int64_t total = (o.mode == EXEC_OC_PASSES || o.mode == EXEC_OC_PASSES_SCRATCH) ? P : P * o.oc;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    int oc, oh, ow;
    int64_t oidx;
    if (o.kind == OP_CONV_I8_NHWC || (o.kind == OP_DW_I8 && o.coff)) { /* pixel-major, channel fastest */
        oc = (int)(t % o.oc); int64_t p = t / o.oc; oh = (int)(p / o.ow); ow = (int)(p % o.ow);
        oidx = t;
    } else {
        if (o.mode == EXEC_OC_PASSES || o.mode == EXEC_OC_PASSES_SCRATCH) { oc = o.pass_oc; }
        else { oc = (int)(t / P); t -= (int64_t)oc * P; }
        oh = (int)(t / o.ow); ow = (int)(t % o.ow);
        oidx = (int64_t)oc * P + t;
    }
    if (o.kind == OP_CONV_F32_NCHW) {
        float r = conv_f32_nchw_point<XL>(im, o, oc, oh, ow);
        float *dst = o.use_scratch ? reinterpret_cast<float *>(im.scratch) + t
                                   : reinterpret_cast<float *>(wr_ptr(im, o.out)) + oidx;
        *dst = r;
        return;
    }
    int8_t r;
    if (o.kind == OP_CONV_I8_NCHW) r = conv_i8_nchw_point<XL>(im, o, oc, oh, ow);
    else if (o.kind == OP_CONV_I8_NHWC) r = conv_i8_nhwc_point<XL>(im, o, oc, oh, ow);
    else r = dw_i8_point<XL>(im, o, o.coff, oc, oh, ow);
    if (o.post_relu && r < 0) r = 0;
    if (o.post_lut >= 0) r = (int8_t)v.cpool[o.post_lut + (int)r + 128];
    if (o.use_scratch) im.scratch[t] = (uint8_t)r;
    else wr_ptr(im, o.out)[oidx] = (uint8_t)r;
}

/* second half of a staged pass: scratch plane -> output plane pass_oc */
__global__ void __launch_bounds__(256) k_pass_commit(ArenaView v, KOp o, int es) {
    const Img im = make_img(v, blockIdx.y);
    const int64_t bytes = (int64_t)o.oh * o.ow * es;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < bytes) wr_ptr(im, o.out)[(int64_t)o.pass_oc * bytes + t] = im.scratch[t];
}

/* NHWC conv in place: thread per pixel, output channels in reference order with immediate stores */
template <bool XL>
__global__ void __launch_bounds__(128) k_conv_nhwc_pixel_serial(ArenaView v, KOp o) {
    const Img im = make_img(v, blockIdx.y);
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (int64_t)o.oh * o.ow) return;
    int oh = (int)(p / o.ow), ow = (int)(p % o.ow);
    volatile uint8_t *out = wr_ptr(im, o.out) + p * o.oc;
    for (int oc = 0; oc < o.oc; oc++) {
        int8_t r = conv_i8_nhwc_point<XL>(im, o, oc, oh, ow);
        if (o.post_relu && r < 0) r = 0;
        if (o.post_lut >= 0) r = (int8_t)v.cpool[o.post_lut + (int)r + 128];
        out[oc] = (uint8_t)r;
        __threadfence_block();
    }
}


/* NCHW 1x1 conv whose output planes ARE its input planes (out == in, same plane size): the
 * reference's oc-outermost loop (src/mars/mxu_conv.c:642-669) makes pass oc read planes ic < oc in
 * their OUTPUT form.  A 1x1 conv couples no two pixels, so one thread per pixel walking oc in
 * order -- and feeding each result back into its private copy of the pixel's channel vector --
 * reproduces that exactly.  Channel vector and weights live in shared memory as 4-byte words
 * (dp4a); blockDim.x = 128 pixels.  smem = (Ci/4)*128*4 + Co*(Ci/4)*4 bytes. */
__global__ void __launch_bounds__(128) k_conv1x1_nchw_inplace(ArenaView v, KOp o) {
    extern __shared__ uint32_t smem_w[];
    const Img im = make_img(v, blockIdx.y);
    const int ci4 = o.ic >> 2;
    uint32_t *xs = smem_w;                  /* [ci4][128] */
    uint32_t *ws = smem_w + ci4 * 128;      /* [oc][ci4]  */
    const int64_t P = (int64_t)o.oh * o.ow;
    const int64_t p = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const uint8_t *in = im.s_minus_W + o.in0;
    const uint8_t *w = (o.w < im.W ? im.w : (const uint8_t *)im.s_minus_W) + o.w;
    for (int i = threadIdx.x; i < o.oc * ci4; i += 128) {
        const uint8_t *q = w + (int64_t)i * 4;
        ws[i] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
    }
    if (p < P)
        for (int k = 0; k < ci4; k++) {
            const uint8_t *q = in + (int64_t)(4 * k) * P + p;
            xs[k * 128 + threadIdx.x] = (uint32_t)q[0] | ((uint32_t)q[P] << 8) | ((uint32_t)q[2 * P] << 16) | ((uint32_t)q[3 * P] << 24);
        }
    __syncthreads();
    if (p >= P) return;
    uint8_t *out = wr_ptr(im, o.out);
    for (int oc = 0; oc < o.oc; oc++) {
        int acc = bias_i32<false>(im, o, oc);
        const uint32_t *wr = ws + oc * ci4;
        for (int k = 0; k < ci4; k++) acc = __dp4a((int)xs[k * 128 + threadIdx.x], (int)wr[k], acc);
        const int8_t r = requant_conv(acc, o.f0);
        out[(int64_t)oc * P + p] = (uint8_t)r;
        if (oc < o.ic) { /* feed the result back: later passes read it as input plane oc */
            const int sh = (oc & 3) * 8;
            uint32_t &word = xs[(oc >> 2) * 128 + threadIdx.x];
            word = (word & ~(0xFFu << sh)) | ((uint32_t)(uint8_t)r << sh);
        }
    }
}

template <bool XL>
__global__ void __launch_bounds__(256) k_flat(ArenaView v, KOp o) {
    const Img im = make_img(v, blockIdx.y);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)o.n) flat_point<XL>(im, o, v.cpool, i);
}

/* maxpool / upsample / concat: one thread per output byte, NHWC index order (c fastest) */
template <bool XL>
__global__ void __launch_bounds__(256) k_spatial(ArenaView v, KOp o) {
    const Img im = make_img(v, blockIdx.y);
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint8_t *out = wr_ptr(im, o.out);
    if (o.kind == OP_CONCAT) {
        if (t >= (int64_t)o.n) return; /* n = oh*ow*ic */
        int c = (int)(t % o.ic);
        int64_t p = t / o.ic;
        Rd<XL> in(im, o.in0);
        out[p * o.oc + o.coff + c] = in.u8(t);
        return;
    }
    if (o.kind == OP_CONCAT_PERIODIC) { /* A[src0+j] = A0[src0 + (j mod s)], s <= j < len+s */
        if (t >= (int64_t)o.n) return;
        int64_t j = t + o.coff;
        out[j] = out[j % o.coff];
        return;
    }
    int64_t total = (int64_t)o.oh * o.ow * o.ic;
    if (t >= total) return;
    int c = (int)(t % o.ic);
    int64_t p = t / o.ic;
    int oh = (int)(p / o.ow), ow = (int)(p % o.ow);
    out[t] = (uint8_t)(o.kind == OP_MAXPOOL ? maxpool_point<XL>(im, o, c, oh, ow) : upsample_point<XL>(im, o, c, oh, ow));
}

/* literal reference loop order, one thread per image: exact under ANY aliasing */
template <bool XL>
__global__ void k_serial(ArenaView v, KOp o) {
    if (threadIdx.x != 0) return;
    const Img im = make_img(v, blockIdx.x);
    volatile uint8_t *out = wr_ptr(im, o.out);
    switch (o.kind) {
        case OP_CONV_I8_NCHW: /* oc, oh, ow (src/mars/mxu_conv.c:642-645) */
            for (int oc = 0; oc < o.oc; oc++)
                for (int oh = 0; oh < o.oh; oh++)
                    for (int ow = 0; ow < o.ow; ow++) {
                        int8_t r = conv_i8_nchw_point<true>(im, o, oc, oh, ow);
                        if (o.post_relu && r < 0) r = 0;
                        if (o.post_lut >= 0) r = (int8_t)v.cpool[o.post_lut + (int)r + 128];
                        out[((int64_t)oc * o.oh + oh) * o.ow + ow] = (uint8_t)r;
                    }
            break;
        case OP_CONV_I8_NHWC: /* oh, ow, oc (src/mars/mxu_conv.c:726-730) */
            for (int oh = 0; oh < o.oh; oh++)
                for (int ow = 0; ow < o.ow; ow++)
                    for (int oc = 0; oc < o.oc; oc++) {
                        int8_t r = conv_i8_nhwc_point<true>(im, o, oc, oh, ow);
                        if (o.post_relu && r < 0) r = 0;
                        if (o.post_lut >= 0) r = (int8_t)v.cpool[o.post_lut + (int)r + 128];
                        out[((int64_t)oh * o.ow + ow) * o.oc + oc] = (uint8_t)r;
                    }
            break;
        case OP_DW_I8:
            for (int c = 0; c < o.oc; c++)
                for (int oh = 0; oh < o.oh; oh++)
                    for (int ow = 0; ow < o.ow; ow++) {
                        int8_t r = dw_i8_point<true>(im, o, o.coff, c, oh, ow);
                        int64_t oi = o.coff ? ((int64_t)oh * o.ow + ow) * o.oc + c : ((int64_t)c * o.oh + oh) * o.ow + ow;
                        out[oi] = (uint8_t)r;
                    }
            break;
        case OP_CONV_F32_NCHW: {
            volatile float *fo = reinterpret_cast<volatile float *>(wr_ptr(im, o.out));
            for (int oc = 0; oc < o.oc; oc++)
                for (int oh = 0; oh < o.oh; oh++)
                    for (int ow = 0; ow < o.ow; ow++)
                        fo[((int64_t)oc * o.oh + oh) * o.ow + ow] = conv_f32_nchw_point<true>(im, o, oc, oh, ow);
            break;
        }
        case OP_MAXPOOL: /* c, oh, ow (src/mars/mars_runtime.c:934-936) */
            for (int c = 0; c < o.ic; c++)
                for (int oh = 0; oh < o.oh; oh++)
                    for (int ow = 0; ow < o.ow; ow++)
                        out[((int64_t)oh * o.ow + ow) * o.ic + c] = (uint8_t)maxpool_point<true>(im, o, c, oh, ow);
            break;
        case OP_UPSAMPLE: /* oh, ow, c (:1027-1035) */
            for (int oh = 0; oh < o.oh; oh++)
                for (int ow = 0; ow < o.ow; ow++)
                    for (int c = 0; c < o.ic; c++)
                        out[((int64_t)oh * o.ow + ow) * o.ic + c] = (uint8_t)upsample_point<true>(im, o, c, oh, ow);
            break;
        case OP_CONCAT: { /* h, w, c ascending (:987-994) */
            Rd<true> in(im, o.in0);
            for (int64_t t = 0; t < (int64_t)o.n; t++) {
                int c = (int)(t % o.ic);
                int64_t p = t / o.ic;
                out[p * o.oc + o.coff + c] = in.u8(t);
            }
            break;
        }
        default: /* flat layers, ascending index */
            for (int64_t i = 0; i < (int64_t)o.n; i++) flat_point<true>(im, o, v.cpool, i);
            break;
    }
}

} // namespace marsb200
