/*
 * kernels_fast.cuh -- vectorised variants of the memory-bound layer kernels.
 *
 * Each is bit-identical to its point function in kernels_exact.cuh (same fp32 operation
 * sequence, same tables) and is only selected for hazard-free ops (EXEC_PARALLEL, operands not
 * straddling the weights/slot boundary) whose operands are 16-byte aligned; everything else
 * stays on the exact kernels.  16 bytes per thread per access, 256-byte tables staged in
 * shared memory, grid = (chunks, images).
 */
#pragma once
#include "kernels_exact.cuh"

namespace marsb200 {

__device__ __forceinline__ uint32_t lut4(const uint8_t *lut, uint32_t w) {
    return (uint32_t)lut[(uint8_t)((w & 0xFF) ^ 0x80)] | ((uint32_t)lut[(uint8_t)(((w >> 8) & 0xFF) ^ 0x80)] << 8) |
           ((uint32_t)lut[(uint8_t)(((w >> 16) & 0xFF) ^ 0x80)] << 16) | ((uint32_t)lut[(uint8_t)((w >> 24) ^ 0x80)] << 24);
}

/* out[i] = table[in[i]] (sigmoid / relu / fused unary chains), table index = value + 128 */
__global__ void __launch_bounds__(256) k_lut16(ArenaView v, KOp o) {
    __shared__ uint8_t lut[256];
    lut[threadIdx.x] = v.cpool[o.lut + threadIdx.x];
    __syncthreads();
    const Img im = make_img(v, blockIdx.y);
    const uint8_t *in = im.s_minus_W + o.in0;
    uint8_t *out = im.s_minus_W + o.out;
    const int64_t nv = (int64_t)(o.n >> 4);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 a = reinterpret_cast<const uint4 *>(in)[i];
        a.x = lut4(lut, a.x); a.y = lut4(lut, a.y); a.z = lut4(lut, a.z); a.w = lut4(lut, a.w);
        reinterpret_cast<uint4 *>(out)[i] = a;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(o.n & 15)) {
        const int64_t i = (nv << 4) + threadIdx.x;
        out[i] = lut[(uint8_t)(in[i] ^ 0x80)];
    }
}

__device__ __forceinline__ uint32_t bin4(uint32_t a, uint32_t b, int is_mul, float sa, float sb, float inv) {
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float va = __fmul_rn((float)(int8_t)(a >> (8 * k)), sa), vb = __fmul_rn((float)(int8_t)(b >> (8 * k)), sb);
        const float y = is_mul ? __fmul_rn(va, vb) : __fadd_rn(va, vb);
        r |= (uint32_t)(uint8_t)requant_mul_inv(y, inv) << (8 * k);
    }
    return r;
}

/* int8 mul / add (reference src/mars/mars_runtime.c:818-835, :885-902) */
__global__ void __launch_bounds__(256) k_bin16(ArenaView v, KOp o) {
    const Img im = make_img(v, blockIdx.y);
    const uint8_t *a = im.s_minus_W + o.in0, *b = im.s_minus_W + o.in1;
    uint8_t *out = im.s_minus_W + o.out;
    const int is_mul = o.kind == OP_MUL_I8;
    const int64_t nv = (int64_t)(o.n >> 4);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 x = reinterpret_cast<const uint4 *>(a)[i], y = reinterpret_cast<const uint4 *>(b)[i];
        uint4 r;
        r.x = bin4(x.x, y.x, is_mul, o.f0, o.f1, o.f2); r.y = bin4(x.y, y.y, is_mul, o.f0, o.f1, o.f2);
        r.z = bin4(x.z, y.z, is_mul, o.f0, o.f1, o.f2); r.w = bin4(x.w, y.w, is_mul, o.f0, o.f1, o.f2);
        reinterpret_cast<uint4 *>(out)[i] = r;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(o.n & 15)) {
        const int64_t i = (nv << 4) + threadIdx.x;
        const float va = __fmul_rn((float)(int8_t)a[i], o.f0), vb = __fmul_rn((float)(int8_t)b[i], o.f1);
        out[i] = (uint8_t)requant_mul_inv(is_mul ? __fmul_rn(va, vb) : __fadd_rn(va, vb), o.f2);
    }
}

/* concat input whose channel count equals the output's (every concat of the shipped YOLO files,
 * SURVEY C.4): out[t + coff] = in[t], a shifted flat copy.  VEC = bytes per access. */
template <int VEC>
__global__ void __launch_bounds__(256) k_shift_copy(ArenaView v, KOp o, int periodic) {
    const Img im = make_img(v, blockIdx.y);
    uint8_t *out = im.s_minus_W + o.out + o.coff;
    const uint8_t *in = periodic ? (const uint8_t *)(im.s_minus_W + o.out) : (const uint8_t *)(im.s_minus_W + o.in0);
    const int64_t nv = (int64_t)o.n / VEC;
    const int64_t per = o.coff / VEC; /* periodic: source index wraps every coff bytes */
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t si = periodic ? i % per : i;
        if (VEC == 16) reinterpret_cast<uint4 *>(out)[i] = reinterpret_cast<const uint4 *>(in)[si];
        else if (VEC == 4) reinterpret_cast<uint32_t *>(out)[i] = reinterpret_cast<const uint32_t *>(in)[si];
        else out[i] = in[si];
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)((int64_t)o.n % VEC)) {
        const int64_t i = nv * VEC + threadIdx.x;
        out[i] = periodic ? in[i % o.coff] : in[i];
    }
}

static inline unsigned fast_grid(uint64_t items, unsigned per_block = 256) {
    uint64_t b = (items + per_block - 1) / per_block;
    return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(b, 148 * 16));
}

/* ---- selection + launch ---------------------------------------------------------- */
static inline bool aligned16(int64_t x) { return (x & 15) == 0; }

static inline bool fast_flat_ok(const ArenaView &v, const KOp &o) {
    const int64_t W = (int64_t)v.W;
    if (o.mode != EXEC_PARALLEL || o.n < 64 || o.in0 < W || o.out < W) return false;
    switch (o.kind) {
        case OP_SIGMOID_I8: case OP_RELU_I8: case OP_LUT_I8:
            return o.lut >= 0 && aligned16(o.in0 - W) && aligned16(o.out - W);
        case OP_MUL_I8: case OP_ADD_I8:
            return o.in1 >= W && aligned16(o.in0 - W) && aligned16(o.in1 - W) && aligned16(o.out - W);
        default: return false;
    }
}
static inline void launch_fast_flat(const ArenaView &v, const KOp &o, int n_img, cudaStream_t s) {
    dim3 g(fast_grid(o.n >> 4), n_img);
    if (o.kind == OP_MUL_I8 || o.kind == OP_ADD_I8) k_bin16<<<g, 256, 0, s>>>(v, o);
    else k_lut16<<<g, 256, 0, s>>>(v, o);
}

/* maxpool / upsample over 4 channels per thread (reference src/mars/mars_runtime.c:934-956, :1027-1040; NHWC indexing of
 * shape[1..3] whatever the tag): signed per-byte max with __vmaxs4, window clipped at the bottom/right edge,
 * pads ignored, exactly like the point functions. */
__global__ void __launch_bounds__(256) k_spatial_vec4(ArenaView v, KOp o) {
    const Img im = make_img(v, blockIdx.y);
    const uint32_t *in = reinterpret_cast<const uint32_t *>(im.s_minus_W + o.in0);
    uint32_t *out = reinterpret_cast<uint32_t *>(im.s_minus_W + o.out);
    const int c4 = o.ic >> 2;
    const int64_t total = (int64_t)o.oh * o.ow * c4;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % c4);
        const int64_t p = t / c4;
        const int oh = (int)(p / o.ow), ow = (int)(p - (int64_t)oh * o.ow);
        uint32_t r;
        if (o.kind == OP_MAXPOOL) {
            r = 0x80808080u; /* four times -128 */
            const int ih0 = oh * o.sh, iw0 = ow * o.sw;
            const int ky = min(o.kh, o.ih - ih0), kx = min(o.kw, o.iw - iw0);
            for (int y = 0; y < ky; y++) {
                const uint32_t *row = in + ((int64_t)(ih0 + y) * o.iw + iw0) * c4 + c;
                for (int x = 0; x < kx; x++) r = __vmaxs4(r, row[(int64_t)x * c4]);
            }
        } else {
            int ih = oh / o.sh; if (ih >= o.ih) ih = o.ih - 1;
            int iw = ow / o.sw; if (iw >= o.iw) iw = o.iw - 1;
            r = in[((int64_t)ih * o.iw + iw) * c4 + c];
        }
        out[t] = r;
    }
}

static inline bool fast_spatial_ok(const ArenaView &v, const KOp &o) {
    if (o.mode != EXEC_PARALLEL || o.out < (int64_t)v.W) return false;
    if (o.kind == OP_CONCAT) return o.in0 >= (int64_t)v.W && o.ic == o.oc && o.n >= 64;
    if (o.kind == OP_CONCAT_PERIODIC) return o.coff > 0 && o.n >= 64;
    if (o.kind == OP_MAXPOOL || o.kind == OP_UPSAMPLE)
        return o.in0 >= (int64_t)v.W && o.ic > 0 && o.ic % 4 == 0 && ((o.in0 - (int64_t)v.W) & 3) == 0 && ((o.out - (int64_t)v.W) & 3) == 0 &&
               o.ih > 0 && o.iw > 0 && o.sh > 0 && o.sw > 0;
    return false;
}
static inline void launch_fast_spatial(const ArenaView &v, const KOp &o, int n_img, cudaStream_t s) {
    if (o.kind == OP_MAXPOOL || o.kind == OP_UPSAMPLE) {
        dim3 g(fast_grid((uint64_t)o.oh * o.ow * (o.ic >> 2)), n_img);
        k_spatial_vec4<<<g, 256, 0, s>>>(v, o);
        return;
    }
    const int periodic = o.kind == OP_CONCAT_PERIODIC;
    const int64_t dst = o.out + o.coff, src = periodic ? o.out : o.in0;
    /* slot bases are 1 KiB aligned and W-relative offsets keep their low bits: alignment of the
     * arena offsets minus W is what counts, and v.W is subtracted from both alike */
    const int64_t rel_d = dst - (int64_t)v.W, rel_s = src - (int64_t)v.W;
    int vec = 1;
    if (((rel_d | rel_s) & 15) == 0 && (!periodic || (o.coff & 15) == 0)) vec = 16;
    else if (((rel_d | rel_s) & 3) == 0 && (!periodic || (o.coff & 3) == 0)) vec = 4;
    dim3 g(fast_grid(o.n / vec), n_img);
    if (vec == 16) k_shift_copy<16><<<g, 256, 0, s>>>(v, o, periodic);
    else if (vec == 4) k_shift_copy<4><<<g, 256, 0, s>>>(v, o, periodic);
    else k_shift_copy<1><<<g, 256, 0, s>>>(v, o, periodic);
}

static inline bool fast_conv_nchw_ok(const KOp &) { return false; }
static inline void launch_fast_conv_nchw(const ArenaView &, const KOp &, int, cudaStream_t) {}

} // namespace marsb200
