/*
 * conv_tf32.cu -- float32 convolution as an implicit GEMM on tcgen05 (kind::tf32, fp32 accumulators in TMEM), sm_100a.
 *
 * Replaces the reference's conv2d_float32_mxu (src/mars/mxu_conv.c:673-710: fp32 NCHW x OIHW, `sum = bias; sum += in * w` in
 * ic -> kh -> kw order, no activation) for hazard-free layers of float32 models.  This is the TOLERANCE path of the north
 * star (<= 1e-3 relative on the logits); the exact-order FFMA-free kernel (kernels_exact.cuh, conv_f32_nchw_point) stays the
 * bit-exact control and is what mars_b200_set_f32_mode(model, 0) selects.
 *
 *   GEMM view   D[M = 128 pixels][N = out channels] = A[M][K] * B[N][K],  K = taps x in channels, both operands K-major.
 *   A operand   a channel-innermost fp32 copy of the input written by a pre-pass (k_f32_to_kmajor): rows padded with k-1
 *               zero columns for stride 1, a 2x2 phase split for stride 2, so that a kernel tap is a flat row shift (the
 *               layouts of conv_tc.cu's int8 copies with 4-byte elements), channels padded to a multiple of 8.
 *   precision   kind::tf32 uses 10 mantissa bits of each fp32 operand.  Mode 1 (tf32): the pre-pass and the weight repack round
 *               to nearest tf32 first (unbiased, ~5e-4 relative per product).  Mode 2 (tf32x3, default): every operand is
 *               split exactly, x = hi + lo with hi = x rounded to tf32 (stored with its 13 low bits zero, so the tensor core
 *               sees exactly hi whether it truncates or rounds -- measured: passing the raw x as `hi` left 2e-4 of error)
 *               and lo = x - hi, and three MMAs per k-step accumulate hi*hi + lo*hi + hi*lo -- ~2^-21 relative per product,
 *               fp32 accumulation, so the result differs from the reference's sequential fp32 sum by summation order only.
 *   epilogue    + fp32 bias, one coalesced 128-byte store per channel and warp into the NCHW output plane.
 *   roles       warps 0..7 epilogue (two per TMEM lane quadrant), warp 8 = TMEM allocator + MMA issuer, warp 9 = TMA producer;
 *               persistent CTAs, one per SM, walking (image, M tile, N tile) units.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "conv_tf32.h"

namespace marsb200 {

namespace {

constexpr int F_BM = 128, F_MAX_TAPS = 36, F_EPI = 8, F_MAX_CO = 1024;

struct F32Params {
    int Co, Ho, Wo, Wp, mflat, plane;
    int n_tile, n_tiles, m_tiles, ntaps, ksteps, stages, acc_bufs, tmem_cols, x3;
    uint32_t idesc, layout, bk; /* bk: bytes of K per pipeline step (128 / 64 / 32) */
    uint32_t a_tile_bytes, b_tile_bytes, stage_bytes;
    int a_shift[F_MAX_TAPS];
    int lo_rows;             /* row offset of the `lo` copy inside the A tensor map (tf32x3) */
    const float *bias;
    uint8_t *out_base;       /* slot 0 of the launch */
    unsigned long long slot_stride;
    long long out_off;
    unsigned wp_magic;
    int img0, n_img;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}\n" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
    for (;;) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(64);
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, {%7, %7, %7, %7}, p;\n\t}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}

struct TileIter {
    int img, rem, tpi, step_img, step_rem;
    __device__ __forceinline__ TileIter(int first, int G, int tiles_per_img) : tpi(tiles_per_img) {
        img = first / tpi; rem = first - img * tpi;
        step_img = G / tpi; step_rem = G - step_img * tpi;
    }
    __device__ __forceinline__ void next() {
        img += step_img; rem += step_rem;
        if (rem >= tpi) { rem -= tpi; img++; }
    }
};

__global__ void __launch_bounds__((F_EPI + 2) * 32, 1)
k_conv_tf32(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const F32Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[8], bar_empty[8], bar_tfull[4], bar_tempty[4];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float s_bias[F_MAX_CO];
    __shared__ int s_shift[F_MAX_TAPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tiles_per_img = p.m_tiles * p.n_tiles;
    const int nsteps = p.ntaps * p.ksteps;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; s++) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        for (int b = 0; b < p.acc_bufs; b++) { mbar_init(smem_u32(&bar_tfull[b]), 1); mbar_init(smem_u32(&bar_tempty[b]), F_EPI); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == F_EPI) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < F_MAX_CO; i += blockDim.x) s_bias[i] = (p.bias && i < p.Co) ? p.bias[i] : 0.0f;
    if (threadIdx.x < F_MAX_TAPS) s_shift[threadIdx.x] = p.a_shift[threadIdx.x];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;

    if (warp < F_EPI) {
        /* ===== epilogue: thread = output pixel (TMEM lane), 16 channels per TMEM load ===== */
        const int quad = warp & 3, part = warp >> 2, parts = F_EPI >> 2;
        const int r = quad * 32 + lane;
        const int n_units = p.n_tile >> 4;
        const long long plane = p.plane;
        const uint32_t acc_lane = tmem_d + ((uint32_t)(quad * 32) << 16);
        int ab = 0, aph = 0;
        for (TileIter ti(blockIdx.x, gridDim.x, tiles_per_img); ti.img < p.n_img; ti.next()) {
            mbar_wait_relaxed(smem_u32(&bar_tfull[ab]), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int mt = p.n_tiles == 1 ? ti.rem : ti.rem / p.n_tiles, n0 = (ti.rem - mt * p.n_tiles) * p.n_tile;
            const int q = mt * F_BM + r;
            const int oh = (int)__umulhi((unsigned)q, p.wp_magic), ow = q - oh * p.Wp;
            const bool valid = q < p.mflat && ow < p.Wo;
            float *o = reinterpret_cast<float *>(p.out_base + (unsigned long long)ti.img * p.slot_stride + p.out_off) + ((long long)n0 * plane + (long long)oh * p.Wo + ow);
            for (int u = part; u < n_units; u += parts) {
                uint32_t v[16];
                tmem_ld16(acc_lane + (uint32_t)(ab * p.n_tile + u * 16), v);
                if (u + parts >= n_units) { /* last read of this accumulator by this warp */
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[ab]));
                }
                const int c0 = n0 + u * 16;
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        if (c0 + j < p.Co) o[(long long)(u * 16 + j) * plane] = __fadd_rn(__uint_as_float(v[j]), s_bias[c0 + j]);
                }
            }
            if (part >= n_units) { /* narrow N tile: this warp only releases */
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[ab]));
            }
            if (++ab == p.acc_bufs) { ab = 0; aph ^= 1; }
        }
    } else if (warp == F_EPI) {
        if (lane == 0) { /* ===== MMA issuer ===== */
            const uint32_t hi_k = ((8u * p.bk) >> 4) | (1u << 14) | (p.layout << 29); /* SBO = 8 rows, version 1, swizzle */
            const uint32_t a_lo0 = (smem_base >> 4) | (1u << 16);
            const uint32_t st16 = p.stage_bytes >> 4, at16 = p.a_tile_bytes >> 4, bt16 = p.b_tile_bytes >> 4;
            const int nj = (int)(p.bk >> 5); /* K = 8 tf32 = 32 bytes per MMA */
            int s = 0, ph = 0, buf = 0, aph = 1;
            for (TileIter ti(blockIdx.x, gridDim.x, tiles_per_img); ti.img < p.n_img; ti.next()) {
                mbar_wait(smem_u32(&bar_tempty[buf]), aph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_d + (uint32_t)(buf * p.n_tile);
                for (int i = 0; i < nsteps; i++) {
                    mbar_wait(smem_u32(&bar_full[s]), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    /* stage: [A hi | A lo | B hi | B lo] (x3) or [A | B] */
                    const uint32_t a_hi_d = a_lo0 + s * st16, a_lo_d = a_hi_d + at16;
                    const uint32_t b_hi_d = a_hi_d + (p.x3 ? 2u : 1u) * at16, b_lo_d = b_hi_d + bt16;
                    for (int j = 0; j < nj; j++) {
                        umma_tf32(acc, a_hi_d + 2u * j, hi_k, b_hi_d + 2u * j, hi_k, p.idesc, (uint32_t)((i | j) != 0));
                        if (p.x3) {
                            umma_tf32(acc, a_lo_d + 2u * j, hi_k, b_hi_d + 2u * j, hi_k, p.idesc, 1u);
                            umma_tf32(acc, a_hi_d + 2u * j, hi_k, b_lo_d + 2u * j, hi_k, p.idesc, 1u);
                        }
                    }
                    umma_commit(smem_u32(&bar_empty[s]));
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
                umma_commit(smem_u32(&bar_tfull[buf]));
                if (++buf == p.acc_bufs) { buf = 0; aph ^= 1; }
            }
        }
    } else if (lane == 0) { /* ===== TMA producer ===== */
        const uint32_t tx = (p.x3 ? 2u : 1u) * (p.a_tile_bytes + (uint32_t)p.n_tile * p.bk);
        int s = 0, ph = 1;
        for (TileIter ti(blockIdx.x, gridDim.x, tiles_per_img); ti.img < p.n_img; ti.next()) {
            const int mt = p.n_tiles == 1 ? ti.rem : ti.rem / p.n_tiles, n0 = (ti.rem - mt * p.n_tiles) * p.n_tile;
            const int q0 = mt * F_BM, zc = p.img0 + ti.img;
            for (int tap = 0; tap < p.ntaps; tap++) {
                const int qa = q0 + s_shift[tap];
                for (int kb = 0; kb < p.ksteps; kb++) {
                    mbar_wait(smem_u32(&bar_empty[s]), ph);
                    const uint32_t full = smem_u32(&bar_full[s]), dst = smem_base + s * p.stage_bytes;
                    mbar_expect_tx(full, tx);
                    tma_load_3d(dst, &mapA, full, kb * (int)p.bk, qa, zc);
                    if (p.x3) {
                        tma_load_3d(dst + p.a_tile_bytes, &mapA, full, kb * (int)p.bk, qa + p.lo_rows, zc);
                        tma_load_3d(dst + 2 * p.a_tile_bytes, &mapB, full, kb * (int)p.bk, n0, 2 * tap);
                        tma_load_3d(dst + 2 * p.a_tile_bytes + p.b_tile_bytes, &mapB, full, kb * (int)p.bk, n0, 2 * tap + 1);
                    } else tma_load_3d(dst + p.a_tile_bytes, &mapB, full, kb * (int)p.bk, n0, tap);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == F_EPI) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

__device__ __forceinline__ float tf32_round(float x) { /* round to nearest, ties away (cvt.rna); the result has 13 zero low bits */
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

/* pre-pass: arena NCHW fp32 -> channel-innermost fp32 copy (pixel indexing as k_to_nhwc in conv_tc.cu): dst[pix][c] for
 * pix < npix, c < Cp (channels >= C are zero); split != 0: a second copy behind the first (and `guard` zero rows) holds lo = x - hi;
 * split == 0: the values are rounded to tf32.  32 pixels x 32 channels per block through shared memory. */
__global__ void __launch_bounds__(256) k_f32_to_kmajor(const uint8_t *src_base, unsigned long long src_stride, uint8_t *dst_base, unsigned long long dst_stride,
                                                      int C, int Cp, int H, int W, int Wp, int plane, int npix, int stride2, int pt, int pl, int split, int guard) {
    __shared__ float tile[32][33];
    const float *src = reinterpret_cast<const float *>(src_base + (unsigned long long)blockIdx.z * src_stride);
    float *dst = reinterpret_cast<float *>(dst_base + (unsigned long long)blockIdx.z * dst_stride);
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    {
        const int pix = p0 + tx;
        int ih = -1, iw = -1;
        if (pix < npix) {
            if (stride2) {
                int ph = pix / plane, r = pix - ph * plane, a = r / Wp, b = r - a * Wp;
                ih = 2 * a + (ph >> 1) - pt; iw = 2 * b + (ph & 1) - pl;
            } else {
                ih = pix / Wp; iw = pix - ih * Wp - pl;
            }
        }
        const bool inb = ih >= 0 && ih < H && iw >= 0 && iw < W;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int c = c0 + ty + 8 * k;
            tile[ty + 8 * k][tx] = (inb && c < C) ? src[((long long)c * H + ih) * W + iw] : 0.0f;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int pix = p0 + ty + 8 * k, c = c0 + tx;
        if (pix < npix + guard && c < Cp) { /* rows npix .. npix+guard-1: zeros (taps of the last tiles run past the copy; with two
                                             * copies in one tensor map the TMA's out-of-bounds zero fill only protects the second) */
            const float x = tile[tx][ty + 8 * k];
            if (split) {
                const float hi = tf32_round(x); /* stored with its 13 low bits zero: whatever the tensor core does with them, it sees hi */
                dst[(long long)pix * Cp + c] = hi;
                if (pix < npix) dst[((long long)npix + guard + pix) * Cp + c] = __fsub_rn(x, hi); /* exact; |lo| <= 2^-11 |x| */
            } else if (pix < npix) dst[(long long)pix * Cp + c] = tf32_round(x);
        }
    }
}

/* OIHW fp32 -> [tap][2][Co_pad][Cip] (split: hi = w, lo = w - trunc(w)) or [tap][Co_pad][Cip] (rounded), zero padded */
__global__ void k_repack_f32(const float *w, float *dst, int Co, int Co_pad, int Ci, int Cip, int ntaps, int split) {
    const long long per = (long long)Co_pad * Cip, total = (long long)ntaps * per;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cip);
        const long long r = i / Cip;
        const int co = (int)(r % Co_pad), tap = (int)(r / Co_pad);
        const float x = (co < Co && ci < Ci) ? w[((long long)co * Ci + ci) * ntaps + tap] : 0.0f;
        if (split) {
            const float hi = tf32_round(x);
            dst[(2ll * tap) * per + (long long)co * Cip + ci] = hi;
            dst[(2ll * tap + 1) * per + (long long)co * Cip + ci] = __fsub_rn(x, hi);
        } else dst[(long long)tap * per + (long long)co * Cip + ci] = tf32_round(x);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_f() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)sym;
    }
    return fn;
}
/* byte-typed 3-d map: dims (row bytes, rows, images / taps) */
bool make_map3b(CUtensorMap *m, void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2, uint32_t b0, uint32_t b1, CUtensorMapSwizzle sw) {
    EncodeTiledFn enc = encode_tiled_f();
    if (!enc) return false;
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {s1, s2};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled (tf32) failed: %d (dims %llu %llu %llu strides %llu %llu box %u %u)", (int)r, (unsigned long long)d0,
                       (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)s1, (unsigned long long)s2, b0, b1);
        return false;
    }
    return true;
}
int round_up_i(int x, int a) { return (x + a - 1) / a * a; }

struct F32Geom {
    bool ok = false;
    int prepass = 0; /* 1: rows padded with k-1 zero columns (stride 1); 2: 2x2 phase split (stride 2) */
    int Wp = 0, plane = 0, npix = 0, ntaps = 0, Cp = 0, guard = 0;
    size_t scratch_bytes = 0;
};
F32Geom f32_geometry(const Op &o, int mode) {
    F32Geom g;
    if (mode <= 0 || o.kind != OP_CONV_F32_NCHW || o.mode != EXEC_PARALLEL || o.xlat) return g;
    if (o.oc < 1 || o.sh != o.sw || o.sh < 1 || o.sh > 2 || o.kh < 1 || o.kh != o.kw) return g;
    if (o.oh <= 0 || o.ow <= 0 || o.ih <= 0 || o.iw <= 0 || o.ic <= 0 || round_up_i(o.oc, 16) > F_MAX_CO) return g;
    if (o.pt < 0 || o.pl < 0 || o.pl >= o.kw || o.pt >= o.kh) return g;
    if ((long long)o.oh * o.ow * o.oc * o.ic * o.kh * o.kw < (1ll << 22)) return g; /* tiny layers stay on the exact kernel */
    g.ntaps = o.kh * o.kw;
    if (g.ntaps > F_MAX_TAPS) return g;
    g.Cp = round_up_i(o.ic, 8);
    if (o.sh == 1) {
        g.prepass = 1; g.Wp = o.iw + o.kw - 1; g.plane = o.ih * g.Wp; g.npix = g.plane;
    } else {
        g.prepass = 2; g.Wp = o.ow + (o.kw - 1) / 2;
        const int rows = std::max((o.ih - 1 + o.pt) / 2 + 1, o.oh + (o.kh - 1) / 2);
        g.plane = rows * g.Wp; g.npix = 4 * g.plane;
    }
    if (o.ow > g.Wp) return g;
    if ((unsigned long long)(o.oh * (long long)g.Wp + 2 * g.Wp) * (unsigned)g.Wp >= (1ull << 32)) return g;
    /* zero rows between the hi and the lo copy: the deepest tap of the last M tile */
    const int max_shift = g.prepass == 1 ? (o.kh - 1) * g.Wp + o.kw - 1 : 3 * g.plane + ((o.kh - 1) / 2) * g.Wp + (o.kw - 1) / 2;
    g.guard = mode >= 2 ? F_BM + (g.prepass == 1 ? max_shift : ((o.kh - 1) / 2) * g.Wp + (o.kw - 1) / 2) + 8 : 0;
    (void)max_shift;
    g.scratch_bytes = (size_t)(mode >= 2 ? 2 * g.npix + g.guard : g.npix) * g.Cp * 4;
    g.ok = true;
    return g;
}

struct F32PlanImpl {
    CUtensorMap mapA, mapB;
    F32Params p;
    int prepass = 0, C = 0, Cp = 0, H = 0, W = 0, pt = 0, pl = 0, plane = 0, npix = 0, split = 0, guard = 0, sms = 148;
    const uint8_t *src_slot0 = nullptr;
    uint8_t *scratch = nullptr;
    size_t scratch_stride = 0, slot_stride = 0, smem = 0;
    float *d_wr = nullptr;
};

} // namespace

bool tf32_supported(const Op &o, int mode) { return f32_geometry(o, mode).ok; }
size_t tf32_scratch_need(const Op &o, int mode) { const F32Geom g = f32_geometry(o, mode); return g.ok ? g.scratch_bytes : 0; }

bool tf32_plan(const Op &o, const ArenaGeom &ag, int mode, uint8_t *scratch, size_t scratch_stride, TcPlan *plan) {
    const F32Geom g = f32_geometry(o, mode);
    if (!g.ok || !encode_tiled_f()) return false;
    if (o.in0 < (int64_t)ag.W || o.out < (int64_t)ag.W || o.w >= (int64_t)ag.W || o.w % 4 || g.scratch_bytes > scratch_stride) return false;
    if (o.bias >= 0 && (o.bias % 4 || o.bias + 4 * (int64_t)o.oc > (int64_t)ag.W)) return false;
    F32PlanImpl *t = new F32PlanImpl();
    F32Params &p = t->p;
    memset(&p, 0, sizeof p);
    const int x3 = mode >= 2 ? 1 : 0;
    p.x3 = x3;
    p.Co = o.oc; p.Ho = o.oh; p.Wo = o.ow; p.Wp = g.Wp; p.mflat = o.oh * g.Wp; p.plane = o.oh * o.ow;
    const int co_pad = round_up_i(o.oc, 16);
    p.n_tile = co_pad <= 256 ? co_pad : (co_pad % 256 == 0 ? 256 : 128);
    p.n_tiles = (co_pad + p.n_tile - 1) / p.n_tile;
    const int kbytes = g.Cp * 4; /* K extent of one tap in bytes */
    p.bk = kbytes % 128 == 0 ? 128 : (kbytes % 64 == 0 ? 64 : 32);
    p.ksteps = kbytes / (int)p.bk;
    p.ntaps = g.ntaps;
    p.layout = p.bk == 128 ? 2u : (p.bk == 64 ? 4u : 6u);
    p.a_tile_bytes = (uint32_t)(F_BM * p.bk);
    p.b_tile_bytes = (uint32_t)round_up_i(p.n_tile * (int)p.bk, 1024);
    p.stage_bytes = (x3 ? 2u : 1u) * (p.a_tile_bytes + p.b_tile_bytes);
    p.stages = std::max(2, std::min(8, (200 * 1024) / (int)p.stage_bytes));
    t->smem = 1024 + (size_t)p.stages * p.stage_bytes;
    p.acc_bufs = std::max(1, std::min(4, 512 / p.n_tile));
    if (p.acc_bufs < 2) p.acc_bufs = 2; /* n_tile <= 256: two buffers always fit the 512 columns */
    p.tmem_cols = 512;
    /* cute/arch/mma_sm100_desc.hpp InstrDescriptor: c = F32 (1), a = b = TF32 (2), both K-major, N >> 3, M >> 4 */
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(F_BM >> 4) << 24);
    for (int kh = 0; kh < o.kh; kh++)
        for (int kw = 0; kw < o.kw; kw++) {
            const int tap = kh * o.kw + kw;
            p.a_shift[tap] = g.prepass == 1 ? (kh - o.pt) * g.Wp + kw : ((kh & 1) * 2 + (kw & 1)) * g.plane + (kh / 2) * g.Wp + kw / 2;
        }
    p.lo_rows = g.npix + g.guard;
    p.bias = o.bias >= 0 ? reinterpret_cast<const float *>(ag.d_weights + o.bias) : nullptr;
    p.slot_stride = ag.slot_stride;
    p.out_off = o.out - (int64_t)ag.W;
    p.m_tiles = (p.mflat + F_BM - 1) / F_BM;
    p.wp_magic = (unsigned)((1ull << 32) / (unsigned)g.Wp) + 1u;
    if ((long long)p.m_tiles * p.n_tiles * ag.capacity >= (1ll << 31)) { delete t; return false; }
    t->prepass = g.prepass; t->C = o.ic; t->Cp = g.Cp; t->H = o.ih; t->W = o.iw; t->pt = o.pt; t->pl = o.pl; t->plane = g.plane; t->npix = g.npix;
    t->split = x3; t->guard = g.guard;
    t->src_slot0 = ag.d_slots + (o.in0 - (int64_t)ag.W);
    t->scratch = scratch; t->scratch_stride = scratch_stride; t->slot_stride = ag.slot_stride;
    const size_t wr_floats = (size_t)g.ntaps * (x3 ? 2 : 1) * co_pad * g.Cp;
    if (cudaMalloc(&t->d_wr, wr_floats * 4) != cudaSuccess) { delete t; return false; }
    k_repack_f32<<<256, 256>>>(reinterpret_cast<const float *>(ag.d_weights + o.w), t->d_wr, o.oc, co_pad, o.ic, g.Cp, g.ntaps, x3);
    bool ok = cudaDeviceSynchronize() == cudaSuccess;
    const CUtensorMapSwizzle ksw = p.bk == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.bk == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    ok = ok && make_map3b(&t->mapA, scratch, (uint64_t)kbytes, (uint64_t)(x3 ? 2 * g.npix + g.guard : g.npix), (uint64_t)ag.capacity, (uint64_t)kbytes, scratch_stride,
                          p.bk, F_BM, ksw);
    ok = ok && make_map3b(&t->mapB, t->d_wr, (uint64_t)kbytes, (uint64_t)co_pad, (uint64_t)g.ntaps * (x3 ? 2 : 1), (uint64_t)kbytes,
                          (uint64_t)co_pad * kbytes, p.bk, (uint32_t)p.n_tile, ksw);
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&t->sms, cudaDevAttrMultiProcessorCount, dev); if (t->sms <= 0) t->sms = 148; }
    ok = ok && cudaFuncSetAttribute((const void *)k_conv_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024) == cudaSuccess;
    if (!ok) { cudaFree(t->d_wr); delete t; return false; }
    if (getenv("MARS_TC_VERBOSE"))
        fprintf(stderr, "tf32_plan layer %d: %dx%d k%d s%d ci %d co %d | n_tile %d stages %d bk %u x3 %d smem %zu\n", o.layer, o.oh, o.ow, o.kh, o.sh, o.ic,
                o.oc, p.n_tile, p.stages, p.bk, x3, t->smem);
    plan->impl = t;
    plan->valid = true;
    plan->f32 = true;
    return true;
}

bool tf32_launch(const TcPlan &plan, uint8_t *slots_base, int first, int n, cudaStream_t s, uint64_t *launches) {
    F32PlanImpl *t = static_cast<F32PlanImpl *>(plan.impl);
    if (!t) return false;
    const uint8_t *src = t->src_slot0 + (size_t)first * t->slot_stride;
    uint8_t *scr = t->scratch + (size_t)first * t->scratch_stride;
    dim3 g((t->npix + t->guard + 31) / 32, (t->Cp + 31) / 32, n);
    k_f32_to_kmajor<<<g, 256, 0, s>>>(src, t->slot_stride, scr, t->scratch_stride, t->C, t->Cp, t->H, t->W, t->p.Wp, t->plane, t->npix, t->prepass == 2,
                                      t->pt, t->pl, t->split, t->guard);
    (*launches)++;
    F32Params p = t->p;
    p.out_base = slots_base + (size_t)first * t->slot_stride;
    p.img0 = first;
    p.n_img = n;
    const long long total = (long long)p.m_tiles * p.n_tiles * n;
    const unsigned grid = (unsigned)std::min<long long>(total, (long long)t->sms);
    k_conv_tf32<<<grid, (F_EPI + 2) * 32, t->smem, s>>>(t->mapA, t->mapB, p);
    (*launches)++;
    return cudaGetLastError() == cudaSuccess;
}

void tf32_release_one(TcPlan &pl) {
    F32PlanImpl *t = static_cast<F32PlanImpl *>(pl.impl);
    if (t) { cudaFree(t->d_wr); delete t; }
    pl.impl = nullptr;
    pl.valid = false;
    pl.f32 = false;
}

} // namespace marsb200
