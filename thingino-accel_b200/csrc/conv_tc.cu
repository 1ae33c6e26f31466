/*
 * conv_tc.cu -- int8 convolution as a TMA-fed implicit GEMM on tcgen05 (sm_100a).
 *
 * Replaces the reference's conv2d_int8_mxu (src/mars/mxu_conv.c:630-670; MIPS variant
 * :144-408 built on the 64-byte S4MACSSB dot product) for hazard-free NCHW/OIHW layers.
 *
 *   GEMM view   D[M = pixels][N = out channels] = A[M][K] * B[N][K],  K = taps x in channels.
 *   A operand   1x1 convs: the NCHW activation planes themselves -- one K-row = 128 consecutive
 *               pixels of one input channel (MN-major A, legal for kind::i8), loaded by TMA
 *               straight from the arena with the 128-byte swizzle.
 *               kxk convs: a kernel tap (kh,kw) is a FLAT pixel shift, so the k-loop walks
 *               (tap, channel block) with no im2col buffer.  TMA faults ("illegal instruction",
 *               measured on B200) when the innermost box coordinate is not 16-byte aligned, so
 *               the 1-pixel shifts cannot be applied to NCHW planes; a small pre-pass kernel
 *               writes a channel-innermost (NHWC) copy instead -- rows padded with k-1 zero
 *               columns for stride 1 (the flat shift then never wraps into real pixels), a 2x2
 *               phase split for stride 2 (every tap becomes a stride-1 shift inside one phase
 *               plane) -- and A is K-major like B.  The copy is also what allows the fused
 *               outputs below to overwrite the layer's own input buffer, as the reference's
 *               work-buffer aliasing demands.
 *   B operand   weights repacked once at load to [tap][Co][Ci] (K-major), TMA + swizzle.
 *   D           int32 in TMEM (128 lanes = pixels, N columns); read back with tcgen05.ld.
 *   epilogue    + int32 bias (wrap-around), fp32 requantisation with the x86 float->int rule
 *               (SURVEY A.1), then -- when the planner fused the following SIGMOID and MUL
 *               layers -- two 256-entry tables give the sigmoid and the SiLU product of the
 *               same element; all observable tensors are stored NCHW.
 *   roles       warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected
 *               lane), warps 2..5 = epilogue (one TMEM lane quadrant each).  One output tile
 *               per CTA, two CTAs per SM so one tile's epilogue overlaps the other's mainloop.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "conv_tc.h"

namespace marsb200 {

#define TC_MAX_TAPS 36
#define TC_THREADS 320 /* warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quadrant) */
#define TC_BM 128

struct TcParams {
    int Ci, Co, Ho, Wo, Wp;  /* Wp = row pitch of the flat pixel index */
    int mflat;               /* Ho * Wp */
    int n_tile, n_tiles, bk, ksteps_per_tap, ntaps, stages;
    int tmem_cols;
    uint32_t idesc;
    uint32_t b_layout;       /* UMMA LayoutType of B: 2 = SW128, 4 = SW64, 6 = SW32 */
    int a_kmajor;            /* 0: A = NCHW planes (MN-major, SW128); 1: A = NHWC copy (K-major, swizzle = bk) */
    uint32_t a_stage_bytes, b_stage_bytes, tx_bytes;
    int a_shift[TC_MAX_TAPS];
    int a_cbase[TC_MAX_TAPS];
    const int32_t *bias;     /* device pointer or null */
    float cs;
    int post_relu;
    uint8_t *out_base;       /* slot 0 of the launch */
    unsigned long long slot_stride;
    long long out_y, out_s, out_z; /* slot-relative byte offsets, -1 = not stored */
    const uint8_t *lut_s, *lut_z;  /* 256-byte tables or null */
    int img0;                /* first image (TMA coordinate of the slot dimension) */
    int vec_store;           /* tile rows are consecutive, 16-byte aligned output pixels */
    int m_tiles, tiles_per_cta;
};

/* ---- PTX wrappers ------------------------------------------------------------- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

/* UMMA shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp layout):
 * [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout type */
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}

/* reference src/mars/mxu_conv.c:663-666 with the x86 cvttss2si rule: NaN and |v| >= 2^31 become
 * INT_MIN, which then clamps to -128 (cvt.rzi.sat alone would give +127 / 0) */
__device__ __forceinline__ int requant_i8(int32_t acc, float cs) {
    const float scaled = __fmul_rn(__int2float_rn(acc), cs);
    const float biased = __fadd_rn(scaled, copysignf(0.5f, scaled)); /* scaled >= 0 ? +0.5 : -0.5; -0.0f cannot occur (int * positive or any cs: sign of zero only matters when scaled == 0, where +-0.5 both truncate to 0) */
    int r;
    asm("cvt.rzi.sat.s8.f32 %0, %1;" : "=r"(r) : "f"(biased)); /* truncate + clamp to [-128,127]; NaN -> 0 */
    return (biased < 2147483648.0f) ? r : -128;
}

/* ---- the kernel -------------------------------------------------------------- */
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

/* Each CTA walks `tiles_per_cta` consecutive 128-pixel tiles of one (image, N tile): the TMA ring,
 * the barriers and the TMEM allocation are set up once, and the accumulator is double buffered in
 * TMEM so the epilogue of tile t overlaps the MMAs of tile t+1. */
__global__ void __launch_bounds__(TC_THREADS, 2)
k_conv_tc(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[8], bar_empty[8], bar_tmem_full[2], bar_tmem_empty[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) int32_t s_bias[256];
    __shared__ uint8_t s_lut[512];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_blk = blockIdx.y, img = blockIdx.z;
    const int tile0 = blockIdx.x * p.tiles_per_cta;
    const int ntiles = min(p.tiles_per_cta, p.m_tiles - tile0);
    const int n0 = n_blk * p.n_tile;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base, b_base = smem_base + p.stages * p.a_stage_bytes;
    uint8_t *stage_out = smem_raw + (smem_base - smem_u32(smem_raw)) + p.stages * (p.a_stage_bytes + p.b_stage_bytes);
    const int nsteps = p.ntaps * p.ksteps_per_tap;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; s++) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        for (int b = 0; b < 2; b++) { mbar_init(smem_u32(&bar_tmem_full[b]), 1); mbar_init(smem_u32(&bar_tmem_empty[b]), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) { /* TMEM allocation is a warp-wide operation; this warp also frees it */
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_bias[i] = (p.bias && i < p.n_tile && n0 + i < p.Co) ? p.bias[n0 + i] : 0;
    if (p.lut_s) for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = p.lut_s[i];
    if (p.lut_z) for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[256 + i] = p.lut_z[i];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) { /* ===== TMA producer ===== */
            const uint32_t tx = p.tx_bytes;
            int it = 0;
            for (int tl = 0; tl < ntiles; tl++) {
                const int q0 = (tile0 + tl) * TC_BM;
                for (int i = 0; i < nsteps; i++, it++) {
                    const int s = it % p.stages, ph = (it / p.stages) & 1;
                    const int tap = i / p.ksteps_per_tap, kb = i - tap * p.ksteps_per_tap;
                    mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
                    const uint32_t full = smem_u32(&bar_full[s]);
                    mbar_expect_tx(full, tx);
                    if (p.a_kmajor) tma_load_3d(a_base + s * p.a_stage_bytes, &mapA, full, kb * p.bk, q0 + p.a_shift[tap], p.img0 + img);
                    else tma_load_3d(a_base + s * p.a_stage_bytes, &mapA, full, q0, kb * p.bk, p.img0 + img);
                    tma_load_3d(b_base + s * p.b_stage_bytes, &mapB, full, kb * p.bk, n0, tap);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) { /* ===== MMA issuer ===== */
            const uint32_t b_sbo = 8u * (uint32_t)p.bk;
            int it = 0;
            for (int tl = 0; tl < ntiles; tl++) {
                const int buf = tl & 1;
                mbar_wait(smem_u32(&bar_tmem_empty[buf]), ((tl >> 1) & 1) ^ 1); /* epilogue drained this accumulator */
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_d + (uint32_t)(buf * p.n_tile);
                for (int i = 0; i < nsteps; i++, it++) {
                    const int s = it % p.stages, ph = (it / p.stages) & 1;
                    mbar_wait(smem_u32(&bar_full[s]), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = a_base + s * p.a_stage_bytes, b_addr = b_base + s * p.b_stage_bytes;
                    for (int j = 0; j < p.bk / 32; j++) {
                        /* A MN-major, 128B swizzle: 32 K-rows of 128 bytes = 4 atoms of 8 rows, 1024 B apart;
                         * A K-major: same layout rules as B */
                        const uint64_t da = p.a_kmajor ? umma_desc(a_addr + j * 32u, 16u, b_sbo, p.b_layout)
                                                       : umma_desc(a_addr + j * 4096u, 0, 1024u, 2u);
                        /* B: K-major, swizzle = bk bytes: rows of bk bytes, 8-row groups 8*bk apart; advance 32 B per MMA */
                        const uint64_t db = umma_desc(b_addr + j * 32u, 16u, b_sbo, p.b_layout);
                        umma_i8(acc, da, db, p.idesc, (uint32_t)((i | j) != 0));
                    }
                    umma_commit(smem_u32(&bar_empty[s])); /* frees the stage when these MMAs retire */
                }
                umma_commit(smem_u32(&bar_tmem_full[buf]));
            }
        }
    } else { /* ===== epilogue: TMEM -> registers -> requant (+ SiLU tables) -> NCHW stores ===== */
        /* Two warps share each TMEM lane quadrant and split the accumulator columns in units of 16.
         * Valid tile rows (pad columns of the flat pixel index skipped) are CONSECUTIVE output pixels,
         * so each unit is transposed through shared memory -- staged at its final 16-byte phase --
         * and stored as 16-byte vectors along the pixel axis, bytes only at the ragged ends. */
        const int ew = warp - 2, grp = ew >> 2, quad = warp & 3;
        const int r = quad * 32 + lane;       /* accumulator row = pixel of the tile */
        const int et = (ew & 3) * 32 + lane;  /* 0..127 inside the group */
        const int plane = p.Ho * p.Wo;
        uint8_t *img_base = p.out_base + (unsigned long long)img * p.slot_stride;
        uint8_t *stg = stage_out + grp * (3 * 16 * 144);
        const long long outs[3] = {p.out_y, p.out_s, p.out_z};
        const int n_units = p.n_tile >> 4;
        const uint32_t bar_id = 1 + grp;
        for (int tl = 0; tl < ntiles; tl++) {
            const int buf = tl & 1;
            const int q0 = (tile0 + tl) * TC_BM;
            const int q = q0 + r;
            const int oh = q / p.Wp, ow = q - oh * p.Wp;
            const bool valid = q < p.mflat && ow < p.Wo;
            /* output pixel index of flat q = number of valid flat indices before it */
            const int oh0 = q0 / p.Wp, ow0 = q0 - oh0 * p.Wp;
            const int pix_first = oh0 * p.Wo + min(ow0, p.Wo);
            const int qe = min(q0 + TC_BM, p.mflat), ohe = qe / p.Wp, owe = qe - ohe * p.Wp;
            const int pix_end = ohe * p.Wo + min(owe, p.Wo);
            const int mis = p.vec_store ? (pix_first & 15) : 0;
            const int sidx = oh * p.Wo + ow - pix_first + mis; /* staging column of this row */
            const int len = pix_end - pix_first;
            mbar_wait(smem_u32(&bar_tmem_full[buf]), (tl >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * p.n_tile);
            if (grp >= n_units) { /* nothing to read for this group: release the accumulator at once */
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bar_tmem_empty[buf]));
            }
            for (int u = grp; u < n_units; u += 2) {
                uint32_t v[16];
                tmem_ld16(acc + (uint32_t)(u * 16), v);
                if (u + 2 >= n_units) { /* last read of this accumulator by this warp: hand it back */
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bar_tmem_empty[buf]));
                }
                if (valid) {
                    const int4 *bv = reinterpret_cast<const int4 *>(s_bias + u * 16);
#pragma unroll
                    for (int j4 = 0; j4 < 4; j4++) {
                        const int4 b4 = bv[j4];
                        const int bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int j = j4 * 4 + k;
                            int y = requant_i8((int32_t)(v[j] + (uint32_t)bb[k]), p.cs);
                            if (p.post_relu) y = max(y, 0);
                            if (p.out_y >= 0) stg[j * 144 + sidx] = (uint8_t)y;
                            if (p.out_s >= 0) stg[16 * 144 + j * 144 + sidx] = s_lut[y + 128];
                            if (p.out_z >= 0) stg[32 * 144 + j * 144 + sidx] = s_lut[256 + y + 128];
                        }
                    }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
                for (int t = 0; t < 3; t++) {
                    if (outs[t] < 0) continue;
                    for (int slot = et; slot < 16 * 9; slot += 128) {
                        const int jl = slot / 9, ch = slot - jl * 9;
                        const int co = n0 + u * 16 + jl;
                        const int lo = max(ch * 16, mis), hi = min(ch * 16 + 16, mis + len);
                        if (co >= p.Co || lo >= hi) continue;
                        const uint8_t *src = stg + t * (16 * 144) + jl * 144;
                        uint8_t *dst = img_base + outs[t] + (long long)co * plane + (pix_first - mis);
                        if (p.vec_store && hi - lo == 16) *reinterpret_cast<uint4 *>(dst + ch * 16) = *reinterpret_cast<const uint4 *>(src + ch * 16);
                        else for (int b2 = lo; b2 < hi; b2++) dst[b2] = src[b2];
                    }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

/* ---- pre-pass: arena NCHW -> channel-innermost copy with zero padding --------------- */
/* dst[pix][c] for pix < npix; source pixel of `pix`:
 *   stride 1: ih = pix / Wp, iw = pix % Wp - pl                          (rows padded to Wp)
 *   stride 2: ph = pix / plane, a = (pix % plane) / Wp, b = .. % Wp, ih = 2a + ph/2 - pt, iw = 2b + ph%2 - pl
 * 32 pixels x 32 channels per block through shared memory: coalesced reads along pixels,
 * coalesced writes along channels. */
__global__ void __launch_bounds__(256) k_to_nhwc(const uint8_t *src_base, unsigned long long src_stride, uint8_t *dst_base,
                                                 unsigned long long dst_stride, int C, int H, int W, int Wp, int plane, int npix,
                                                 int stride2, int pt, int pl) {
    __shared__ uint8_t tile[32][33];
    const uint8_t *src = src_base + (unsigned long long)blockIdx.z * src_stride;
    uint8_t *dst = dst_base + (unsigned long long)blockIdx.z * dst_stride;
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
            tile[ty + 8 * k][tx] = (inb && c < C) ? src[((long long)c * H + ih) * W + iw] : (uint8_t)0;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int pix = p0 + ty + 8 * k, c = c0 + tx;
        if (pix < npix && c < C) dst[(long long)pix * C + c] = tile[tx][ty + 8 * k];
    }
}
/* ---- pre-pass for small-Ci convs (the 6x6 stride-2 stem, Ci = 3): explicit im2col ------
 * dst[q][k], q = oh*Wo + ow, k = (ci*KH + y)*KW + x (the OIHW row order, so the weights need no
 * permutation), zero for k >= Kt and for taps outside the input.  One thread writes 4 k's. */
__global__ void __launch_bounds__(256) k_im2col(const uint8_t *src_base, unsigned long long src_stride, uint8_t *dst_base,
                                                unsigned long long dst_stride, int C, int H, int W, int Ho, int Wo, int KH, int KW,
                                                int S, int pt, int pl, int Kt, int Kp) {
    __shared__ int s_off[256]; /* per k: packed (ci, y, x) */
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
        int ci = k / (KH * KW), r = k - ci * KH * KW, y = r / KW, x = r - y * KW;
        s_off[k] = k < Kt ? (ci << 16) | (y << 8) | x : -1;
    }
    __syncthreads();
    const uint8_t *src = src_base + (unsigned long long)blockIdx.z * src_stride;
    uint8_t *dst = dst_base + (unsigned long long)blockIdx.z * dst_stride;
    const int words = Kp >> 2;
    const long long total = (long long)Ho * Wo * words;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(t / words), wk = (int)(t - (long long)q * words);
        const int oh = q / Wo, ow = q - oh * Wo;
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int code = s_off[wk * 4 + b];
            if (code >= 0) {
                const int ci = code >> 16, ih = oh * S - pt + ((code >> 8) & 0xFF), iw = ow * S - pl + (code & 0xFF);
                if (ih >= 0 && ih < H && iw >= 0 && iw < W) word |= (uint32_t)src[((long long)ci * H + ih) * W + iw] << (8 * b);
            }
        }
        reinterpret_cast<uint32_t *>(dst)[t] = word;
    }
}
/* OIHW rows (Kt bytes) -> [Co_pad][Kp], zero padded */
__global__ void k_repack_rows(const int8_t *w, int8_t *dst, int Co, int Co_pad, int Kt, int Kp) {
    long long total = (long long)Co_pad * Kp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int k = (int)(i % Kp), co = (int)(i / Kp);
        dst[i] = (co < Co && k < Kt) ? w[(long long)co * Kt + k] : (int8_t)0;
    }
}
/* OIHW -> [tap][Co_pad][Ci], rows beyond Co zero */
__global__ void k_repack_weights(const int8_t *w, int8_t *dst, int Co, int Co_pad, int Ci, int ntaps) {
    long long total = (long long)ntaps * Co_pad * Ci;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ci = (int)(i % Ci);
        long long r = i / Ci;
        int co = (int)(r % Co_pad), tap = (int)(r / Co_pad);
        dst[i] = co < Co ? w[((long long)co * Ci + ci) * ntaps + tap] : (int8_t)0;
    }
}

/* ---- host side ------------------------------------------------------------------ */
struct TcPlanImpl {
    CUtensorMap mapA, mapB;
    TcParams p;
    int prepass = 0;
    int C = 0, H = 0, W = 0, pt = 0, pl = 0, plane = 0, npix = 0, Kp = 0, KH = 0, KW = 0, S = 1, Ho = 0, Wo = 0;
    const uint8_t *src_slot0 = nullptr; /* input tensor in slot 0 */
    uint8_t *scratch = nullptr;
    size_t scratch_stride = 0, slot_stride = 0;
    int8_t *d_wr = nullptr;
    int m_tiles = 0;
    size_t smem = 0;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)sym;
    }
    return fn;
}

static bool make_map3(CUtensorMap *m, void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2, uint32_t b0,
                      uint32_t b1, CUtensorMapSwizzle sw) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {s1, s2};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed: %d (dims %llu %llu %llu strides %llu %llu box %u %u)", (int)r,
                       (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)s1,
                       (unsigned long long)s2, b0, b1);
        return false;
    }
    return true;
}


static int round_up(int x, int a) { return (x + a - 1) / a * a; }

/* geometry shared by tc_scratch_need and tc_plan */
struct TcGeom {
    bool ok = false;
    int prepass = 0; /* 0: A from the arena (1x1); 1: NHWC copy, rows padded (stride 1); 2: NHWC 2x2 phase split (stride 2);
                        3: explicit im2col rows of Kp bytes (small Ci) */
    int Wp = 0, plane = 0, npix = 0, ntaps = 0, Kp = 0;
    size_t scratch_bytes = 0;
};
static TcGeom tc_geometry(const Op &o) {
    TcGeom g;
    if (o.kind != OP_CONV_I8_NCHW || o.mode != EXEC_PARALLEL || o.xlat) return g;
    if (o.oc < 16 || o.sh != o.sw || o.sh < 1 || o.kh < 1 || o.kw < 1) return g;
    if (o.oh <= 0 || o.ow <= 0 || o.ih <= 0 || o.iw <= 0 || o.ic <= 0) return g;
    if (o.ic < 32 || o.ic % 32 || o.kh != o.kw) {
        /* small / odd channel counts: im2col rows, if one row stays small */
        const int Kt = o.ic * o.kh * o.kw;
        if (Kt > 256 || o.kh > 255 || o.kw > 255 || (long long)o.oh * o.ow < 4096) return g;
        g.prepass = 3; g.Kp = round_up(Kt, 32); g.ntaps = 1; g.Wp = o.ow; g.npix = o.oh * o.ow;
        g.scratch_bytes = (size_t)g.npix * g.Kp;
        g.ok = true;
        return g;
    }
    g.ntaps = o.kh * o.kw;
    if (g.ntaps > TC_MAX_TAPS) return g;
    if (o.kh == 1 && o.sh == 1 && o.pt == 0 && o.pl == 0 && o.oh <= o.ih && o.ow == o.iw && ((long long)o.ih * o.iw) % 16 == 0) {
        g.prepass = 0; g.Wp = o.iw;
    } else if (o.sh == 1) {
        if (o.pl >= o.kw || o.pt >= o.kh) return g;
        g.prepass = 1; g.Wp = o.iw + o.kw - 1; g.plane = o.ih * g.Wp; g.npix = g.plane;
    } else if (o.sh == 2) {
        if (o.pl >= o.kw || o.pt >= o.kh) return g;
        /* phase plane: rows a = 0 .. , pitch Wp >= ow + (k-1)/2; enough zero rows below that the
         * deepest tap of the last output row stays inside its own plane */
        g.prepass = 2; g.Wp = o.ow + (o.kw - 1) / 2;
        const int rows = std::max((o.ih - 1 + o.pt) / 2 + 1, o.oh + (o.kh - 1) / 2);
        g.plane = rows * g.Wp; g.npix = 4 * g.plane;
    } else return g;
    if (o.ow > g.Wp) return g;
    g.scratch_bytes = g.prepass ? (size_t)g.npix * o.ic : 0;
    g.ok = true;
    return g;
}

size_t tc_scratch_need(const Op &o) {
    TcGeom g = tc_geometry(o);
    return g.ok ? g.scratch_bytes : 0;
}
bool tc_supported(const Op &o) {
    if (!tc_geometry(o).ok) return false;
    const int co_pad = round_up(o.oc, 16);
    return co_pad <= 256 || co_pad % 128 == 0;
}
int tc_n_tiles(int oc) {
    const int co_pad = round_up(oc, 16);
    const int nt = co_pad <= 256 ? co_pad : (co_pad % 256 == 0 ? 256 : 128);
    return (co_pad + nt - 1) / nt;
}
bool tc_uses_copy(const Op &o) { return tc_geometry(o).prepass != 0; }

bool tc_plan(const Op &o, const ArenaGeom &ag, const uint8_t *d_cpool, uint8_t *scratch, size_t scratch_stride, TcPlan *plan) {
    TcGeom g = tc_geometry(o);
    if (!g.ok || !encode_tiled()) return false;
    if (o.in0 < (int64_t)ag.W || o.out < (int64_t)ag.W || o.w >= (int64_t)ag.W) return false;
    if (g.scratch_bytes > scratch_stride) return false;
    TcPlanImpl *t = new TcPlanImpl();
    TcParams &p = t->p;
    memset(&p, 0, sizeof p);
    const int ci_eff = g.prepass == 3 ? g.Kp : o.ic; /* K extent of one tap */
    p.Ci = ci_eff; p.Co = o.oc; p.Ho = o.oh; p.Wo = o.ow; p.Wp = g.Wp; p.mflat = o.oh * g.Wp;
    const int co_pad = round_up(o.oc, 16);
    p.n_tile = co_pad <= 256 ? co_pad : (co_pad % 256 == 0 ? 256 : 128);
    if (co_pad > 256 && co_pad % 128) { delete t; return false; }
    p.n_tiles = (co_pad + p.n_tile - 1) / p.n_tile;
    p.bk = ci_eff % 64 == 0 ? 64 : 32;
    p.ksteps_per_tap = ci_eff / p.bk;
    p.ntaps = g.ntaps;
    p.a_stage_bytes = (uint32_t)(TC_BM * p.bk);
    p.b_stage_bytes = (uint32_t)round_up(p.n_tile * p.bk, 1024);
    p.tx_bytes = p.a_stage_bytes + (uint32_t)(p.n_tile * p.bk);
    const int stage_bytes = (int)(p.a_stage_bytes + p.b_stage_bytes);
    p.stages = std::max(2, std::min(8, (94 * 1024) / stage_bytes));
    p.tmem_cols = 32;
    while (p.tmem_cols < 2 * p.n_tile) p.tmem_cols <<= 1; /* two accumulators */
    /* cute/arch/mma_sm100_desc.hpp InstrDescriptor: c=S32, a=b=signed 8 bit, A MN-major, B K-major, N>>3, M>>4 */
    p.idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    p.b_layout = p.bk == 64 ? 4u : 6u;
    p.a_kmajor = g.prepass != 0;
    if (!p.a_kmajor) p.idesc |= 1u << 15; /* A MN-major */
    for (int kh = 0; kh < (g.prepass == 3 ? 1 : o.kh); kh++)
        for (int kw = 0; kw < (g.prepass == 3 ? 1 : o.kw); kw++) {
            const int tap = kh * o.kw + kw;
            if (g.prepass == 0 || g.prepass == 3) { if (tap < TC_MAX_TAPS) p.a_shift[tap] = 0; }
            else if (g.prepass == 1) p.a_shift[tap] = (kh - o.pt) * g.Wp + kw;
            else p.a_shift[tap] = ((kh & 1) * 2 + (kw & 1)) * g.plane + (kh / 2) * g.Wp + kw / 2;
            if (tap < TC_MAX_TAPS) p.a_cbase[tap] = 0;
        }
    p.bias = o.bias >= 0 ? reinterpret_cast<const int32_t *>(ag.d_weights + o.bias) : nullptr;
    if (o.bias >= 0 && (o.bias % 4 || o.bias + 4 * (int64_t)o.oc > (int64_t)ag.W)) { delete t; return false; }
    p.cs = o.f0;
    p.post_relu = o.post_relu;
    p.vec_store = (((long long)o.oh * o.ow) % 16 == 0 && (ag.slot_stride % 16) == 0) ? 1 : 0;
    p.slot_stride = ag.slot_stride;
    p.out_y = o.store_y ? o.out - (int64_t)ag.W : -1;
    p.out_s = o.out_s >= 0 ? o.out_s - (int64_t)ag.W : -1;
    p.out_z = o.out_z >= 0 ? o.out_z - (int64_t)ag.W : -1;
    p.lut_s = o.lut_s >= 0 ? d_cpool + o.lut_s : nullptr;
    p.lut_z = o.lut_z >= 0 ? d_cpool + o.lut_z : nullptr;
    t->prepass = g.prepass; t->C = o.ic; t->H = o.ih; t->W = o.iw; t->pt = o.pt; t->pl = o.pl;
    t->plane = g.plane; t->npix = g.npix; t->Kp = g.Kp; t->KH = o.kh; t->KW = o.kw; t->S = o.sh; t->Ho = o.oh; t->Wo = o.ow;
    t->src_slot0 = ag.d_slots + (o.in0 - (int64_t)ag.W);
    t->scratch = scratch; t->scratch_stride = scratch_stride; t->slot_stride = ag.slot_stride;
    t->m_tiles = (p.mflat + TC_BM - 1) / TC_BM;
    p.m_tiles = t->m_tiles;
    t->smem = 1024 + (size_t)p.stages * stage_bytes + 2 * 3 * 16 * 144;
    if (p.tmem_cols > 256) t->smem = std::max<size_t>(t->smem, 120 * 1024); /* one CTA per SM: it owns all of TMEM */

    /* weights: [tap][co_pad][Ci] K-major */
    const size_t wr_bytes = (size_t)g.ntaps * co_pad * ci_eff;
    if (cudaMalloc(&t->d_wr, wr_bytes) != cudaSuccess) { delete t; return false; }
    if (g.prepass == 3)
        k_repack_rows<<<256, 256>>>(reinterpret_cast<const int8_t *>(ag.d_weights + o.w), t->d_wr, o.oc, co_pad, o.ic * o.kh * o.kw, g.Kp);
    else
        k_repack_weights<<<256, 256>>>(reinterpret_cast<const int8_t *>(ag.d_weights + o.w), t->d_wr, o.oc, co_pad, o.ic, g.ntaps);
    if (cudaDeviceSynchronize() != cudaSuccess) { cudaFree(t->d_wr); delete t; return false; }

    bool ok;
    if (g.prepass == 0)
        ok = make_map3(&t->mapA, (void *)t->src_slot0, (uint64_t)o.ih * o.iw, (uint64_t)o.ic, (uint64_t)ag.capacity,
                       (uint64_t)o.ih * o.iw, ag.slot_stride, TC_BM, (uint32_t)p.bk, CU_TENSOR_MAP_SWIZZLE_128B);
    else /* NHWC copy: dims (C, pixels, images), K-major box {bk, 128} */
        ok = make_map3(&t->mapA, scratch, (uint64_t)ci_eff, (uint64_t)g.npix, (uint64_t)ag.capacity, (uint64_t)ci_eff,
                       scratch_stride, (uint32_t)p.bk, TC_BM, p.bk == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    ok = ok && make_map3(&t->mapB, t->d_wr, (uint64_t)ci_eff, (uint64_t)co_pad, (uint64_t)g.ntaps, (uint64_t)ci_eff,
                         (uint64_t)co_pad * ci_eff, (uint32_t)p.bk, (uint32_t)p.n_tile,
                         p.bk == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    if (ok) ok = cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess;
    if (!ok) { cudaFree(t->d_wr); delete t; return false; }
    plan->impl = t;
    plan->valid = true;
    return true;
}

bool tc_launch(const TcPlan &plan, uint8_t *slots_base, int first, int n, cudaStream_t s, uint64_t *launches) {
    TcPlanImpl *t = static_cast<TcPlanImpl *>(plan.impl);
    if (!t) return false;
    const uint8_t *src = t->src_slot0 + (size_t)first * t->slot_stride;
    uint8_t *scr = t->scratch + (size_t)first * t->scratch_stride;
    if (t->prepass == 3) {
        const long long total = (long long)t->npix * (t->Kp / 4);
        dim3 g((unsigned)std::min<long long>((total + 255) / 256, 148 * 32), 1, n);
        k_im2col<<<g, 256, 0, s>>>(src, t->slot_stride, scr, t->scratch_stride, t->C, t->H, t->W, t->Ho, t->Wo, t->KH, t->KW, t->S,
                                   t->pt, t->pl, t->C * t->KH * t->KW, t->Kp);
        (*launches)++;
    } else if (t->prepass) {
        dim3 g((t->npix + 31) / 32, (t->C + 31) / 32, n);
        k_to_nhwc<<<g, 256, 0, s>>>(src, t->slot_stride, scr, t->scratch_stride, t->C, t->H, t->W, t->p.Wp, t->plane, t->npix,
                                    t->prepass == 2, t->pt, t->pl);
        (*launches)++;
    }
    TcParams p = t->p;
    p.out_base = slots_base + (size_t)first * t->slot_stride;
    p.img0 = first;
    /* enough CTAs for a few waves of 2 per SM, each amortising its setup over several tiles */
    const long long total_tiles = (long long)t->m_tiles * p.n_tiles * n;
    int tpc = (int)std::min<long long>(16, std::max<long long>(1, total_tiles / (148 * 2 * 4)));
    tpc = std::min(tpc, t->m_tiles);
    p.tiles_per_cta = tpc;
    dim3 grid((t->m_tiles + tpc - 1) / tpc, p.n_tiles, n);
    k_conv_tc<<<grid, TC_THREADS, t->smem, s>>>(t->mapA, t->mapB, p);
    (*launches)++;
    return cudaGetLastError() == cudaSuccess;
}

void tc_release(std::vector<TcPlan> &plans) {
    for (auto &pl : plans) {
        TcPlanImpl *t = static_cast<TcPlanImpl *>(pl.impl);
        if (t) { cudaFree(t->d_wr); delete t; }
        pl.impl = nullptr;
        pl.valid = false;
    }
    plans.clear();
}

} // namespace marsb200
