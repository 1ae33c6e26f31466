/*
 * nna_layout.cuh -- the NNA-native tensor layouts the .mars format tags name, as device kernels.
 *
 *   NMHWSOIB2 weights  [N_OFP][M_IFP][KH][KW][OFP = 32][IFP = 32], 1024-byte blocks, zero padded channels
 *                      (reference include/mars.h:46-56; packer mars-compiler/src/mars_format.rs:436-470, size :472-476;
 *                      unpackers mgk-decompiler/mgk_decompiler.py:470-540 and scripts/extract_weights_nmhwsoib2.py:52-80)
 *   NDHWC32 features   [N][D_C32 = ceil(C / 32)][H][W][32], zero padded channels
 *                      (reference mars-compiler/src/mars_format.rs:478-531; byte size src/mars/mars_runtime.c:93-101)
 *
 * Pure byte permutations: bit-exact by construction, HBM bound.  The feature conversions go through a shared-memory tile
 * so that both sides move whole lines: a block transposes 32 channels x 128 pixels -- plane rows in (128 contiguous bytes
 * per channel), pixel rows out (32 contiguous bytes per pixel, 4 KiB per tile) -- and back.  The weight packers are
 * one thread per destination byte (weights are a few MB once per model).
 * A [pixel][32-byte channel group] row is also exactly the K-major A operand row the tcgen05 conv reads at bk = 32
 * (conv_tc.cu), and a 32 x 32 NMHWSOIB2 block is a K-major B tile of that kernel: these are the layouts a layer would
 * consume without the pre-pass copies.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace marsb200 {

/* dst byte index -> (n, m, h, w, ofp, ifp); source OIHW index ((o * Ci + i) * KH + h) * KW + w with o = 32 n + ofp, i = 32 m + ifp */
__global__ void __launch_bounds__(256) k_pack_nmhwsoib2(const int8_t *__restrict__ oihw, uint8_t *__restrict__ packed, int Co, int Ci, int KH, int KW,
                                                        long long total) {
    for (long long d = (long long)blockIdx.x * 256 + threadIdx.x; d < total; d += (long long)gridDim.x * 256) {
        const int ifp = (int)(d & 31), ofp = (int)((d >> 5) & 31);
        long long blk = d >> 10;
        const int w = (int)(blk % KW); blk /= KW;
        const int h = (int)(blk % KH); blk /= KH;
        const int m_ifp = (Ci + 31) >> 5;
        const int m = (int)(blk % m_ifp), n = (int)(blk / m_ifp);
        const int o = n * 32 + ofp, i = m * 32 + ifp;
        packed[d] = (o < Co && i < Ci) ? (uint8_t)oihw[(((long long)o * Ci + i) * KH + h) * KW + w] : (uint8_t)0;
    }
}

/* the inverse: one thread per OIHW byte (padding bytes of the packed blocks are ignored) */
__global__ void __launch_bounds__(256) k_unpack_nmhwsoib2(const uint8_t *__restrict__ packed, int8_t *__restrict__ oihw, int Co, int Ci, int KH, int KW,
                                                          long long total) {
    const int m_ifp = (Ci + 31) >> 5;
    for (long long s = (long long)blockIdx.x * 256 + threadIdx.x; s < total; s += (long long)gridDim.x * 256) {
        long long t = s;
        const int w = (int)(t % KW); t /= KW;
        const int h = (int)(t % KH); t /= KH;
        const int i = (int)(t % Ci), o = (int)(t / Ci);
        const long long d = ((((long long)(o >> 5) * m_ifp + (i >> 5)) * KH + h) * KW + w) * 1024 + (o & 31) * 32 + (i & 31);
        oihw[s] = (int8_t)packed[d];
    }
}

#define NNA_TILE_PX 128
/* NCHW -> NDHWC32.  grid = (pixel tiles, D_C32, N); block = 256.  Tile in shared memory: [32 channels][128 + 4 pixels]. */
__global__ void __launch_bounds__(256) k_nchw_to_ndhwc32(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int C, long long HW) {
    __shared__ __align__(16) uint8_t tile[32][NNA_TILE_PX + 4];
    const int d = blockIdx.y, n = blockIdx.z, D = gridDim.y;
    const long long p0 = (long long)blockIdx.x * NNA_TILE_PX;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npx = (int)min((long long)NNA_TILE_PX, HW - p0);
    const bool vec_in = (HW & 3) == 0 && ((uintptr_t)src & 3) == 0;
    for (int cc = warp; cc < 32; cc += 8) { /* a warp brings 128 contiguous bytes of one channel plane */
        const int c = d * 32 + cc;
        uint32_t v = 0;
        if (c < C) {
            const uint8_t *pl = src + ((long long)n * C + c) * HW + p0;
            if (vec_in) { if (4 * lane < npx) v = *reinterpret_cast<const uint32_t *>(pl + 4 * lane); }
            else
                for (int k = 0; k < 4; k++) if (4 * lane + k < npx) v |= (uint32_t)pl[4 * lane + k] << (8 * k);
        }
        *reinterpret_cast<uint32_t *>(&tile[cc][4 * lane]) = v; /* channels beyond C: zeros (mars_format.rs:500, vec![0u8; ..]) */
    }
    __syncthreads();
    /* thread t: 16 channels (half a pixel row) of pixel t >> 1; a warp writes 512 contiguous bytes */
    const int px = threadIdx.x >> 1, half = threadIdx.x & 1;
    if (px < npx) {
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int c0 = half * 16 + q * 4;
            o[q] = (uint32_t)tile[c0][px] | ((uint32_t)tile[c0 + 1][px] << 8) | ((uint32_t)tile[c0 + 2][px] << 16) | ((uint32_t)tile[c0 + 3][px] << 24);
        }
        uint8_t *out = dst + ((((long long)n * D + d) * HW + p0 + px) * 32 + half * 16);
        *reinterpret_cast<uint4 *>(out) = make_uint4(o[0], o[1], o[2], o[3]); /* 32-byte pixel rows of a cudaMalloc'ed / 16-byte aligned buffer */
    }
}

/* NCHW -> NDHWC32 without shared memory (planes of a multiple of 4 pixels, 4-byte aligned): a lane owns four adjacent pixels.  It reads
 * their word from each of the 32 channel planes of the group (a warp reads 128 contiguous bytes of one plane per instruction),
 * transposes 4 channels x 4 pixels at a time in registers (8 PRMT) and ends up with the complete 32-byte rows of its four pixels:
 * 128 contiguous output bytes per lane, 4 KiB per warp.  0.8 instructions per byte against 2.2 for the byte gather out of the tile.
 * grid = (chunks of 8 warps x 128 pixels, D_C32, N); block = 256. */
__global__ void __launch_bounds__(256) k_nchw_to_ndhwc32_reg(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int C, long long HW) {
    const int d = blockIdx.y, n = blockIdx.z, D = gridDim.y;
    const long long q = (long long)blockIdx.x * 256 + threadIdx.x; /* pixel quad */
    if (4 * q >= HW) return;
    const int c_lo = d * 32, nc = min(32, C - c_lo);
    const uint32_t *pl = reinterpret_cast<const uint32_t *>(src + ((long long)n * C + c_lo) * HW) + q;
    const long long pw = HW >> 2; /* words per plane */
    uint32_t o[4][8]; /* [pixel][4-channel word] */
#pragma unroll
    for (int g = 0; g < 8; g++) {
        const uint32_t a = 4 * g + 0 < nc ? __ldg(pl + (4 * g + 0) * pw) : 0u, b = 4 * g + 1 < nc ? __ldg(pl + (4 * g + 1) * pw) : 0u;
        const uint32_t c = 4 * g + 2 < nc ? __ldg(pl + (4 * g + 2) * pw) : 0u, e = 4 * g + 3 < nc ? __ldg(pl + (4 * g + 3) * pw) : 0u;
        const uint32_t t0 = __byte_perm(a, b, 0x5140u), t1 = __byte_perm(a, b, 0x7362u), u0 = __byte_perm(c, e, 0x5140u), u1 = __byte_perm(c, e, 0x7362u);
        o[0][g] = __byte_perm(t0, u0, 0x5410u); o[1][g] = __byte_perm(t0, u0, 0x7632u);
        o[2][g] = __byte_perm(t1, u1, 0x5410u); o[3][g] = __byte_perm(t1, u1, 0x7632u);
    }
    uint4 *out = reinterpret_cast<uint4 *>(dst + (((long long)n * D + d) * HW + 4 * q) * 32);
#pragma unroll
    for (int p = 0; p < 4; p++) {
        out[2 * p] = make_uint4(o[p][0], o[p][1], o[p][2], o[p][3]);
        out[2 * p + 1] = make_uint4(o[p][4], o[p][5], o[p][6], o[p][7]);
    }
}

/* NDHWC32 -> NCHW: the same tile the other way round */
__global__ void __launch_bounds__(256) k_ndhwc32_to_nchw(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int C, long long HW) {
    __shared__ __align__(16) uint8_t tile[32][NNA_TILE_PX + 4];
    const int d = blockIdx.y, n = blockIdx.z, D = gridDim.y;
    const long long p0 = (long long)blockIdx.x * NNA_TILE_PX;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int npx = (int)min((long long)NNA_TILE_PX, HW - p0);
    const int px = threadIdx.x >> 1, half = threadIdx.x & 1;
    if (px < npx) {
        const uint4 v = *reinterpret_cast<const uint4 *>(src + ((((long long)n * D + d) * HW + p0 + px) * 32 + half * 16));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int k = 0; k < 4; k++) tile[half * 16 + q * 4 + k][px] = (uint8_t)(w[q] >> (8 * k));
    }
    __syncthreads();
    const bool vec_out = (HW & 3) == 0 && ((uintptr_t)dst & 3) == 0;
    for (int cc = warp; cc < 32; cc += 8) {
        const int c = d * 32 + cc;
        if (c >= C) continue; /* padding channels are dropped */
        uint8_t *pl = dst + ((long long)n * C + c) * HW + p0;
        const uint32_t v = *reinterpret_cast<const uint32_t *>(&tile[cc][4 * lane]);
        if (vec_out) { if (4 * lane < npx) *reinterpret_cast<uint32_t *>(pl + 4 * lane) = v; }
        else
            for (int k = 0; k < 4; k++) if (4 * lane + k < npx) pl[4 * lane + k] = (uint8_t)(v >> (8 * k));
    }
}

} // namespace marsb200
