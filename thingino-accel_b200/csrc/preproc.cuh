/*
 * preproc.cuh -- letterbox pre-processing on the GPU: the step in front of mars_run in the reference's
 * callers, load_image() of src/mars/mars_yolo_test.c:40-77 minus the file decode:
 *     scale = fminf(tw/ow, th/oh); nw = (int)(ow*scale); nh = (int)(oh*scale); px = (tw-nw)/2; py = (th-nh)/2;
 *     stbir_resize_uint8(img, ow, oh, 0, rsz, nw, nh, 0, 3);  out = -17 everywhere, (int8)(rsz - 128) inside, NCHW or NHWC.
 *
 * stbir_resize_uint8 is the vendored include/stb/stb_image_resize.h (v0.9x): linear colour space, clamp edges, default
 * filters -- Catmull-Rom when an axis is enlarged (ratio > 1), Mitchell otherwise -- horizontal pass first, float
 * intermediate, every product and sum rounded separately.  Whatever route it takes (gather when enlarging, scatter when
 * shrinking), an output sample is a sum of (input sample x coefficient) terms added onto 0.0f in ascending source order;
 * build_resize_taps() restates its coefficient construction (stb_image_resize.h:1008-1231) on the host and hands the
 * kernels one tap list per output index, in that order, so the device sums are bit-identical:
 *     k_pre_horizontal    : H[f][r][x][c] = sum_t (u8 / 255.0f) * wh[t]            (stb :1243-1282, :1441-1652)
 *     k_pre_vertical_pack : v = sum_t H[f][src_t][x][c] * wv[t]; u8 = (int)((double)(sat(v) * 255.0f) + 0.5);
 *                           out = (int8)(u8 - 128), border -17  -- or RGBA u8, border 114 (stb :1692-1760, :1866-2061)
 */
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <vector>

namespace marsb200 {

/* per output index o: taps [start[o], start[o+1]) = (source index already clamped to the image, coefficient) */
struct ResizeTaps {
    std::vector<int> start, src;
    std::vector<float> w;
};

/* the two default kernels (stb :810-836); x86-64 evaluates every operation in float */
static inline float pre_catmullrom(float x) {
    x = fabsf(x);
    if (x < 1.0f) return 1 - x * x * (2.5f - 1.5f * x);
    if (x < 2.0f) return 2 - x * (4 + x * (0.5f * x - 2.5f));
    return 0.0f;
}
static inline float pre_mitchell(float x) {
    x = fabsf(x);
    if (x < 1.0f) return (16 + x * x * (21 * x - 36)) / 18;
    if (x < 2.0f) return (32 + x * (-60 + x * (36 - 7 * x))) / 18;
    return 0.0f;
}

/* Coefficients live in one flat array of `width` floats per contributor, exactly as in the reference: a contributor with
 * width + 1 candidates spills its last (zero) candidate into its neighbour's first cell, which the neighbour then
 * overwrites -- kept, so that every read sees what the reference's read sees. */
static inline void build_resize_taps(int in_size, int out_size, ResizeTaps *t) {
    const float scale = ((float)out_size / in_size) / (1.0f - 0.0f);
    const float shift = 0.0f * out_size / (1.0f - 0.0f);
    const bool up = scale > 1;
    const float support = 2.0f; /* both default filters */
    const int width = (int)ceil(support * 2);
    t->start.assign(1, 0);
    t->src.clear();
    t->w.clear();
    auto clampi = [&](int v) { return v < 0 ? 0 : (v >= in_size ? in_size - 1 : v); };
    if (up) {
        /* one contributor per OUTPUT index: which input samples it gathers (stb :1008-1021, :1037-1085) */
        const int nc = out_size;
        std::vector<float> coef((size_t)nc * width + 16, 0.0f);
        std::vector<int> n0(nc), n1(nc);
        const float radius = support * scale;
        for (int n = 0; n < nc; n++) {
            const float center = (float)n + 0.5f;
            const float lo = center - radius, hi = center + radius;
            const float in_lo = (lo + shift) / scale, in_hi = (hi + shift) / scale;
            const float in_center = (center + shift) / scale;
            int first = (int)floor(in_lo + 0.5), last = (int)floor(in_hi - 0.5);
            float *g = &coef[(size_t)n * width];
            float total = 0;
            n0[n] = first; n1[n] = last;
            for (int i = 0; i <= last - first; i++) {
                const float pc = (float)(i + first) + 0.5f;
                g[i] = pre_catmullrom(in_center - pc);
                if (i == 0 && !g[i]) { /* a leading zero is dropped and the window slides */
                    n0[n] = ++first;
                    i--;
                    continue;
                }
                total += g[i];
            }
            const float fs = 1 / total;
            for (int i = 0; i <= last - first; i++) g[i] *= fs;
            for (int i = last - first; i >= 0; i--) {
                if (g[i]) break;
                n1[n] = n0[n] + i - 1;
            }
        }
        for (int o = 0; o < out_size; o++) {
            for (int k = n0[o]; k <= n1[o]; k++) {
                t->src.push_back(clampi(k));
                t->w.push_back(coef[(size_t)o * width + (k - n0[o])]);
            }
            t->start.push_back((int)t->src.size());
        }
    } else {
        /* one contributor per INPUT index (with a margin of clamped virtual samples either side): which output samples it
         * feeds (stb :1023-1035, :1087-1115), then the per-output normalisation and clean-up (stb :1117-1190) */
        const int pixel_width = (int)ceil(support * 2 / scale);
        const int margin = pixel_width / 2;
        const int nc = in_size + margin * 2;
        std::vector<float> coef((size_t)nc * width + 16, 0.0f);
        coef[(size_t)nc * width + 8] = 1.0f; /* stops a zero-skip that would otherwise run off the end */
        std::vector<int> n0(nc), n1(nc);
        const float radius = support / scale;
        for (int n = 0; n < nc; n++) {
            const int na = n - margin;
            const float center = (float)na + 0.5f;
            const float lo = center - radius, hi = center + radius;
            const float out_lo = lo * scale - shift, out_hi = hi * scale - shift;
            const float out_center = center * scale - shift;
            const int first = (int)floor(out_lo + 0.5), last = (int)floor(out_hi - 0.5);
            float *g = &coef[(size_t)n * width];
            n0[n] = first; n1[n] = last;
            for (int i = 0; i <= last - first; i++) {
                const float pc = (float)(i + first) + 0.5f;
                const float x = pc - out_center;
                g[i] = pre_mitchell(x) * scale;
            }
            for (int i = last - first; i >= 0; i--) {
                if (g[i]) break;
                n1[n] = n0[n] + i - 1;
            }
        }
        auto at = [&](int j, int c) -> float & { return coef[(size_t)width * j + c]; };
        for (int i = 0; i < out_size; i++) {
            float total = 0;
            for (int j = 0; j < nc; j++) {
                if (i >= n0[j] && i <= n1[j]) total += at(j, i - n0[j]);
                else if (i < n0[j]) break;
            }
            const float s = 1 / total;
            for (int j = 0; j < nc; j++) {
                if (i >= n0[j] && i <= n1[j]) at(j, i - n0[j]) *= s;
                else if (i < n0[j]) break;
            }
        }
        for (int j = 0; j < nc; j++) {
            int skip = 0;
            while (at(j, skip) == 0) skip++;
            n0[j] += skip;
            while (n0[j] < 0) { n0[j]++; skip++; }
            const int range = n1[j] - n0[j] + 1;
            const int mx = width < range ? width : range;
            for (int i = 0; i < mx; i++) {
                if (i + skip >= width) break;
                at(j, i) = at(j, i + skip);
            }
        }
        for (int j = 0; j < nc; j++) n1[j] = n1[j] < out_size - 1 ? n1[j] : out_size - 1;
        /* scatter order = ascending contributor; per output that is the order of its additions */
        std::vector<std::vector<int>> who(out_size);
        for (int j = 0; j < nc; j++)
            for (int k = n0[j]; k <= n1[j]; k++)
                if (k >= 0 && k < out_size) who[k].push_back(j);
        for (int o = 0; o < out_size; o++) {
            for (int j : who[o]) {
                t->src.push_back(clampi(j - margin));
                t->w.push_back(at(j, o - n0[j]));
            }
            t->start.push_back((int)t->src.size());
        }
    }
}

/* the caller's letterbox geometry (mars_yolo_test.c:46-49) */
struct LetterboxGeom {
    int ow, oh, tw, th, nw, nh, px, py;
};
static inline LetterboxGeom letterbox_geom(int ow, int oh, int tw, int th) {
    LetterboxGeom g;
    g.ow = ow; g.oh = oh; g.tw = tw; g.th = th;
    const float scale = fminf((float)tw / ow, (float)th / oh);
    g.nw = (int)(ow * scale);
    g.nh = (int)(oh * scale);
    g.px = (tw - g.nw) / 2;
    g.py = (th - g.nh) / 2;
    return g;
}

/* frames: n x (oh x ow x 3) bytes, H: n x oh x nw x 3 floats.  grid (ceil(nw/128), oh, n) */
__global__ void __launch_bounds__(128) k_pre_horizontal(const uint8_t *frames, size_t frame_stride, int ow, int oh, int nw,
                                                        const int *hstart, const int *hsrc, const float *hw, float *H) {
    __shared__ float dec[256]; /* the decode step: (float)u8 / 255.0f, IEEE division */
    dec[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.0f);
    dec[threadIdx.x + 128] = __fdiv_rn((float)(threadIdx.x + 128), 255.0f);
    __syncthreads();
    const int x = blockIdx.x * 128 + threadIdx.x, r = blockIdx.y, f = blockIdx.z;
    if (x >= nw) return;
    const uint8_t *row = frames + (size_t)f * frame_stride + (size_t)r * ow * 3;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
    for (int t = hstart[x]; t < hstart[x + 1]; t++) {
        const uint8_t *p = row + (size_t)hsrc[t] * 3;
        const float c = hw[t];
        a0 = __fadd_rn(a0, __fmul_rn(dec[p[0]], c));
        a1 = __fadd_rn(a1, __fmul_rn(dec[p[1]], c));
        a2 = __fadd_rn(a2, __fmul_rn(dec[p[2]], c));
    }
    float *o = H + (((size_t)f * oh + r) * nw + x) * 3;
    o[0] = a0; o[1] = a1; o[2] = a2;
}

/* the resized sample as stb stores it: (unsigned char)(int)((double)(saturate(v) * 255.0f) + 0.5) */
__device__ __forceinline__ int pre_encode_u8(float v) {
    const float s = v < 0 ? 0.0f : (v > 1 ? 1.0f : v);
    return (int)(unsigned char)__double2int_rz((double)__fmul_rn(s, 255.0f) + 0.5);
}

/* packing of the target frame: the two int8 tensor layouts of mars_yolo_test.c:62-72 (px - 128, border -17), or the RGBA
 * uint8 frame of examples/yolo_detect.cpp:104-124 (memset 114 -- alpha included -- then R, G, B, 0 inside the image) */
enum { PRE_NCHW_I8 = 0, PRE_NHWC_I8 = 1, PRE_RGBA_U8 = 2 };

/* one thread per target pixel: grid (ceil(tw/128), th, n); dst = target of frame 0 of the launch, frames slot_stride apart */
__global__ void __launch_bounds__(128) k_pre_vertical_pack(const float *H, LetterboxGeom g, const int *vstart, const int *vsrc,
                                                           const float *vw, int8_t *dst, size_t slot_stride, int mode) {
    const int dx = blockIdx.x * 128 + threadIdx.x, dy = blockIdx.y, f = blockIdx.z;
    if (dx >= g.tw) return;
    const int x = dx - g.px, y = dy - g.py;
    const bool inside = x >= 0 && x < g.nw && y >= 0 && y < g.nh;
    int u0 = 0, u1 = 0, u2 = 0;
    if (inside) {
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
        for (int t = vstart[y]; t < vstart[y + 1]; t++) {
            const float *p = H + (((size_t)f * g.oh + vsrc[t]) * g.nw + x) * 3;
            const float c = vw[t];
            a0 = __fadd_rn(a0, __fmul_rn(p[0], c));
            a1 = __fadd_rn(a1, __fmul_rn(p[1], c));
            a2 = __fadd_rn(a2, __fmul_rn(p[2], c));
        }
        u0 = pre_encode_u8(a0); u1 = pre_encode_u8(a1); u2 = pre_encode_u8(a2);
    }
    int8_t *o = dst + (size_t)f * slot_stride;
    if (mode == PRE_RGBA_U8) {
        uchar4 px = inside ? make_uchar4((unsigned char)u0, (unsigned char)u1, (unsigned char)u2, 0) : make_uchar4(114, 114, 114, 114);
        reinterpret_cast<uchar4 *>(o)[(size_t)dy * g.tw + dx] = px;
        return;
    }
    const int r0 = inside ? u0 - 128 : -17, r1 = inside ? u1 - 128 : -17, r2 = inside ? u2 - 128 : -17;
    if (mode == PRE_NHWC_I8) {
        int8_t *q = o + ((size_t)dy * g.tw + dx) * 3;
        q[0] = (int8_t)r0; q[1] = (int8_t)r1; q[2] = (int8_t)r2;
    } else {
        const size_t ps = (size_t)g.tw * g.th, at = (size_t)dy * g.tw + dx;
        o[at] = (int8_t)r0; o[ps + at] = (int8_t)r1; o[2 * ps + at] = (int8_t)r2;
    }
}

/* device copies of the two tap lists of one geometry */
struct LetterboxPlan {
    LetterboxGeom g;
    int *d_hstart = nullptr, *d_hsrc = nullptr, *d_vstart = nullptr, *d_vsrc = nullptr;
    float *d_hw = nullptr, *d_vw = nullptr;
    bool ok = false;
};

static inline void letterbox_plan_release(LetterboxPlan *p) {
    cudaFree(p->d_hstart); cudaFree(p->d_hsrc); cudaFree(p->d_hw);
    cudaFree(p->d_vstart); cudaFree(p->d_vsrc); cudaFree(p->d_vw);
    *p = LetterboxPlan();
}

static inline bool letterbox_plan_build(int ow, int oh, int tw, int th, LetterboxPlan *p) {
    if (p->ok && p->g.ow == ow && p->g.oh == oh && p->g.tw == tw && p->g.th == th) return true;
    letterbox_plan_release(p);
    if (ow <= 0 || oh <= 0 || tw <= 0 || th <= 0) return false;
    p->g = letterbox_geom(ow, oh, tw, th);
    if (p->g.nw <= 0 || p->g.nh <= 0 || p->g.nw > tw || p->g.nh > th) return false;
    ResizeTaps h, v;
    build_resize_taps(ow, p->g.nw, &h);
    build_resize_taps(oh, p->g.nh, &v);
    auto up_i = [](const std::vector<int> &s, int **d) {
        return cudaMalloc(d, (s.size() + 1) * sizeof(int)) == cudaSuccess &&
               cudaMemcpy(*d, s.data(), s.size() * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
    };
    auto up_f = [](const std::vector<float> &s, float **d) {
        return cudaMalloc(d, (s.size() + 1) * sizeof(float)) == cudaSuccess &&
               cudaMemcpy(*d, s.data(), s.size() * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess;
    };
    p->ok = up_i(h.start, &p->d_hstart) && up_i(h.src, &p->d_hsrc) && up_f(h.w, &p->d_hw) && up_i(v.start, &p->d_vstart) &&
            up_i(v.src, &p->d_vsrc) && up_f(v.w, &p->d_vw);
    if (!p->ok) letterbox_plan_release(p);
    return p->ok;
}

/* frames (device) -> int8 tensors at dst + f * slot_stride, f < n; H = scratch of n * oh * nw * 3 floats */
static inline cudaError_t launch_letterbox(const LetterboxPlan &p, const uint8_t *d_frames, size_t frame_stride, int n, float *d_H,
                                           int8_t *dst, size_t slot_stride, int mode, cudaStream_t s) {
    const LetterboxGeom &g = p.g;
    k_pre_horizontal<<<dim3((g.nw + 127) / 128, g.oh, n), 128, 0, s>>>(d_frames, frame_stride, g.ow, g.oh, g.nw, p.d_hstart, p.d_hsrc,
                                                                       p.d_hw, d_H);
    k_pre_vertical_pack<<<dim3((g.tw + 127) / 128, g.th, n), 128, 0, s>>>(d_H, g, p.d_vstart, p.d_vsrc, p.d_vw, dst, slot_stride, mode);
    return cudaGetLastError();
}

} // namespace marsb200
