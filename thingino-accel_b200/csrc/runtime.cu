/*
 * runtime.cu -- libmars_b200.so: the mars C runtime API over an HBM arena.
 *
 * Host side (stays C-callable, mirrors reference src/mars/mars_runtime.c):
 *   mars_load_memory/file : parse the .mars tables, run the reference's work-buffer
 *                           planner (:248-337) to get every tensor's ARENA OFFSET, upload
 *                           the weight blob once, compile the layer table into device ops;
 *   mars_run              : H2D of the input work buffers, replay the ops on one image
 *                           slot, D2H of the output work buffers (synchronous, like the
 *                           reference's :439-459);
 *   mars_b200_*           : the same ops over a batch of image slots resident in HBM,
 *                           plus YOLO decode + NMS on the device.
 * There is no CPU execution path: without a CUDA device every entry point fails with
 * MARS_ERR_NNA_INIT_FAILED.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mxu_ops.h"
#include "../../include/mars_math.h"
#include "../../include/nna.h"
#include "../../include/nna_memory.h"
#include "conv_tc.h"
#include "conv_tf32.h"
#include "kernels_exact.cuh"
#include "kernels_fast.cuh"
#include "mars_internal.h"
#include "postproc.cuh"
#include "preproc.cuh"
#include "nna_layout.cuh"

namespace marsb200 {

/* ---- diagnostics ------------------------------------------------------------ */
static thread_local char g_err[768];
void set_last_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    if (getenv("MARS_VERBOSE")) fprintf(stderr, "mars_b200: %s\n", g_err);
}

#define CU_OK(call, ret)                                                                       \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            set_last_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return ret;                                                                        \
        }                                                                                      \
    } while (0)

bool pdl_enabled() {
    /* measured on B200 (one_step.py, yolov5s-shaped graph, CUDA-graph replay): 5.93 ms with PDL against 5.58 ms without at 128
     * images, 39.0 against 37.6 ms at 1024 -- the early CTAs of the next kernel take shared memory and TMEM from the persistent
     * CTAs still running.  Off unless MARS_PDL=1. */
    static const bool on = getenv("MARS_PDL") && atoi(getenv("MARS_PDL")) != 0;
    return on;
}

bool first_time_on_device(unsigned long long *mask) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return true;
    if ((*mask >> dev) & 1ull) return false;
    *mask |= 1ull << dev;
    return true;
}

/* ---- process-wide device state (reference: g_nna_dev, src/device.c:105-131) -- */
static int g_strict = -1; /* -1: MARS_STRICT from the environment; 1: a failed tensor-core plan is an error, not a silent drop to the direct kernels */
static bool strict_mode() {
    if (g_strict < 0) { const char *e = getenv("MARS_STRICT"); g_strict = (e && atoi(e) != 0) ? 1 : 0; }
    return g_strict != 0;
}
static int g_device = -1;
static bool g_ready = false;
static size_t g_arena_bytes = 0;
struct Model;
static Model *g_last_model = nullptr;

static int pick_device() {
    if (g_device >= 0) return g_device;
    const char *e = getenv("MARS_DEVICE");
    if (e && *e) return atoi(e);
    e = getenv("LOCAL_RANK"); /* one process per GPU under torchrun */
    if (e && *e) {
        int n = 0;
        if (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) return atoi(e) % n;
    }
    return 0;
}

static int device_up() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_last_error("no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
        return NNA_ERROR_DEVICE;
    }
    int d = pick_device();
    if (d >= n) {
        set_last_error("device ordinal %d out of range (%d visible)", d, n);
        return NNA_ERROR_DEVICE;
    }
    e = cudaSetDevice(d);
    if (e != cudaSuccess) {
        set_last_error("cudaSetDevice(%d): %s", d, cudaGetErrorString(e));
        return NNA_ERROR_DEVICE;
    }
    g_device = d;
    g_ready = true;
    return NNA_SUCCESS;
}

/* ---- helpers shared with program.cpp ---------------------------------------- */
/* reference src/mars/mars_runtime.c:80-124: byte size by dtype AND format tag */
size_t tensor_byte_size(const mars_tensor_t *t) {
    size_t es;
    switch (t->dtype) {
        case MARS_DTYPE_FLOAT32: case MARS_DTYPE_INT32: es = 4; break;
        case MARS_DTYPE_INT16: es = 2; break;
        default: es = 1;
    }
    if (t->format == MARS_FORMAT_NDHWC32 && t->ndims >= 4) {
        int c32 = (t->shape[1] + 31) / 32;
        return (size_t)(t->shape[0] * c32 * t->shape[2] * t->shape[3] * 32) * es;
    }
    if (t->format == MARS_FORMAT_NMHWSOIB2 && t->ndims >= 4) {
        int no = (t->shape[0] + 31) / 32, mi = (t->shape[1] + 31) / 32;
        return (size_t)(no * mi * t->shape[2] * t->shape[3] * 1024);
    }
    size_t numel = 1;
    for (uint32_t i = 0; i < t->ndims && i < MARS_MAX_DIMS; i++) numel *= (size_t)t->shape[i];
    if (t->dtype == MARS_DTYPE_UINT4) return (numel + 1) / 2;
    return numel * es;
}

/* reference src/mars/mars_runtime.c:713-721: first table entry with a matching id */
int find_tensor(const mars_header_t &h, const mars_runtime_tensor_t *tensors, uint32_t id) {
    if (id == 0xFFFFFFFFu) return -1;
    for (uint32_t i = 0; i < h.num_tensors; i++)
        if (tensors[i].desc.id == id) return (int)i;
    return -1;
}

/* ---- the private model object ------------------------------------------------ */
static const uint32_t MODEL_TAG = 0x4232304Du; /* "M02B" */

struct Model {
    mars_model_t pub; /* must stay first: applications hold &pub */
    uint32_t tag = MODEL_TAG;
    int device = 0;
    size_t arena_size = 0, weights_size = 0, buffer_size = 0, slot_bytes = 0, slot_stride = 0;
    int num_buffers = 0;
    std::vector<size_t> toff;
    uint8_t *h_arena = nullptr;  /* pinned host mirror of [weights | buf0 | buf1 | (buf2)] */
    uint8_t *d_weights = nullptr;
    uint8_t *d_slots = nullptr;
    int capacity = 0;
    uint8_t *d_scratch = nullptr;
    size_t scratch_stride = 0;
    uint8_t *d_cpool = nullptr;
    uint8_t *d_tc_scratch = nullptr; /* padded / phase-split conv inputs, one region per image slot */
    size_t tc_scratch_stride = 0;
    uint8_t *d_linked = nullptr;     /* producer-written conv input copies (Program::linked_bytes per image), pads stay zero */
    size_t linked_stride = 0;
    Program prog;
    int opt_level = 3, depthwise_mode = 0;
    int f32_mode = 2; /* float32 convolutions: 0 = exact-order fp32 (bit-exact control), 1 = tf32, 2 = tf32x3 (conv_tf32.cu) */
    cudaStream_t stream = nullptr, h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    float last_ms = 0.0f;
    uint64_t launches = 0;
    /* asynchronous batches (mars_b200_submit_batch / wait_batch): one in flight per half of the slot pool */
    struct Pending { bool active = false; int n = 0, maxd = 0; int32_t *counts = nullptr; } pending[2];
    /* letterbox pre-processing (mars_b200_preprocess_batch): tap lists of the last geometry, frame staging, float rows */
    LetterboxPlan pre_plan;
    uint8_t *d_frames = nullptr;
    size_t frames_cap = 0;
    float *d_preH = nullptr;
    size_t preH_cap = 0;
    /* captured steps (mars_b200_step_resident): one CUDA graph per (first, n, threshold, with_detect) seen twice */
    struct StepGraph { int first, n, with_detect; uint32_t thresh_bits; int seen; cudaGraphExec_t exec; uint64_t launches; };
    std::vector<StepGraph> graphs;
    bool graphs_broken = false;
    /* device-resident detections */
    mars_det_t *d_raw = nullptr, *d_det = nullptr;
    int32_t *d_raw_cnt = nullptr, *d_det_cnt = nullptr;
    unsigned *d_nms_mask = nullptr; /* capacity x 1024 x 32 words: suppression bit matrices */
    DecodeTables *d_tab = nullptr;
    float tab_scale = 0.0f;
    bool tab_valid = false;
    /* per-op profile (CUDA events around every op when enabled) */
    int profile = 0;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<double> prof_ms;
    std::vector<uint64_t> prof_calls;
    bool prof_pending = false;
    /* tensor-core conv plans (conv_tc.cu) */
    std::vector<TcPlan> tc;
    bool compiled = false;
    uint64_t weights_hash = 0; /* FNV-1a of the host weight blob the compiled plans were built from */
};

static Model *as_model(mars_model_t *p) {
    if (!p) return nullptr;
    Model *m = reinterpret_cast<Model *>(p);
    return m->tag == MODEL_TAG ? m : nullptr;
}

static inline uint8_t *dev_addr(const Model *m, size_t off, int slot) {
    return off < m->weights_size ? m->d_weights + off : m->d_slots + (size_t)slot * m->slot_stride + (off - m->weights_size);
}

static uint64_t fnv1a(const uint8_t *p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

static void release_graphs(Model *m) {
    for (auto &g : m->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    m->graphs.clear();
}

static void model_release(Model *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    release_graphs(m);
    tc_release(m->tc);
    for (auto e : m->prof_ev) cudaEventDestroy(e);
    cudaFree(m->d_weights); cudaFree(m->d_slots); cudaFree(m->d_scratch); cudaFree(m->d_cpool); cudaFree(m->d_tc_scratch); cudaFree(m->d_linked);
    cudaFree(m->d_raw); cudaFree(m->d_det); cudaFree(m->d_raw_cnt); cudaFree(m->d_det_cnt); cudaFree(m->d_tab); cudaFree(m->d_nms_mask);
    if (m->h_arena) cudaFreeHost(m->h_arena);
    if (m->ev0) cudaEventDestroy(m->ev0);
    if (m->ev1) cudaEventDestroy(m->ev1);
    for (int i = 0; i < 2; i++) {
        if (m->ev_in[i]) cudaEventDestroy(m->ev_in[i]);
        if (m->ev_done[i]) cudaEventDestroy(m->ev_done[i]);
        if (m->ev_out[i]) cudaEventDestroy(m->ev_out[i]);
    }
    letterbox_plan_release(&m->pre_plan);
    cudaFree(m->d_frames);
    cudaFree(m->d_preH);
    if (m->stream) cudaStreamDestroy(m->stream);
    if (m->h2d_stream) cudaStreamDestroy(m->h2d_stream);
    if (m->d2h_stream) cudaStreamDestroy(m->d2h_stream);
    free(m->pub.tensors);
    free(m->pub.layers);
    if (g_last_model == m) g_last_model = nullptr;
    m->tag = 0;
    delete m;
}

/* (re)allocate `capacity` image slots; slot 0 keeps its contents */
static mars_error_t set_capacity(Model *m, int capacity) {
    if (capacity < 1) capacity = 1;
    if (capacity == m->capacity) return MARS_OK;
    uint8_t *ns = nullptr, *nscr = nullptr;
    mars_det_t *nraw = nullptr, *ndet = nullptr;
    int32_t *nrc = nullptr, *ndc = nullptr;
    size_t bytes = (size_t)capacity * m->slot_stride;
    CU_OK(cudaStreamSynchronize(m->stream), MARS_ERR_LAYER_FAILED);
    CU_OK(cudaMalloc(&ns, bytes), MARS_ERR_ALLOC_FAILED);
    CU_OK(cudaMemsetAsync(ns, 0, bytes, m->stream), MARS_ERR_ALLOC_FAILED);
    if (m->d_slots) {
        size_t keep = (size_t)std::min(capacity, m->capacity) * m->slot_stride;
        CU_OK(cudaMemcpyAsync(ns, m->d_slots, keep, cudaMemcpyDeviceToDevice, m->stream), MARS_ERR_ALLOC_FAILED);
    }
    if (m->scratch_stride) {
        CU_OK(cudaMalloc(&nscr, (size_t)capacity * m->scratch_stride), MARS_ERR_ALLOC_FAILED);
    }
    CU_OK(cudaMalloc(&nraw, (size_t)capacity * MARS_MAX_DETS * sizeof(mars_det_t)), MARS_ERR_ALLOC_FAILED);
    CU_OK(cudaMalloc(&ndet, (size_t)capacity * MARS_MAX_DETS * sizeof(mars_det_t)), MARS_ERR_ALLOC_FAILED);
    unsigned *nmask = nullptr;
    CU_OK(cudaMalloc(&nmask, (size_t)capacity * NMS_SCRATCH_WORDS * sizeof(unsigned)), MARS_ERR_ALLOC_FAILED);
    CU_OK(cudaMalloc(&nrc, (size_t)capacity * sizeof(int32_t)), MARS_ERR_ALLOC_FAILED);
    CU_OK(cudaMalloc(&ndc, (size_t)capacity * sizeof(int32_t)), MARS_ERR_ALLOC_FAILED);
    CU_OK(cudaMemsetAsync(nrc, 0, (size_t)capacity * sizeof(int32_t), m->stream), MARS_ERR_ALLOC_FAILED);
    CU_OK(cudaMemsetAsync(ndc, 0, (size_t)capacity * sizeof(int32_t), m->stream), MARS_ERR_ALLOC_FAILED);
    CU_OK(cudaStreamSynchronize(m->stream), MARS_ERR_LAYER_FAILED);
    cudaFree(m->d_slots); cudaFree(m->d_scratch); cudaFree(m->d_raw); cudaFree(m->d_det);
    cudaFree(m->d_raw_cnt); cudaFree(m->d_det_cnt); cudaFree(m->d_nms_mask);
    m->d_nms_mask = nmask;
    m->d_slots = ns; m->d_scratch = nscr; m->d_raw = nraw; m->d_det = ndet; m->d_raw_cnt = nrc; m->d_det_cnt = ndc;
    m->capacity = capacity;
    m->pub.ddr_paddr = m->d_slots;
    for (uint32_t i = 0; i < m->pub.header.num_tensors; i++) m->pub.tensors[i].paddr = dev_addr(m, m->toff[i], 0);
    /* tensor maps embed slot addresses and the slot count */
    tc_release(m->tc);
    cudaFree(m->d_tc_scratch);
    m->d_tc_scratch = nullptr;
    m->tc_scratch_stride = 0;
    cudaFree(m->d_linked);
    m->d_linked = nullptr;
    m->linked_stride = 0;
    m->compiled = false;
    return MARS_OK;
}

static mars_error_t compile_model(Model *m) {
    if (m->compiled) return MARS_OK;
    release_graphs(m); /* kernel parameters (slot addresses, tensor maps, tables) are about to change */
    Program p;
    mars_error_t e = compile_program(m->pub.header, m->pub.tensors, m->pub.layers, m->toff, m->weights_size,
                                     m->arena_size, m->opt_level, m->depthwise_mode, &p, m->f32_mode);
    if (e != MARS_OK) return e;
    m->prog = std::move(p);
    cudaFree(m->d_cpool);
    m->d_cpool = nullptr;
    CU_OK(cudaMalloc(&m->d_cpool, m->prog.const_pool.size()), MARS_ERR_ALLOC_FAILED);
    CU_OK(cudaMemcpy(m->d_cpool, m->prog.const_pool.data(), m->prog.const_pool.size(), cudaMemcpyHostToDevice),
          MARS_ERR_ALLOC_FAILED);
    if (m->prog.scratch_bytes > m->scratch_stride) {
        m->scratch_stride = (m->prog.scratch_bytes + 255) & ~(size_t)255;
        cudaFree(m->d_scratch);
        m->d_scratch = nullptr;
        CU_OK(cudaMalloc(&m->d_scratch, (size_t)m->capacity * m->scratch_stride), MARS_ERR_ALLOC_FAILED);
    }
    tc_release(m->tc);
    m->tc.assign(m->prog.ops.size(), TcPlan());
    size_t need = 0;
    for (const Op &o : m->prog.ops) {
        if (o.impl == CONV_TC_NCHW) need = std::max(need, tc_scratch_need(o));
        if (o.impl == CONV_TC_F32) need = std::max(need, tf32_scratch_need(o, m->f32_mode));
    }
    need = (need + 1023) & ~(size_t)1023;
    if (need > m->tc_scratch_stride || (need && !m->d_tc_scratch)) {
        cudaFree(m->d_tc_scratch);
        m->d_tc_scratch = nullptr;
        m->tc_scratch_stride = need;
        CU_OK(cudaMalloc(&m->d_tc_scratch, (size_t)m->capacity * need), MARS_ERR_ALLOC_FAILED);
    }
    /* the link area is private to one compiled program: pads must be zero, so (re)allocate it zeroed every time */
    cudaFree(m->d_linked);
    m->d_linked = nullptr;
    m->linked_stride = (m->prog.linked_bytes + 1023) & ~(size_t)1023;
    if (m->linked_stride) {
        CU_OK(cudaMalloc(&m->d_linked, (size_t)m->capacity * m->linked_stride), MARS_ERR_ALLOC_FAILED);
        CU_OK(cudaMemset(m->d_linked, 0, (size_t)m->capacity * m->linked_stride), MARS_ERR_ALLOC_FAILED);
    }
    ArenaGeom g{m->d_weights, m->d_slots, m->weights_size, m->slot_stride, m->capacity, m->h_arena, m->prog.const_pool.data()};
    for (size_t i = 0; i < m->prog.ops.size(); i++) {
        Op &o = m->prog.ops[i];
        if (o.impl == CONV_TC_F32) { /* float32 layer on the tf32 tensor path; a failed plan simply stays on the exact kernel */
            if (!tf32_plan(o, g, m->f32_mode, m->d_tc_scratch, m->tc_scratch_stride, &m->tc[i])) {
                if (strict_mode()) {
                    char why[600];
                    snprintf(why, sizeof why, "%s", g_err);
                    set_last_error("strict mode: tf32 plan failed for layer %d (%s)", o.layer, why);
                    tc_release(m->tc);
                    return MARS_ERR_LAYER_FAILED;
                }
                o.impl = CONV_DIRECT;
            }
            continue;
        }
        if (o.impl != CONV_TC_NCHW) continue;
        const Op *consumer = o.nhwc_consumer >= 0 ? &m->prog.ops[o.nhwc_consumer] : nullptr;
        if (!tc_plan(o, g, m->d_tc_scratch, m->tc_scratch_stride, m->d_linked, m->linked_stride, consumer, &m->tc[i])) {
            /* a fused op cannot simply fall back (its followers were folded): recompile exact -- unless the caller asked
             * to be told (mars_b200_set_strict / MARS_STRICT=1): the direct kernels are ~100x slower */
            if (strict_mode()) {
                char why[600];
                snprintf(why, sizeof why, "%s", g_err);
                set_last_error("strict mode: tensor-core plan failed for layer %d (%s)", o.layer, why);
                tc_release(m->tc);
                return MARS_ERR_LAYER_FAILED;
            }
            fprintf(stderr, "mars_b200: tensor-core plan failed for layer %d (%s); using the direct CUDA kernels\n", o.layer, g_err);
            tc_release(m->tc);
            m->opt_level = 0;
            return compile_model(m);
        }
    }
    m->prof_ms.assign(m->prog.ops.size(), 0.0);
    m->prof_calls.assign(m->prog.ops.size(), 0);
    m->weights_hash = fnv1a(m->h_arena, m->weights_size);
    /* tables, the zeroed link area and the repacked weights are complete before a (non-blocking) stream uses them */
    CU_OK(cudaDeviceSynchronize(), MARS_ERR_LAYER_FAILED);
    m->compiled = true;
    return MARS_OK;
}

/* ---- op launch ---------------------------------------------------------------- */
static KOp to_kop(const Op &o) {
    KOp k;
    k.kind = o.kind; k.mode = o.mode;
    k.in0 = o.in0; k.in1 = o.in1; k.in2 = o.in2; k.out = o.out; k.w = o.w; k.bias = o.bias;
    k.ic = o.ic; k.ih = o.ih; k.iw = o.iw; k.oc = o.oc; k.oh = o.oh; k.ow = o.ow;
    k.kh = o.kh; k.kw = o.kw; k.sh = o.sh; k.sw = o.sw; k.pt = o.pt; k.pl = o.pl;
    k.coff = o.coff;
    k.f0 = o.f0; k.f1 = o.f1; k.f2 = o.f2;
    k.n = o.n;
    k.lut = o.lut; k.post_relu = o.post_relu; k.post_lut = o.post_lut;
    k.pass_oc = 0; k.use_scratch = 0;
    k.out_s = o.fused_layers > 0 ? o.out_s : -1;
    k.out_z = (o.fused_layers > 0 && o.store_z) ? o.out_z : -1;
    k.lut_s = o.lut_s; k.lut_z = o.lut_z; k.store_y = o.store_y ? 1 : 0;
    /* int8 mul / add: |a * sa| <= 128 |sa| etc.; with everything finite and the scaled result below 2^30 neither the x86
     * overflow rule nor NaN handling of the reference's float -> int conversion can trigger */
    k.fast_bin = 0;
    if (o.kind == OP_MUL_I8 || o.kind == OP_ADD_I8) {
        const double a = 128.0 * fabs((double)o.f0), b = 128.0 * fabs((double)o.f1), inv = fabs((double)o.f2);
        const double top = (o.kind == OP_MUL_I8 ? a * b : a + b) * inv + 1.0;
        k.fast_bin = (std::isfinite(o.f0) && std::isfinite(o.f1) && std::isfinite(o.f2) && top < 1073741824.0 && a < 1e30 && b < 1e30) ? 1 : 0;
    }
    return k;
}

static inline unsigned blocks_for(uint64_t n, unsigned per) { return (unsigned)((n + per - 1) / per); }

static mars_error_t launch_op(Model *m, size_t op_index, int first, int n, bool full_pass) {
    const Op &o = m->prog.ops[op_index];
    if (o.mode >= 1000) {
        set_last_error("layer %d: %s", o.layer, o.note.c_str());
        return (mars_error_t)(-(o.mode - 1000));
    }
    if (o.kind == OP_NOP || n <= 0) return MARS_OK;
    ArenaView v;
    v.wbase = m->d_weights;
    v.sbase = m->d_slots + (size_t)first * m->slot_stride;
    v.W = m->weights_size;
    v.slot_stride = m->slot_stride;
    v.cpool = m->d_cpool;
    v.scratch = m->d_scratch ? m->d_scratch + (size_t)first * m->scratch_stride : nullptr;
    v.scratch_stride = m->scratch_stride;
    KOp k = to_kop(o);
    cudaStream_t s = m->stream;
    const bool xl = o.xlat;
    const int is_conv = o.kind == OP_CONV_I8_NCHW || o.kind == OP_CONV_I8_NHWC || o.kind == OP_CONV_F32_NCHW || o.kind == OP_DW_I8;
    if (o.mode == EXEC_SERIAL) {
        if (xl) k_serial<true><<<n, 32, 0, s>>>(v, k); else k_serial<false><<<n, 32, 0, s>>>(v, k);
        m->launches++;
    } else if (is_conv) {
        const uint64_t P = (uint64_t)o.oh * o.ow;
        const int es = o.kind == OP_CONV_F32_NCHW ? 4 : 1;
        if (o.impl == CONV_TC_F32 && m->tc[op_index].valid) {
            if (!tf32_launch(m->tc[op_index], m->d_slots, first, n, s, &m->launches)) return MARS_ERR_LAYER_FAILED;
        } else if (o.impl == CONV_TC_NCHW && m->tc[op_index].valid) {
            if (!tc_launch(m->tc[op_index], m->d_slots, first, n, full_pass, s, &m->launches)) return MARS_ERR_LAYER_FAILED;
        } else if (o.mode == EXEC_PARALLEL) {
            if (o.kind == OP_CONV_I8_NCHW && !xl && fast_conv_nchw_ok(k)) {
                launch_fast_conv_nchw(v, k, n, s);
            } else if (o.kind == OP_DW_I8 && !xl && m->opt_level >= 1 && dw_nchw4_ok(v, k)) {
                launch_pdl(k_dw_nchw4, dim3(blocks_for((uint64_t)o.oc * o.oh * (o.ow >> 2), 256), n), dim3(256), (size_t)0, s, v, k);
            } else {
                dim3 g(blocks_for(P * o.oc, 256), n);
                if (xl) k_conv_point<true><<<g, 256, 0, s>>>(v, k); else k_conv_point<false><<<g, 256, 0, s>>>(v, k);
            }
            m->launches++;
        } else if (o.mode == EXEC_OC_PASSES || o.mode == EXEC_OC_PASSES_SCRATCH) {
            dim3 g(blocks_for(P, 256), n);
            for (int oc = 0; oc < o.oc; oc++) {
                k.pass_oc = oc;
                k.use_scratch = o.mode == EXEC_OC_PASSES_SCRATCH;
                if (xl) k_conv_point<true><<<g, 256, 0, s>>>(v, k); else k_conv_point<false><<<g, 256, 0, s>>>(v, k);
                m->launches++;
                if (k.use_scratch) {
                    dim3 gc(blocks_for(P * es, 256), n);
                    k_pass_commit<<<gc, 256, 0, s>>>(v, k, es);
                    m->launches++;
                }
            }
        } else if (o.mode == EXEC_PIXEL_SERIAL && o.kind == OP_CONV_I8_NCHW && m->opt_level >= 1 && inplace_reg_ok(v, k)) {
            launch_inplace_reg(v, k, n, s);
            m->launches++;
        } else if (o.mode == EXEC_PIXEL_SERIAL && o.kind == OP_CONV_I8_NCHW && o.fused_layers > 0) {
            set_last_error("layer %d: fused in-place conv needs the register kernel", o.layer); /* the planner fuses only what inplace_reg_ok accepts */
            return MARS_ERR_LAYER_FAILED;
        } else if (o.mode == EXEC_PIXEL_SERIAL && o.kind == OP_CONV_I8_NCHW) {
            dim3 g(blocks_for(P, 128), n);
            const size_t smem = (size_t)o.ic * 128 + (size_t)o.oc * o.ic;
            static unsigned long long attr_set = 0;
            if (first_time_on_device(&attr_set)) cudaFuncSetAttribute(k_conv1x1_nchw_inplace, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            k_conv1x1_nchw_inplace<<<g, 128, smem, s>>>(v, k);
            m->launches++;
        } else if (o.mode == EXEC_PIXEL_SERIAL) {
            dim3 g(blocks_for(P, 128), n);
            if (xl) k_conv_nhwc_pixel_serial<true><<<g, 128, 0, s>>>(v, k); else k_conv_nhwc_pixel_serial<false><<<g, 128, 0, s>>>(v, k);
            m->launches++;
        }
    } else if (o.kind == OP_MAXPOOL || o.kind == OP_UPSAMPLE || o.kind == OP_CONCAT || o.kind == OP_CONCAT_PERIODIC) {
        uint64_t total = (o.kind == OP_CONCAT || o.kind == OP_CONCAT_PERIODIC) ? o.n : (uint64_t)o.oh * o.ow * o.ic;
        if (o.kind == OP_CONCAT_PERIODIC) {
            /* chunks of `coff` bytes depend on the previous chunk only through position j mod coff,
             * whose source bytes [0, coff) are never written: order-free */
        }
        if (!xl && m->opt_level >= 1 && fast_spatial_ok(v, k)) {
            launch_fast_spatial(v, k, n, s);
        } else {
            dim3 g(blocks_for(total, 256), n);
            if (xl) k_spatial<true><<<g, 256, 0, s>>>(v, k); else k_spatial<false><<<g, 256, 0, s>>>(v, k);
        }
        m->launches++;
    } else {
        if (!xl && m->opt_level >= 1 && fast_flat_ok(v, k)) {
            launch_fast_flat(v, k, n, s);
        } else {
            dim3 g(blocks_for(o.n, 256), n);
            if (xl) k_flat<true><<<g, 256, 0, s>>>(v, k); else k_flat<false><<<g, 256, 0, s>>>(v, k);
        }
        m->launches++;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("layer %d launch failed: %s", o.layer, cudaGetErrorString(e));
        return MARS_ERR_LAYER_FAILED;
    }
    return MARS_OK;
}

/* all ops (or the ops of one layer, or the ops [op_lo, op_hi)) over slots [first, first+n) on m->stream */
static mars_error_t enqueue_ops(Model *m, int first, int n, int only_layer, size_t op_lo = 0, size_t op_hi = (size_t)-1) {
    mars_error_t e = compile_model(m);
    if (e != MARS_OK) return e;
    const size_t nops = m->prog.ops.size();
    if (op_lo != 0 || op_hi < nops) { /* a slice of the op list (micro-batched head / full-batch tail): no per-op events */
        for (size_t i = op_lo; i < std::min(op_hi, nops); i++) {
            e = launch_op(m, i, first, n, true);
            if (e != MARS_OK) return e;
        }
        return MARS_OK;
    }
    if (m->profile && m->prof_ev.size() < nops + 1) {
        while (m->prof_ev.size() < nops + 1) {
            cudaEvent_t ev;
            CU_OK(cudaEventCreate(&ev), MARS_ERR_ALLOC_FAILED);
            m->prof_ev.push_back(ev);
        }
    }
    const bool prof = m->profile && only_layer < 0;
    for (size_t i = 0; i < nops; i++) {
        if (only_layer >= 0 && m->prog.ops[i].layer != only_layer) continue;
        if (prof) cudaEventRecord(m->prof_ev[i], m->stream);
        e = launch_op(m, i, first, n, only_layer < 0);
        if (e != MARS_OK) return e;
    }
    if (prof) {
        cudaEventRecord(m->prof_ev[nops], m->stream);
        m->prof_pending = true;
    }
    return MARS_OK;
}

/* after the stream was synchronised: fold the per-op event intervals of the last full pass */
static void profile_collect(Model *m) {
    if (!m->prof_pending) return;
    m->prof_pending = false;
    const size_t nops = m->prog.ops.size();
    for (size_t i = 0; i < nops; i++) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, m->prof_ev[i], m->prof_ev[i + 1]) == cudaSuccess) {
            m->prof_ms[i] += ms;
            m->prof_calls[i]++;
        }
    }
}

static bool range_ok(Model *m, int first, int n) {
    if (first < 0 || n < 0 || first + n > m->capacity) {
        set_last_error("slots [%d,%d) outside the batch capacity %d (mars_b200_set_batch)", first, first + n, m->capacity);
        return false;
    }
    return true;
}

/* ---- decode + NMS on the device ----------------------------------------------- */
static void build_decode_tables(float scale, DecodeTables *t) {
    for (int v = -128; v < 128; v++) {
        /* reference src/mars/mars_yolo_test.c:84,92 -- host libm expf, one rounding per operation */
        volatile float a = -(float)v * scale;
        volatile float e = expf(a);
        volatile float den = 1.0f + e;
        volatile float obj = 1.0f / den;
        t->obj[v + 128] = obj;
        volatile float s = (float)v * scale;
        volatile float e2 = expf(-s);
        volatile float den2 = 1.0f + e2;
        t->den[v + 128] = den2;
    }
    volatile float e3 = expf(1e9f);
    volatile float d3 = 1.0f + e3;
    t->den[256] = d3;
}

static mars_error_t ensure_tables(Model *m, float scale) {
    if (m->tab_valid && memcmp(&m->tab_scale, &scale, 4) == 0) return MARS_OK;
    DecodeTables t;
    build_decode_tables(scale, &t);
    if (!m->d_tab) CU_OK(cudaMalloc(&m->d_tab, sizeof(DecodeTables)), MARS_ERR_ALLOC_FAILED);
    CU_OK(cudaMemcpyAsync(m->d_tab, &t, sizeof t, cudaMemcpyHostToDevice, m->stream), MARS_ERR_LAYER_FAILED);
    CU_OK(cudaStreamSynchronize(m->stream), MARS_ERR_LAYER_FAILED); /* t is a stack object */
    m->tab_scale = scale;
    m->tab_valid = true;
    return MARS_OK;
}

static mars_error_t enqueue_detect(Model *m, int first, int n, float thresh) {
    if (m->pub.header.num_outputs < 1) return MARS_ERR_INVALID_TENSOR;
    uint32_t oi = m->pub.header.output_tensor_ids[0];
    if (oi >= m->pub.header.num_tensors) return MARS_ERR_INVALID_TENSOR;
    const mars_tensor_t &d = m->pub.tensors[oi].desc;
    if (d.ndims < 3 || d.shape[2] != 85 || d.shape[1] <= 0 || m->toff[oi] < m->weights_size) {
        set_last_error("detect: output 0 is not an int8 [1,N,85] head (reference src/mars/mars_yolo_test.c:187)");
        return MARS_ERR_INVALID_TENSOR;
    }
    mars_error_t e = ensure_tables(m, d.scale);
    if (e != MARS_OK) return e;
    const int8_t *data = reinterpret_cast<const int8_t *>(dev_addr(m, m->toff[oi], first));
    int nms_launches = 1;
    launch_pdl(k_parse_output, dim3(n), dim3(256), 0, m->stream, data, m->slot_stride, d.shape[1], d.scale, m->d_tab,
                                            m->d_raw + (size_t)first * MARS_MAX_DETS, m->d_raw_cnt + first, 1000,
                                            MARS_MAX_DETS);
    CU_OK(launch_nms_center(m->d_raw + (size_t)first * MARS_MAX_DETS, m->d_raw_cnt + first, m->d_det + (size_t)first * MARS_MAX_DETS,
                            m->d_det_cnt + first, MARS_MAX_DETS, thresh, n, m->d_nms_mask + (size_t)first * NMS_SCRATCH_WORDS, m->stream, &nms_launches),
          MARS_ERR_LAYER_FAILED);
    m->launches += 1 + nms_launches;
    CU_OK(cudaGetLastError(), MARS_ERR_LAYER_FAILED);
    return MARS_OK;
}

static inline double now_us() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

} // namespace marsb200

using namespace marsb200;

/* ================================================================================ */
/* the reference API (include/mars_runtime.h)                                        */
/* ================================================================================ */
extern "C" {

/* reference src/mars/mars_runtime.c:58-76 */
const char *mars_get_error_string(mars_error_t err) {
    static const char *s[] = {"OK", "Invalid magic number", "Version mismatch", "Memory allocation failed",
                              "Invalid file format", "NNA initialization failed", "Layer execution failed",
                              "Invalid tensor", "Invalid layer"};
    int i = -(int)err;
    return (i >= 0 && i < 9) ? s[i] : "Unknown error";
}

/* reference src/mars/mars_runtime.c:126-349 */
/* The planner's decisions for a .mars blob without touching a device: parses the tables, lays the tensors out as mars_load_memory
 * does (reference src/mars/mars_runtime.c:248-337), compiles the layer table at `opt_level` and writes the op list (one line per
 * op: kind, mode, kernel choice, fused / linked / forwarded / elided streams) to dst.  Returns the length of the description
 * (0 on error).  Host logic only -- the CPU test suite checks fusion, concat trimming and forwarding through it. */
size_t mars_b200_plan_describe(const void *data, size_t size, size_t arena_bytes, int opt_level, char *dst, size_t cap) {
    if (!data || size < sizeof(mars_header_t)) return 0;
    mars_header_t h;
    memcpy(&h, data, sizeof h);
    if (h.magic != MARS_MAGIC || h.version_major != MARS_VERSION_MAJOR) return 0;
    const size_t tables = sizeof h + (size_t)h.num_tensors * sizeof(mars_tensor_t) + (size_t)h.num_layers * sizeof(mars_layer_t);
    if (size < tables || h.weights_offset > size || h.weights_size > size - h.weights_offset) return 0;
    std::vector<mars_runtime_tensor_t> tensors(h.num_tensors ? h.num_tensors : 1);
    std::vector<mars_runtime_layer_t> layers(h.num_layers ? h.num_layers : 1);
    memset(tensors.data(), 0, tensors.size() * sizeof(mars_runtime_tensor_t));
    memset(layers.data(), 0, layers.size() * sizeof(mars_runtime_layer_t));
    const uint8_t *p = (const uint8_t *)data + sizeof h;
    for (uint32_t i = 0; i < h.num_tensors; i++, p += sizeof(mars_tensor_t)) memcpy(&tensors[i].desc, p, sizeof(mars_tensor_t));
    for (uint32_t i = 0; i < h.num_layers; i++, p += sizeof(mars_layer_t)) memcpy(&layers[i].desc, p, sizeof(mars_layer_t));
    const size_t arena = arena_bytes ? arena_bytes : (size_t)8 << 20, W = (size_t)h.weights_size;
    if (W > arena) return 0;
    size_t remaining = arena - W, maxsz = 0;
    for (uint32_t i = 0; i < h.num_tensors; i++)
        if (tensors[i].desc.data_size == 0) maxsz = std::max(maxsz, (tensor_byte_size(&tensors[i].desc) + 63) & ~(size_t)63);
    size_t nb = 3, bs = maxsz;
    if (bs * nb > remaining) nb = 2;
    if (bs * nb > remaining) { bs = (remaining / 2) & ~(size_t)63; if (bs < 65536) return 0; }
    std::vector<size_t> toff(h.num_tensors ? h.num_tensors : 1);
    uint32_t k = 0;
    for (uint32_t i = 0; i < h.num_tensors; i++) {
        if (tensors[i].desc.data_size > 0) { toff[i] = (size_t)tensors[i].desc.data_offset; tensors[i].alloc_size = (size_t)tensors[i].desc.data_size; }
        else { toff[i] = W + (size_t)(k % nb) * bs; tensors[i].alloc_size = bs; k++; }
    }
    Program prog;
    if (compile_program(h, tensors.data(), layers.data(), toff, W, arena, opt_level, 0, &prog, 2 /* the library's default float32 mode: tf32x3 */) != MARS_OK) return 0;
    const std::string s = describe_program(prog);
    if (dst && cap) {
        const size_t n = std::min(cap - 1, s.size());
        memcpy(dst, s.data(), n);
        dst[n] = 0;
    }
    return s.size();
}

mars_error_t mars_load_memory(const void *data, size_t size, mars_model_t **out_model) {
    if (!data || !out_model || size < sizeof(mars_header_t)) return MARS_ERR_INVALID_FILE;
    mars_header_t h;
    memcpy(&h, data, sizeof h);
    if (h.magic != MARS_MAGIC) {
        fprintf(stderr, "Mars: Invalid magic 0x%08x (expected 0x%08x)\n", h.magic, MARS_MAGIC);
        return MARS_ERR_INVALID_MAGIC;
    }
    if (h.version_major != MARS_VERSION_MAJOR) {
        fprintf(stderr, "Mars: Version mismatch %d.%d (expected %d.x)\n", h.version_major, h.version_minor, MARS_VERSION_MAJOR);
        return MARS_ERR_VERSION_MISMATCH;
    }
    const size_t tables = sizeof h + (size_t)h.num_tensors * sizeof(mars_tensor_t) + (size_t)h.num_layers * sizeof(mars_layer_t);
    if (size < tables || h.weights_offset > size || h.weights_size > size - h.weights_offset) {
        set_last_error("file truncated: %zu bytes, tables need %zu, weights [%llu,+%llu)", size, tables,
                       (unsigned long long)h.weights_offset, (unsigned long long)h.weights_size);
        return MARS_ERR_INVALID_FILE; /* the reference would read past the buffer here */
    }
    if (!g_ready && device_up() != NNA_SUCCESS) return MARS_ERR_NNA_INIT_FAILED;
    CU_OK(cudaSetDevice(g_device), MARS_ERR_NNA_INIT_FAILED);

    Model *m = new (std::nothrow) Model();
    if (!m) return MARS_ERR_ALLOC_FAILED;
    memset(&m->pub, 0, sizeof m->pub);
    m->device = g_device;
    { const char *e = getenv("MARS_F32_MODE"); if (e && *e) m->f32_mode = std::max(0, std::min(2, atoi(e))); }
    m->pub.header = h;
    m->pub.tensors = (mars_runtime_tensor_t *)calloc(h.num_tensors ? h.num_tensors : 1, sizeof(mars_runtime_tensor_t));
    m->pub.layers = (mars_runtime_layer_t *)calloc(h.num_layers ? h.num_layers : 1, sizeof(mars_runtime_layer_t));
    if (!m->pub.tensors || !m->pub.layers) { model_release(m); return MARS_ERR_ALLOC_FAILED; }
    const uint8_t *p = (const uint8_t *)data + sizeof h;
    for (uint32_t i = 0; i < h.num_tensors; i++, p += sizeof(mars_tensor_t)) memcpy(&m->pub.tensors[i].desc, p, sizeof(mars_tensor_t));
    for (uint32_t i = 0; i < h.num_layers; i++, p += sizeof(mars_layer_t)) memcpy(&m->pub.layers[i].desc, p, sizeof(mars_layer_t));

    /* arena size: the reference's literal 8 MiB (:209) unless the application asked for more */
    size_t arena = g_arena_bytes;
    if (!arena) {
        const char *e = getenv("MARS_ARENA_BYTES");
        if (e && *e) arena = (size_t)strtoull(e, nullptr, 0);
    }
    if (!arena) arena = (size_t)8 << 20;
    m->arena_size = arena;
    m->weights_size = (size_t)h.weights_size;
    if (m->weights_size > arena) {
        fprintf(stderr, "Mars: Weights too large (%zu > %zu)\n", m->weights_size, arena);
        model_release(m);
        return MARS_ERR_ALLOC_FAILED;
    }

    /* the planner of reference :248-337, producing arena offsets */
    size_t remaining = arena - m->weights_size, maxsz = 0;
    for (uint32_t i = 0; i < h.num_tensors; i++)
        if (m->pub.tensors[i].desc.data_size == 0) {
            size_t sz = (tensor_byte_size(&m->pub.tensors[i].desc) + 63) & ~(size_t)63;
            maxsz = std::max(maxsz, sz);
        }
    size_t nb = 3, bs = maxsz;
    if (bs * nb > remaining) nb = 2;
    if (bs * nb > remaining) {
        bs = (remaining / 2) & ~(size_t)63;
        if (bs < 65536) {
            fprintf(stderr, "Mars: Not enough DDR for work buffers (set MARS_ARENA_BYTES / mars_b200_set_arena_bytes)\n");
            model_release(m);
            return MARS_ERR_ALLOC_FAILED;
        }
    }
    m->num_buffers = (int)nb;
    m->buffer_size = bs;
    m->slot_bytes = nb * bs;
    m->slot_stride = (m->slot_bytes + 1023) & ~(size_t)1023;
    if (m->slot_stride == 0) m->slot_stride = 1024;
    m->toff.resize(h.num_tensors ? h.num_tensors : 1);

    if (cudaHostAlloc((void **)&m->h_arena, arena + 4096, cudaHostAllocDefault) != cudaSuccess) {
        set_last_error("pinned host mirror of %zu bytes: %s", arena, cudaGetErrorString(cudaGetLastError()));
        model_release(m);
        return MARS_ERR_ALLOC_FAILED;
    }
    memset(m->h_arena, 0, arena + 4096);
    memcpy(m->h_arena, (const uint8_t *)data + h.weights_offset, m->weights_size);
    /* 64 KiB of zeroed slack behind the weights keeps stray device reads mapped */
    if (cudaMalloc(&m->d_weights, m->weights_size + 65536) != cudaSuccess ||
        cudaMemset(m->d_weights, 0, m->weights_size + 65536) != cudaSuccess ||
        cudaMemcpy(m->d_weights, m->h_arena, m->weights_size, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&m->h2d_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&m->d2h_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&m->ev0) != cudaSuccess || cudaEventCreate(&m->ev1) != cudaSuccess) {
        set_last_error("device setup: %s", cudaGetErrorString(cudaGetLastError()));
        model_release(m);
        return MARS_ERR_ALLOC_FAILED;
    }
    for (int i = 0; i < 2; i++) {
        cudaEventCreateWithFlags(&m->ev_in[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&m->ev_done[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&m->ev_out[i], cudaEventDisableTiming);
    }

    uint32_t k = 0;
    for (uint32_t i = 0; i < h.num_tensors; i++) {
        mars_runtime_tensor_t &t = m->pub.tensors[i];
        if (t.desc.data_size > 0) { /* weight tensor: points into the blob (:321-326) */
            m->toff[i] = (size_t)t.desc.data_offset;
            t.alloc_size = (size_t)t.desc.data_size;
        } else { /* runtime tensor: round-robin work buffer (:316-334) */
            m->toff[i] = m->weights_size + (size_t)(k % nb) * bs;
            t.alloc_size = bs;
            k++;
        }
        t.vaddr = m->toff[i] <= arena ? m->h_arena + m->toff[i] : nullptr;
        t.is_external = false;
    }
    m->pub.ddr_base = m->h_arena;
    m->pub.ddr_size = arena;
    m->pub.oram_base = nullptr;
    m->pub.oram_paddr = nullptr;
    m->pub.oram_size = 0;
    m->pub.weights = m->h_arena;
    m->pub.weights_size = m->weights_size;

    /* An image slot must reach as far as any layer reads or writes.  In the reference an access behind the last work buffer
     * lands in unused arena (extents taken from another tensor or dtype, the reduced-buffer case); here it would land in the
     * next image's slot, so the slot is sized to the furthest access of a dry compile (it stays private and zero-initialised). */
    {
        Program dry;
        mars_error_t de = compile_program(m->pub.header, m->pub.tensors, m->pub.layers, m->toff, m->weights_size, m->arena_size, 0,
                                          m->depthwise_mode, &dry);
        if (de != MARS_OK) { model_release(m); return de; }
        if (dry.max_extent > m->weights_size + m->slot_bytes) {
            m->slot_bytes = dry.max_extent - m->weights_size;
            m->slot_stride = (m->slot_bytes + 1023) & ~(size_t)1023;
        }
    }
    mars_error_t e = set_capacity(m, 1);
    if (e == MARS_OK) e = compile_model(m);
    if (e != MARS_OK) { model_release(m); return e; }
    g_last_model = m;
    *out_model = &m->pub;
    return MARS_OK;
}

/* reference src/mars/mars_runtime.c:351-386 */
mars_error_t mars_load_file(const char *path, mars_model_t **model) {
    if (!path || !model) return MARS_ERR_INVALID_FILE;
    FILE *f = fopen(path, "rb");
    if (!f) {
        fprintf(stderr, "Mars: Cannot open %s\n", path);
        return MARS_ERR_INVALID_FILE;
    }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz <= 0) { fclose(f); return MARS_ERR_INVALID_FILE; }
    void *buf = malloc((size_t)sz);
    if (!buf) { fclose(f); return MARS_ERR_ALLOC_FAILED; }
    size_t got = fread(buf, 1, (size_t)sz, f);
    fclose(f);
    if (got != (size_t)sz) { free(buf); return MARS_ERR_INVALID_FILE; }
    mars_error_t e = mars_load_memory(buf, (size_t)sz, model);
    free(buf); /* the model owns a copy of the weights inside the arena (:384) */
    return e;
}

void mars_free(mars_model_t *model) { model_release(as_model(model)); }

/* reference src/mars/mars_runtime.c:395-411: the header ids are used as table indices */
mars_runtime_tensor_t *mars_get_input(mars_model_t *model, int index) {
    if (!model || index < 0 || (uint32_t)index >= model->header.num_inputs) return nullptr;
    uint32_t id = model->header.input_tensor_ids[index];
    return id < model->header.num_tensors ? &model->tensors[id] : nullptr;
}
mars_runtime_tensor_t *mars_get_output(mars_model_t *model, int index) {
    if (!model || index < 0 || (uint32_t)index >= model->header.num_outputs) return nullptr;
    uint32_t id = model->header.output_tensor_ids[index];
    return id < model->header.num_tensors ? &model->tensors[id] : nullptr;
}
int mars_get_num_inputs(mars_model_t *model) { return model ? (int)model->header.num_inputs : 0; }
int mars_get_num_outputs(mars_model_t *model) { return model ? (int)model->header.num_outputs : 0; }

/* reference src/mars/mars_runtime.c:421-434 */
void mars_print_summary(mars_model_t *model) {
    if (!model) return;
    Model *m = as_model(model);
    printf("Mars Model Summary:\n");
    printf("  Layers: %u\n", model->header.num_layers);
    printf("  Tensors: %u\n", model->header.num_tensors);
    printf("  Inputs: %u\n", model->header.num_inputs);
    printf("  Outputs: %u\n", model->header.num_outputs);
    printf("  Weights: %zu bytes\n", model->weights_size);
    printf("  DDR: %zu bytes @ %p\n", model->ddr_size, model->ddr_base);
    if (m) printf("  B200: device %d, %d work buffers x %zu bytes, %d image slot(s), %zu device ops\n", m->device,
                  m->num_buffers, m->buffer_size, m->capacity, m->prog.ops.size());
}

/* reference src/mars/mars_runtime.c:439-459: synchronous, one image */
mars_error_t mars_run(mars_model_t *model) {
    Model *m = as_model(model);
    if (!m) return MARS_ERR_INVALID_FILE;
    CU_OK(cudaSetDevice(m->device), MARS_ERR_NNA_INIT_FAILED);
    double t0 = now_us();
    static int full_mirror = -1;
    if (full_mirror < 0) { const char *e = getenv("MARS_MIRROR"); full_mirror = (e && !strcmp(e, "full")) ? 1 : 0; }
    const size_t W = m->weights_size;
    if (full_mirror) {
        CU_OK(cudaMemcpyAsync(m->d_slots, m->h_arena + W, m->slot_bytes, cudaMemcpyHostToDevice, m->stream), MARS_ERR_LAYER_FAILED);
    } else {
        bool done[3] = {false, false, false};
        for (uint32_t i = 0; i < model->header.num_inputs && i < 4; i++) {
            uint32_t ti = model->header.input_tensor_ids[i];
            if (ti >= model->header.num_tensors || m->toff[ti] < W) continue;
            size_t b = (m->toff[ti] - W) / m->buffer_size;
            if (b < 3 && !done[b]) {
                done[b] = true;
                CU_OK(cudaMemcpyAsync(m->d_slots + b * m->buffer_size, m->h_arena + W + b * m->buffer_size, m->buffer_size,
                                      cudaMemcpyHostToDevice, m->stream), MARS_ERR_LAYER_FAILED);
            }
        }
    }
    mars_error_t e = enqueue_ops(m, 0, 1, -1);
    if (e != MARS_OK) { cudaStreamSynchronize(m->stream); return e; }
    if (full_mirror) {
        CU_OK(cudaMemcpyAsync(m->h_arena + W, m->d_slots, m->slot_bytes, cudaMemcpyDeviceToHost, m->stream), MARS_ERR_LAYER_FAILED);
    } else {
        bool done[3] = {false, false, false};
        for (uint32_t i = 0; i < model->header.num_outputs && i < 4; i++) {
            uint32_t ti = model->header.output_tensor_ids[i];
            if (ti >= model->header.num_tensors || m->toff[ti] < W) continue;
            size_t b = (m->toff[ti] - W) / m->buffer_size;
            if (b < 3 && !done[b]) {
                done[b] = true;
                CU_OK(cudaMemcpyAsync(m->h_arena + W + b * m->buffer_size, m->d_slots + b * m->buffer_size, m->buffer_size,
                                      cudaMemcpyDeviceToHost, m->stream), MARS_ERR_LAYER_FAILED);
            }
        }
    }
    cudaError_t ce = cudaStreamSynchronize(m->stream);
    if (ce != cudaSuccess) {
        set_last_error("mars_run: %s", cudaGetErrorString(ce));
        return MARS_ERR_LAYER_FAILED;
    }
    for (uint32_t i = 0; i < model->header.num_layers; i++) model->layers[i].is_executed = true;
    model->inference_count++;
    model->total_inference_us += (uint64_t)(now_us() - t0); /* the reference declares but never fills this */
    return MARS_OK;
}

/* ================================================================================ */
/* additive B200 entry points (include/mars_b200.h)                                  */
/* ================================================================================ */
const char *mars_b200_version(void) { return "mars-b200 0.1 (sm_100a)"; }
int mars_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return (int)MARS_ERR_NNA_INIT_FAILED;
    return n;
}
int mars_b200_set_device(int ordinal) {
    g_device = ordinal;
    g_ready = false;
    return device_up() == NNA_SUCCESS ? 0 : (int)MARS_ERR_NNA_INIT_FAILED;
}
void mars_b200_set_arena_bytes(size_t bytes) { g_arena_bytes = bytes; }
const char *mars_b200_last_error(void) { return g_err; }

mars_error_t mars_b200_arena_upload(mars_model_t *model, int slot) {
    Model *m = as_model(model);
    if (!m) return MARS_ERR_INVALID_FILE;
    CU_OK(cudaSetDevice(m->device), MARS_ERR_NNA_INIT_FAILED);
    if (!range_ok(m, slot, 1)) return MARS_ERR_INVALID_TENSOR;
    /* the host mirror of the weights is writable (model->weights, tensor->vaddr, as in the reference): the tensor-core plans
     * hold a repacked copy and requantisation bounds derived from the blob, so an edit invalidates the compiled program */
    if (m->compiled && fnv1a(m->h_arena, m->weights_size) != m->weights_hash) { cudaStreamSynchronize(m->stream); m->compiled = false; }
    CU_OK(cudaMemcpyAsync(m->d_weights, m->h_arena, m->weights_size, cudaMemcpyHostToDevice, m->stream), MARS_ERR_LAYER_FAILED);
    size_t nbytes = std::min(m->slot_bytes, m->arena_size - m->weights_size);
    CU_OK(cudaMemcpyAsync(m->d_slots + (size_t)slot * m->slot_stride, m->h_arena + m->weights_size, nbytes,
                          cudaMemcpyHostToDevice, m->stream), MARS_ERR_LAYER_FAILED);
    CU_OK(cudaStreamSynchronize(m->stream), MARS_ERR_LAYER_FAILED);
    return compile_model(m);
}

mars_error_t mars_b200_arena_download(mars_model_t *model, int slot, void *host_dst, size_t bytes) {
    Model *m = as_model(model);
    if (!m) return MARS_ERR_INVALID_FILE;
    CU_OK(cudaSetDevice(m->device), MARS_ERR_NNA_INIT_FAILED);
    if (!range_ok(m, slot, 1)) return MARS_ERR_INVALID_TENSOR;
    uint8_t *dst = host_dst ? (uint8_t *)host_dst : m->h_arena;
    if (!host_dst || bytes > m->weights_size + m->slot_bytes) bytes = std::min(m->arena_size, m->weights_size + m->slot_bytes);
    size_t wb = std::min(bytes, m->weights_size);
    CU_OK(cudaMemcpyAsync(dst, m->d_weights, wb, cudaMemcpyDeviceToHost, m->stream), MARS_ERR_LAYER_FAILED);
    if (bytes > wb)
        CU_OK(cudaMemcpyAsync(dst + wb, m->d_slots + (size_t)slot * m->slot_stride, bytes - wb, cudaMemcpyDeviceToHost,
                              m->stream), MARS_ERR_LAYER_FAILED);
    CU_OK(cudaStreamSynchronize(m->stream), MARS_ERR_LAYER_FAILED);
    return MARS_OK;
}

mars_error_t mars_b200_arena_clear(mars_model_t *model) {
    Model *m = as_model(model);
    if (!m) return MARS_ERR_INVALID_FILE;
    CU_OK(cudaSetDevice(m->device), MARS_ERR_NNA_INIT_FAILED);
    CU_OK(cudaMemsetAsync(m->d_slots, 0, (size_t)m->capacity * m->slot_stride, m->stream), MARS_ERR_LAYER_FAILED);
    CU_OK(cudaStreamSynchronize(m->stream), MARS_ERR_LAYER_FAILED);
    memset(m->h_arena + m->weights_size, 0, m->arena_size - m->weights_size);
    return MARS_OK;
}

mars_error_t mars_b200_run_layer(mars_model_t *model, uint32_t layer) {
    Model *m = as_model(model);
    if (!m) return MARS_ERR_INVALID_FILE;
    if (layer >= model->header.num_layers) return MARS_ERR_INVALID_LAYER;
    CU_OK(cudaSetDevice(m->device), MARS_ERR_NNA_INIT_FAILED);
    mars_error_t e = enqueue_ops(m, 0, 1, (int)layer);
    cudaError_t ce = cudaStreamSynchronize(m->stream);
    if (e == MARS_OK && ce != cudaSuccess) {
        set_last_error("run_layer %u: %s", layer, cudaGetErrorString(ce));
        e = MARS_ERR_LAYER_FAILED;
    }
    return e;
}

size_t mars_b200_describe(mars_model_t *model, char *dst, size_t cap) {
    Model *m = as_model(model);
    if (!m) return 0;
    if (compile_model(m) != MARS_OK) return 0;
    std::string s = describe_program(m->prog);
    if (dst && cap) {
        size_t n = std::min(cap - 1, s.size());
        memcpy(dst, s.data(), n);
        dst[n] = 0;
    }
    return s.size();
}

void mars_b200_set_opt_level(mars_model_t *model, int level) {
    Model *m = as_model(model);
    if (!m || m->opt_level == level) return;
    cudaSetDevice(m->device);
    cudaStreamSynchronize(m->stream);
    m->opt_level = level;
    m->compiled = false;
}
void mars_b200_set_strict(int on) { g_strict = on ? 1 : 0; }
void mars_b200_set_f32_mode(mars_model_t *model, int mode) {
    Model *m = as_model(model);
    if (!m || m->f32_mode == mode) return;
    cudaSetDevice(m->device);
    cudaStreamSynchronize(m->stream);
    m->f32_mode = mode < 0 ? 0 : (mode > 2 ? 2 : mode);
    m->compiled = false;
}
void mars_b200_set_depthwise_mode(mars_model_t *model, int mode) {
    Model *m = as_model(model);
    if (!m || m->depthwise_mode == mode) return;
    cudaSetDevice(m->device);
    cudaStreamSynchronize(m->stream);
    m->depthwise_mode = mode;
    m->compiled = false;
}

mars_error_t mars_b200_set_batch(mars_model_t *model, int capacity) {
    Model *m = as_model(model);
    if (!m) return MARS_ERR_INVALID_FILE;
    CU_OK(cudaSetDevice(m->device), MARS_ERR_NNA_INIT_FAILED);
    mars_error_t e = set_capacity(m, capacity);
    if (e == MARS_OK) e = compile_model(m);
    return e;
}
int mars_b200_get_batch(mars_model_t *model) {
    Model *m = as_model(model);
    return m ? m->capacity : 0;
}

static size_t io_bytes(mars_model_t *model, int output) {
    mars_runtime_tensor_t *t = output ? mars_get_output(model, 0) : mars_get_input(model, 0);
    if (!t) return 0;
    size_t n = 1;
    for (uint32_t i = 0; i < t->desc.ndims && i < MARS_MAX_DIMS; i++) n *= (size_t)(t->desc.shape[i] < 0 ? 0 : t->desc.shape[i]);
    size_t es = (t->desc.dtype == MARS_DTYPE_FLOAT32 || t->desc.dtype == MARS_DTYPE_INT32) ? 4 : (t->desc.dtype == MARS_DTYPE_INT16 ? 2 : 1);
    return n * es;
}
size_t mars_b200_input_bytes(mars_model_t *model) { return io_bytes(model, 0); }
size_t mars_b200_output_bytes(mars_model_t *model) { return io_bytes(model, 1); }

static mars_error_t copy_io(Model *m, int first, int n, void *host, size_t stride, int output, cudaStream_t s) {
    mars_model_t *model = &m->pub;
    if (!range_ok(m, first, n)) return MARS_ERR_INVALID_TENSOR;
    uint32_t ti = output ? model->header.output_tensor_ids[0] : model->header.input_tensor_ids[0];
    if ((output ? model->header.num_outputs : model->header.num_inputs) < 1 || ti >= model->header.num_tensors ||
        m->toff[ti] < m->weights_size)
        return MARS_ERR_INVALID_TENSOR;
    size_t bytes = io_bytes(model, output);
    if (bytes > m->buffer_size) bytes = m->buffer_size;
    if (n == 0 || bytes == 0) return MARS_OK;
    if (stride < bytes) stride = bytes;
    uint8_t *dev = dev_addr(m, m->toff[ti], first);
    if (output)
        CU_OK(cudaMemcpy2DAsync(host, stride, dev, m->slot_stride, bytes, (size_t)n, cudaMemcpyDeviceToHost, s), MARS_ERR_LAYER_FAILED);
    else
        CU_OK(cudaMemcpy2DAsync(dev, m->slot_stride, host, stride, bytes, (size_t)n, cudaMemcpyHostToDevice, s), MARS_ERR_LAYER_FAILED);
    return MARS_OK;
}

mars_error_t mars_b200_upload_inputs(mars_model_t *model, int first, int n, const void *host, size_t stride) {
    Model *m = as_model(model);
    if (!m || !host) return MARS_ERR_INVALID_FILE;
    CU_OK(cudaSetDevice(m->device), MARS_ERR_NNA_INIT_FAILED);
    mars_error_t e = copy_io(m, first, n, const_cast<void *>(host), stride, 0, m->stream);
    if (e != MARS_OK) return e;
    CU_OK(cudaStreamSynchronize(m->stream), MARS_ERR_LAYER_FAILED);
    return MARS_OK;
}
mars_error_t mars_b200_download_outputs(mars_model_t *model, int first, int n, void *host, size_t stride) {
    Model *m = as_model(model);
    if (!m || !host) return MARS_ERR_INVALID_FILE;
    CU_OK(cudaSetDevice(m->device), MARS_ERR_NNA_INIT_FAILED);
    mars_error_t e = copy_io(m, first, n, host, stride, 1, m->stream);
    if (e != MARS_OK) return e;
    CU_OK(cudaStreamSynchronize(m->stream), MARS_ERR_LAYER_FAILED);
    return MARS_OK;
}

/* Pre-processing in front of mars_run, as the reference's caller does it on the CPU (src/mars/mars_yolo_test.c:40-77 and
 * :157-165): n RGB8 frames of w x h pixels (host memory, frame_stride bytes apart, >= w*h*3) are letterboxed into input
 * tensor 0 of slots [first, first+n) -- stbir_resize_uint8 semantics, gray border -17, px - 128, NCHW or NHWC as the
 * tensor's format field says.  Synchronous like mars_b200_upload_inputs. */
mars_error_t mars_b200_preprocess_batch(mars_model_t *model, int first, int n, const uint8_t *frames, size_t frame_stride, int w, int h) {
    Model *m = as_model(model);
    if (!m || !frames) return MARS_ERR_INVALID_FILE;
    CU_OK(cudaSetDevice(m->device), MARS_ERR_NNA_INIT_FAILED);
    if (!range_ok(m, first, n)) return MARS_ERR_INVALID_TENSOR;
    if (m->pub.header.num_inputs < 1) return MARS_ERR_INVALID_TENSOR;
    const uint32_t ti = m->pub.header.input_tensor_ids[0];
    if (ti >= m->pub.header.num_tensors || m->toff[ti] < m->weights_size) return MARS_ERR_INVALID_TENSOR;
    const mars_tensor_t &d = m->pub.tensors[ti].desc;
    if (d.dtype != MARS_DTYPE_INT8 || d.ndims < 4 || w <= 0 || h <= 0) {
        set_last_error("preprocess: input 0 must be a 4-d int8 tensor (reference src/mars/mars_yolo_test.c:157-159)");
        return MARS_ERR_INVALID_TENSOR;
    }
    const int nhwc = d.format == MARS_FORMAT_NHWC;
    const int th = nhwc ? d.shape[1] : d.shape[2], tw = nhwc ? d.shape[2] : d.shape[3];
    if (th <= 0 || tw <= 0 || (size_t)tw * th * 3 > m->buffer_size) return MARS_ERR_INVALID_TENSOR;
    const size_t frame_bytes = (size_t)w * h * 3;
    if (frame_stride < frame_bytes) frame_stride = frame_bytes;
    if (!letterbox_plan_build(w, h, tw, th, &m->pre_plan)) {
        set_last_error("preprocess: cannot letterbox %dx%d into %dx%d", w, h, tw, th);
        return MARS_ERR_INVALID_TENSOR;
    }
    const size_t h_per = (size_t)h * m->pre_plan.g.nw * 3 * sizeof(float);
    int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n, ((size_t)256 << 20) / std::max<size_t>(h_per, 1)));
    if ((size_t)chunk * frame_bytes > m->frames_cap) {
        cudaFree(m->d_frames);
        m->d_frames = nullptr; m->frames_cap = 0;
        CU_OK(cudaMalloc(&m->d_frames, (size_t)chunk * frame_bytes), MARS_ERR_ALLOC_FAILED);
        m->frames_cap = (size_t)chunk * frame_bytes;
    }
    if ((size_t)chunk * h_per > m->preH_cap) {
        cudaFree(m->d_preH);
        m->d_preH = nullptr; m->preH_cap = 0;
        CU_OK(cudaMalloc(&m->d_preH, (size_t)chunk * h_per), MARS_ERR_ALLOC_FAILED);
        m->preH_cap = (size_t)chunk * h_per;
    }
    cudaEventRecord(m->ev0, m->stream);
    for (int done = 0; done < n; done += chunk) {
        const int c = std::min(chunk, n - done);
        CU_OK(cudaMemcpy2DAsync(m->d_frames, frame_bytes, frames + (size_t)done * frame_stride, frame_stride, frame_bytes, (size_t)c,
                                cudaMemcpyHostToDevice, m->stream), MARS_ERR_LAYER_FAILED);
        CU_OK(launch_letterbox(m->pre_plan, m->d_frames, frame_bytes, c, m->d_preH,
                               reinterpret_cast<int8_t *>(dev_addr(m, m->toff[ti], first + done)), m->slot_stride, nhwc, m->stream),
              MARS_ERR_LAYER_FAILED);
        m->launches += 2;
    }
    cudaEventRecord(m->ev1, m->stream);
    CU_OK(cudaStreamSynchronize(m->stream), MARS_ERR_LAYER_FAILED);
    cudaEventElapsedTime(&m->last_ms, m->ev0, m->ev1);
    return MARS_OK;
}

/* Layer-group micro-batching (replaces the reference's one-image-at-a-time schedule, src/mars/mars_runtime.c:439-459): the
 * head of the op list -- the layers at the large resolutions, whose activations dominate the traffic -- runs over
 * MARS_MICRO_BATCH images at a time, all its layers back to back, so that a layer finds its producer's output in L2 instead
 * of HBM; the tail (small planes, where a micro-batch would not fill the 296 persistent CTAs) runs over the whole batch.
 * MARS_MICRO_OPS = number of head ops (default: everything in front of the first conv whose output plane is <= 80 x 80;
 * 0 = the whole list per micro-batch).  Images are independent, so any order across images is the same computation.
 * Off by default (MARS_MICRO_BATCH unset): profiles/r02h_micro_batch_sweep.txt has the measurement. */
static int micro_batch() {
    static int mb = -1;
    if (mb < 0) {
        const char *e = getenv("MARS_MICRO_BATCH");
        mb = (e && *e) ? atoi(e) : 0;
    }
    return mb;
}
static size_t micro_split(Model *m) {
    const char *e = getenv("MARS_MICRO_OPS");
    if (e && *e) return (size_t)std::max(0, atoi(e));
    for (size_t i = 0; i < m->prog.ops.size(); i++) {
        const Op &o = m->prog.ops[i];
        if ((o.kind == OP_CONV_I8_NCHW || o.kind == OP_CONV_I8_NHWC) && (long long)o.oh * o.ow <= 6400) return i;
    }
    return 0;
}

static mars_error_t enqueue_run(Model *m, int first, int n) {
    const int mb = micro_batch();
    if (mb <= 0 || mb >= n || m->profile) return enqueue_ops(m, first, n, -1);
    mars_error_t e = compile_model(m);
    if (e != MARS_OK) return e;
    const size_t nops = m->prog.ops.size(), split = micro_split(m);
    const size_t head = (split == 0 || split > nops) ? nops : split;
    for (int i = 0; i < n; i += mb) {
        e = enqueue_ops(m, first + i, std::min(mb, n - i), -1, 0, head);
        if (e != MARS_OK) return e;
    }
    if (head < nops) e = enqueue_ops(m, first, n, -1, head, nops);
    return e;
}

mars_error_t mars_b200_run_resident(mars_model_t *model, int first, int n) {
    Model *m = as_model(model);
    if (!m) return MARS_ERR_INVALID_FILE;
    CU_OK(cudaSetDevice(m->device), MARS_ERR_NNA_INIT_FAILED);
    if (!range_ok(m, first, n)) return MARS_ERR_INVALID_TENSOR;
    cudaEventRecord(m->ev0, m->stream);
    mars_error_t e = enqueue_run(m, first, n);
    cudaEventRecord(m->ev1, m->stream);
    cudaError_t ce = cudaStreamSynchronize(m->stream);
    if (e == MARS_OK && ce != cudaSuccess) {
        set_last_error("run_resident: %s", cudaGetErrorString(ce));
        e = MARS_ERR_LAYER_FAILED;
    }
    if (e == MARS_OK) { cudaEventElapsedTime(&m->last_ms, m->ev0, m->ev1); profile_collect(m); }
    return e;
}

mars_error_t mars_b200_detect_resident(mars_model_t *model, int first, int n, float nms_thresh) {
    Model *m = as_model(model);
    if (!m) return MARS_ERR_INVALID_FILE;
    CU_OK(cudaSetDevice(m->device), MARS_ERR_NNA_INIT_FAILED);
    if (!range_ok(m, first, n)) return MARS_ERR_INVALID_TENSOR;
    cudaEventRecord(m->ev0, m->stream);
    mars_error_t e = enqueue_detect(m, first, n, nms_thresh);
    cudaEventRecord(m->ev1, m->stream);
    cudaError_t ce = cudaStreamSynchronize(m->stream);
    if (e == MARS_OK && ce != cudaSuccess) {
        set_last_error("detect_resident: %s", cudaGetErrorString(ce));
        e = MARS_ERR_LAYER_FAILED;
    }
    if (e == MARS_OK) { cudaEventElapsedTime(&m->last_ms, m->ev0, m->ev1); profile_collect(m); }
    return e;
}

/* The ~110 launches of a step replayed as one CUDA graph: the second time a (first, n, threshold) combination is seen the
 * launch sequence is captured from the compute stream, afterwards it is replayed -- the kernels and their parameters are
 * the same, only the host-side launch work and the inter-kernel launch latency go.  MARS_GRAPH=0 disables it; profiling
 * (per-op events) and anything that cannot be captured fall back to plain launches. */
static bool graphs_enabled() {
    static const bool on = !(getenv("MARS_GRAPH") && atoi(getenv("MARS_GRAPH")) == 0);
    return on;
}

static mars_error_t enqueue_step(Model *m, int first, int n, float nms_thresh, int with_detect) {
    mars_error_t e = enqueue_run(m, first, n);
    if (e == MARS_OK && with_detect) e = enqueue_detect(m, first, n, nms_thresh);
    return e;
}

static mars_error_t enqueue_step_graphed(Model *m, int first, int n, float nms_thresh, int with_detect) {
    if (!graphs_enabled() || m->graphs_broken || m->profile || n <= 0) return enqueue_step(m, first, n, nms_thresh, with_detect);
    uint32_t tb;
    memcpy(&tb, &nms_thresh, 4);
    Model::StepGraph *g = nullptr;
    for (auto &c : m->graphs)
        if (c.first == first && c.n == n && c.with_detect == with_detect && c.thresh_bits == tb) g = &c;
    if (!g) {
        if (m->graphs.size() >= 16) release_graphs(m);
        m->graphs.push_back(Model::StepGraph{first, n, with_detect, tb, 0, nullptr, 0});
        g = &m->graphs.back();
    }
    if (g->exec) {
        if (cudaGraphLaunch(g->exec, m->stream) != cudaSuccess) { m->graphs_broken = true; return MARS_ERR_LAYER_FAILED; }
        m->launches += g->launches;
        return MARS_OK;
    }
    if (++g->seen < 2) return enqueue_step(m, first, n, nms_thresh, with_detect); /* first time: plain (also does the lazy one-time setup) */
    const uint64_t before = m->launches;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        m->graphs_broken = true;
        cudaGetLastError();
        return enqueue_step(m, first, n, nms_thresh, with_detect);
    }
    mars_error_t e = enqueue_step(m, first, n, nms_thresh, with_detect);
    const cudaError_t ce = cudaStreamEndCapture(m->stream, &graph);
    if (e != MARS_OK || ce != cudaSuccess || !graph || cudaGraphInstantiate(&g->exec, graph, 0) != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        g->exec = nullptr;
        m->graphs_broken = true;
        cudaGetLastError();
        m->launches = before;
        return e != MARS_OK ? e : enqueue_step(m, first, n, nms_thresh, with_detect);
    }
    cudaGraphDestroy(graph);
    g->launches = m->launches - before;
    if (cudaGraphLaunch(g->exec, m->stream) != cudaSuccess) { m->graphs_broken = true; return MARS_ERR_LAYER_FAILED; }
    return MARS_OK;
}

/* run + detect as one timed region (the bench's resident step) */
mars_error_t mars_b200_step_resident(mars_model_t *model, int first, int n, float nms_thresh, int with_detect) {
    Model *m = as_model(model);
    if (!m) return MARS_ERR_INVALID_FILE;
    CU_OK(cudaSetDevice(m->device), MARS_ERR_NNA_INIT_FAILED);
    if (!range_ok(m, first, n)) return MARS_ERR_INVALID_TENSOR;
    mars_error_t e = compile_model(m);
    if (e != MARS_OK) return e;
    cudaEventRecord(m->ev0, m->stream);
    e = enqueue_step_graphed(m, first, n, nms_thresh, with_detect);
    cudaEventRecord(m->ev1, m->stream);
    cudaError_t ce = cudaStreamSynchronize(m->stream);
    if (e == MARS_OK && ce != cudaSuccess) {
        set_last_error("step_resident: %s", cudaGetErrorString(ce));
        e = MARS_ERR_LAYER_FAILED;
    }
    if (e == MARS_OK) { cudaEventElapsedTime(&m->last_ms, m->ev0, m->ev1); profile_collect(m); }
    return e;
}

/* the same step, enqueued only: the call returns as soon as the launches (one CUDA-graph replay from the second time on) are on the
 * model's compute stream; the caller orders its own work behind them there (mars_b200_compute_stream) and synchronises itself */
mars_error_t mars_b200_enqueue_step_resident(mars_model_t *model, int first, int n, float nms_thresh, int with_detect) {
    Model *m = as_model(model);
    if (!m) return MARS_ERR_INVALID_FILE;
    CU_OK(cudaSetDevice(m->device), MARS_ERR_NNA_INIT_FAILED);
    if (!range_ok(m, first, n)) return MARS_ERR_INVALID_TENSOR;
    mars_error_t e = compile_model(m);
    if (e != MARS_OK) return e;
    return enqueue_step_graphed(m, first, n, nms_thresh, with_detect);
}

static mars_error_t copy_dets(Model *m, int first, int n, mars_det_t *dets, int32_t *counts, int maxd, cudaStream_t s) {
    if (maxd > 1000) maxd = 1000;
    if (maxd < 0) maxd = 0;
    CU_OK(cudaMemcpyAsync(counts, m->d_det_cnt + first, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s), MARS_ERR_LAYER_FAILED);
    if (maxd > 0)
        CU_OK(cudaMemcpy2DAsync(dets, (size_t)maxd * sizeof(mars_det_t), m->d_det + (size_t)first * MARS_MAX_DETS,
                                (size_t)MARS_MAX_DETS * sizeof(mars_det_t), (size_t)maxd * sizeof(mars_det_t), (size_t)n,
                                cudaMemcpyDeviceToHost, s), MARS_ERR_LAYER_FAILED);
    return MARS_OK;
}

mars_error_t mars_b200_download_detections(mars_model_t *model, int first, int n, mars_det_t *dets, int32_t *counts, int maxd) {
    Model *m = as_model(model);
    if (!m || !dets || !counts) return MARS_ERR_INVALID_FILE;
    CU_OK(cudaSetDevice(m->device), MARS_ERR_NNA_INIT_FAILED);
    if (!range_ok(m, first, n)) return MARS_ERR_INVALID_TENSOR;
    mars_error_t e = copy_dets(m, first, n, dets, counts, maxd, m->stream);
    if (e != MARS_OK) return e;
    CU_OK(cudaStreamSynchronize(m->stream), MARS_ERR_LAYER_FAILED);
    for (int i = 0; i < n; i++) if (counts[i] > maxd) counts[i] = maxd;
    return MARS_OK;
}

/* end to end from host buffers, double-buffered over the two halves of the slot pool:
 * H2D of chunk c+1 and D2H of chunk c-1 overlap the kernels of chunk c */
static mars_error_t batch_pipeline(Model *m, int n, const void *inputs, size_t in_stride, void *outputs, size_t out_stride,
                                   mars_det_t *dets, int32_t *counts, int maxd, float thresh) {
    if (n <= 0) return MARS_OK;
    const size_t in_bytes = io_bytes(&m->pub, 0);
    if (in_stride < in_bytes) in_stride = in_bytes;
    const size_t out_bytes = io_bytes(&m->pub, 1);
    if (out_stride < out_bytes) out_stride = out_bytes;
    if (maxd > 1000) maxd = 1000;
    const int chunk = m->capacity >= 2 ? m->capacity / 2 : 1;
    const int halves = m->capacity >= 2 ? 2 : 1;
    mars_error_t e = MARS_OK;
    cudaEventRecord(m->ev0, m->stream);
    int c = 0;
    for (int done = 0; done < n && e == MARS_OK; done += chunk, c++) {
        const int h = c % halves, first = h * chunk, cnt = std::min(chunk, n - done);
        /* the half is free once its previous results were copied out */
        if (c >= halves) cudaStreamWaitEvent(m->h2d_stream, m->ev_out[h], 0);
        e = copy_io(m, first, cnt, (uint8_t *)const_cast<void *>(inputs) + (size_t)done * in_stride, in_stride, 0, m->h2d_stream);
        if (e != MARS_OK) break;
        cudaEventRecord(m->ev_in[h], m->h2d_stream);
        cudaStreamWaitEvent(m->stream, m->ev_in[h], 0);
        e = enqueue_run(m, first, cnt);
        if (e == MARS_OK && dets) e = enqueue_detect(m, first, cnt, thresh);
        if (e != MARS_OK) break;
        cudaEventRecord(m->ev_done[h], m->stream);
        cudaStreamWaitEvent(m->d2h_stream, m->ev_done[h], 0);
        if (dets) e = copy_dets(m, first, cnt, dets + (size_t)done * maxd, counts + done, maxd, m->d2h_stream);
        if (e == MARS_OK && outputs) e = copy_io(m, first, cnt, (uint8_t *)outputs + (size_t)done * out_stride, out_stride, 1, m->d2h_stream);
        cudaEventRecord(m->ev_out[h], m->d2h_stream);
    }
    cudaEventRecord(m->ev1, m->stream);
    cudaError_t c1 = cudaStreamSynchronize(m->h2d_stream), c2 = cudaStreamSynchronize(m->stream), c3 = cudaStreamSynchronize(m->d2h_stream);
    if (e == MARS_OK && (c1 != cudaSuccess || c2 != cudaSuccess || c3 != cudaSuccess)) {
        set_last_error("batch pipeline: %s", cudaGetErrorString(c1 != cudaSuccess ? c1 : (c2 != cudaSuccess ? c2 : c3)));
        e = MARS_ERR_LAYER_FAILED;
    }
    if (e == MARS_OK) {
        cudaEventElapsedTime(&m->last_ms, m->ev0, m->ev1);
        if (dets) for (int i = 0; i < n; i++) if (counts[i] > maxd) counts[i] = maxd;
    }
    return e;
}

mars_error_t mars_b200_detect_batch(mars_model_t *model, int n, const void *inputs, size_t in_stride, mars_det_t *dets,
                                    int32_t *counts, int maxd, float nms_thresh) {
    Model *m = as_model(model);
    if (!m || !inputs || !dets || !counts) return MARS_ERR_INVALID_FILE;
    CU_OK(cudaSetDevice(m->device), MARS_ERR_NNA_INIT_FAILED);
    return batch_pipeline(m, n, inputs, in_stride, nullptr, 0, dets, counts, maxd, nms_thresh);
}

mars_error_t mars_b200_run_batch(mars_model_t *model, int n, const void *inputs, size_t in_stride, void *outputs, size_t out_stride) {
    Model *m = as_model(model);
    if (!m || !inputs || !outputs) return MARS_ERR_INVALID_FILE;
    CU_OK(cudaSetDevice(m->device), MARS_ERR_NNA_INIT_FAILED);
    return batch_pipeline(m, n, inputs, in_stride, outputs, out_stride, nullptr, nullptr, 0, 0.0f);
}

/* Asynchronous form of mars_b200_detect_batch: the slot pool is split into two halves ("pools" 0 and 1); a batch of up to
 * capacity/2 images is queued on one half -- H2D on the copy stream, layers + decode + NMS on the compute stream, D2H of the
 * detection records on the read-back stream -- and the call returns.  While it runs the caller submits the next batch to the
 * other half, so the host->device copy of batch k+1 and the read-back of batch k-1 overlap the kernels of batch k.
 * inputs / dets / counts must stay valid (and should be pinned, e.g. nna_malloc) until mars_b200_wait_batch(pool) returns. */
mars_error_t mars_b200_submit_batch(mars_model_t *model, int pool, int n, const void *inputs, size_t in_stride, mars_det_t *dets,
                                    int32_t *counts, int maxd, float nms_thresh) {
    Model *m = as_model(model);
    if (!m || !inputs || !dets || !counts || pool < 0 || pool > 1) return MARS_ERR_INVALID_FILE;
    CU_OK(cudaSetDevice(m->device), MARS_ERR_NNA_INIT_FAILED);
    const int half = m->capacity / 2;
    if (half < 1 || n < 0 || n > half) {
        set_last_error("submit_batch: %d images do not fit half of the slot pool (capacity %d; mars_b200_set_batch)", n, m->capacity);
        return MARS_ERR_INVALID_TENSOR;
    }
    if (m->pending[pool].active) {
        set_last_error("submit_batch: pool %d still has a batch in flight (mars_b200_wait_batch first)", pool);
        return MARS_ERR_LAYER_FAILED;
    }
    if (maxd > 1000) maxd = 1000;
    if (maxd < 0) maxd = 0;
    const size_t in_bytes = io_bytes(&m->pub, 0);
    if (in_stride < in_bytes) in_stride = in_bytes;
    const int first = pool * half;
    mars_error_t e = compile_model(m); /* also drops captured graphs of an older program (set_opt_level, set_depthwise_mode) */
    if (e != MARS_OK) return e;
    if (n > 0) {
        /* the half is free: its previous batch was waited for (read-back complete) before this call */
        e = copy_io(m, first, n, const_cast<void *>(inputs), in_stride, 0, m->h2d_stream);
        if (e != MARS_OK) return e;
        cudaEventRecord(m->ev_in[pool], m->h2d_stream);
        cudaStreamWaitEvent(m->stream, m->ev_in[pool], 0);
        e = enqueue_step_graphed(m, first, n, nms_thresh, 1);
        if (e != MARS_OK) { cudaStreamSynchronize(m->stream); return e; }
        cudaEventRecord(m->ev_done[pool], m->stream);
        cudaStreamWaitEvent(m->d2h_stream, m->ev_done[pool], 0);
        e = copy_dets(m, first, n, dets, counts, maxd, m->d2h_stream);
        if (e != MARS_OK) return e;
    }
    cudaEventRecord(m->ev_out[pool], m->d2h_stream);
    m->pending[pool].active = true;
    m->pending[pool].n = n; m->pending[pool].maxd = maxd; m->pending[pool].counts = counts;
    return MARS_OK;
}

/* mars_b200_submit_batch without the YOLO post-process: the raw output tensor 0 of every image is read back instead
 * (models whose output is not a [1, N, 85] head: NanoDet-style heads, float32 models).  Same pools, same wait. */
mars_error_t mars_b200_submit_run_batch(mars_model_t *model, int pool, int n, const void *inputs, size_t in_stride, void *outputs, size_t out_stride) {
    Model *m = as_model(model);
    if (!m || !inputs || !outputs || pool < 0 || pool > 1) return MARS_ERR_INVALID_FILE;
    CU_OK(cudaSetDevice(m->device), MARS_ERR_NNA_INIT_FAILED);
    const int half = m->capacity / 2;
    if (half < 1 || n < 0 || n > half) {
        set_last_error("submit_run_batch: %d images do not fit half of the slot pool (capacity %d; mars_b200_set_batch)", n, m->capacity);
        return MARS_ERR_INVALID_TENSOR;
    }
    if (m->pending[pool].active) {
        set_last_error("submit_run_batch: pool %d still has a batch in flight (mars_b200_wait_batch first)", pool);
        return MARS_ERR_LAYER_FAILED;
    }
    const size_t in_bytes = io_bytes(&m->pub, 0), out_bytes = io_bytes(&m->pub, 1);
    if (in_stride < in_bytes) in_stride = in_bytes;
    if (out_stride < out_bytes) out_stride = out_bytes;
    const int first = pool * half;
    mars_error_t e = compile_model(m);
    if (e != MARS_OK) return e;
    if (n > 0) {
        e = copy_io(m, first, n, const_cast<void *>(inputs), in_stride, 0, m->h2d_stream);
        if (e != MARS_OK) return e;
        cudaEventRecord(m->ev_in[pool], m->h2d_stream);
        cudaStreamWaitEvent(m->stream, m->ev_in[pool], 0);
        e = enqueue_step_graphed(m, first, n, 0.0f, 0);
        if (e != MARS_OK) { cudaStreamSynchronize(m->stream); return e; }
        cudaEventRecord(m->ev_done[pool], m->stream);
        cudaStreamWaitEvent(m->d2h_stream, m->ev_done[pool], 0);
        e = copy_io(m, first, n, outputs, out_stride, 1, m->d2h_stream);
        if (e != MARS_OK) return e;
    }
    cudaEventRecord(m->ev_out[pool], m->d2h_stream);
    m->pending[pool].active = true;
    m->pending[pool].n = 0; m->pending[pool].maxd = 0; m->pending[pool].counts = nullptr;
    return MARS_OK;
}

mars_error_t mars_b200_wait_batch(mars_model_t *model, int pool) {
    Model *m = as_model(model);
    if (!m || pool < 0 || pool > 1) return MARS_ERR_INVALID_FILE;
    if (!m->pending[pool].active) return MARS_OK;
    CU_OK(cudaSetDevice(m->device), MARS_ERR_NNA_INIT_FAILED);
    cudaError_t ce = cudaEventSynchronize(m->ev_out[pool]);
    Model::Pending &pd = m->pending[pool];
    pd.active = false;
    if (ce != cudaSuccess) {
        set_last_error("wait_batch: %s", cudaGetErrorString(ce));
        return MARS_ERR_LAYER_FAILED;
    }
    for (int i = 0; i < pd.n; i++) if (pd.counts[i] > pd.maxd) pd.counts[i] = pd.maxd;
    return MARS_OK;
}

/* the CUDA stream (cudaStream_t) every kernel of this model is launched on: work a caller enqueues there -- e.g. an NCCL
 * gather of the resident detection records -- is ordered after the batch that produced them and before the next one */
void *mars_b200_compute_stream(mars_model_t *model) {
    Model *m = as_model(model);
    return m ? (void *)m->stream : nullptr;
}
uint64_t mars_b200_launch_count(mars_model_t *model) {
    Model *m = as_model(model);
    return m ? m->launches : 0;
}
/* device addresses of the resident detection records (for a device-side gather):
 * dets = mars_det_t[capacity][1024], counts = int32[capacity] */
void mars_b200_detections_device(mars_model_t *model, void **dets, void **counts, int *stride_dets) {
    Model *m = as_model(model);
    if (dets) *dets = m ? m->d_det : nullptr;
    if (counts) *counts = m ? m->d_det_cnt : nullptr;
    if (stride_dets) *stride_dets = MARS_MAX_DETS;
}
float mars_b200_last_gpu_ms(mars_model_t *model) {
    Model *m = as_model(model);
    return m ? m->last_ms : 0.0f;
}

/* per-op profile: enable, run, read back "op kind layer ms calls bytes macs" lines */
void mars_b200_set_profile(mars_model_t *model, int on) {
    Model *m = as_model(model);
    if (!m) return;
    m->profile = on;
    if (on) {
        std::fill(m->prof_ms.begin(), m->prof_ms.end(), 0.0);
        std::fill(m->prof_calls.begin(), m->prof_calls.end(), 0);
    }
}
int mars_b200_num_ops(mars_model_t *model) {
    Model *m = as_model(model);
    if (!m || compile_model(m) != MARS_OK) return 0;
    return (int)m->prog.ops.size();
}
/* info[0]=kind [1]=layer [2]=impl [3]=mode [4]=ic [5]=oc [6]=oh [7]=ow [8]=kh [9]=kw [10]=fused_layers [11]=ih [12]=iw;
 * ms = accumulated CUDA-event milliseconds, calls = timed launches of this op */
int mars_b200_op_info(mars_model_t *model, int op, int32_t *info, double *ms, uint64_t *calls, uint64_t *flat_n) {
    Model *m = as_model(model);
    if (!m || op < 0 || (size_t)op >= m->prog.ops.size()) return -1;
    const Op &o = m->prog.ops[op];
    if (info) {
        info[0] = o.kind; info[1] = o.layer; info[2] = o.impl; info[3] = o.mode; info[4] = o.ic; info[5] = o.oc;
        info[6] = o.oh; info[7] = o.ow; info[8] = o.kh; info[9] = o.kw; info[10] = o.fused_layers; info[11] = o.ih; info[12] = o.iw;
    }
    if (ms) *ms = (size_t)op < m->prof_ms.size() ? m->prof_ms[op] : 0.0;
    if (calls) *calls = (size_t)op < m->prof_calls.size() ? m->prof_calls[op] : 0;
    if (flat_n) *flat_n = o.n;
    return 0;
}

/* geometry of the arena, for tests: out[0]=weights_size [1]=buffer_size [2]=num_buffers [3]=slot_stride [4]=arena_size */
void mars_b200_geometry(mars_model_t *model, size_t *out5) {
    Model *m = as_model(model);
    if (!m || !out5) return;
    out5[0] = m->weights_size; out5[1] = m->buffer_size; out5[2] = (size_t)m->num_buffers; out5[3] = m->slot_stride; out5[4] = m->arena_size;
}
size_t mars_b200_tensor_offset(mars_model_t *model, uint32_t index) {
    Model *m = as_model(model);
    if (!m || index >= model->header.num_tensors) return (size_t)-1;
    return m->toff[index];
}

/* ---- several GPUs behind one C call (north_star: "work is partitioned across the GPUs of one box by sharding the
 * image batch, with no inter-GPU traffic on the hot path") --------------------------------------------------------------
 * A group holds one replica of the model per device.  mars_b200_group_detect_batch shards the n images into contiguous
 * blocks of ceil(n / G), runs every block on its GPU from its own host thread (H2D, all layers, decode, NMS, D2H, pipelined
 * in chunks exactly like mars_b200_detect_batch) and returns when all are done; the detection records land directly in the
 * caller's arrays, which is the "gather" of a single-process caller (bench.py, one process per GPU, gathers over NCCL). */
struct mars_b200_group {
    std::vector<mars_model_t *> models;
    std::vector<int> devices;
};

mars_error_t mars_b200_group_load(const void *data, size_t size, const int *devices, int n_devices, int per_gpu_batch, mars_b200_group_t **out) {
    if (!data || !out || n_devices < 1 || per_gpu_batch < 1) return MARS_ERR_INVALID_FILE;
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1) {
        set_last_error("no CUDA device (this library has no CPU fallback)");
        return MARS_ERR_NNA_INIT_FAILED;
    }
    mars_b200_group *g = new (std::nothrow) mars_b200_group();
    if (!g) return MARS_ERR_ALLOC_FAILED;
    const int saved = g_device;
    mars_error_t e = MARS_OK;
    for (int i = 0; i < n_devices && e == MARS_OK; i++) {
        const int d = devices ? devices[i] : i;
        if (d < 0 || d >= visible) { set_last_error("group: device ordinal %d out of range (%d visible)", d, visible); e = MARS_ERR_NNA_INIT_FAILED; break; }
        if (mars_b200_set_device(d) != 0) { e = MARS_ERR_NNA_INIT_FAILED; break; }
        mars_model_t *m = nullptr;
        e = mars_load_memory(data, size, &m);
        if (e == MARS_OK) e = mars_b200_set_batch(m, per_gpu_batch);
        if (e != MARS_OK) { if (m) mars_free(m); break; }
        g->models.push_back(m);
        g->devices.push_back(d);
    }
    if (saved >= 0) mars_b200_set_device(saved);
    if (e != MARS_OK) {
        for (auto *m : g->models) mars_free(m);
        delete g;
        return e;
    }
    *out = g;
    return MARS_OK;
}

void mars_b200_group_free(mars_b200_group_t *g) {
    if (!g) return;
    for (auto *m : g->models) mars_free(m);
    delete g;
}

int mars_b200_group_size(mars_b200_group_t *g) { return g ? (int)g->models.size() : 0; }
mars_model_t *mars_b200_group_model(mars_b200_group_t *g, int i) { return (g && i >= 0 && (size_t)i < g->models.size()) ? g->models[i] : nullptr; }

mars_error_t mars_b200_group_detect_batch(mars_b200_group_t *g, int n, const void *inputs, size_t in_stride, mars_det_t *dets, int32_t *counts,
                                          int maxd, float nms_thresh) {
    if (!g || g->models.empty() || !inputs || !dets || !counts || n < 0) return MARS_ERR_INVALID_FILE;
    const int G = (int)g->models.size();
    const size_t in_bytes = mars_b200_input_bytes(g->models[0]);
    if (in_stride < in_bytes) in_stride = in_bytes;
    if (maxd > 1000) maxd = 1000;
    if (maxd < 0) maxd = 0;
    const int per = (n + G - 1) / G;
    std::vector<mars_error_t> err((size_t)G, MARS_OK);
    std::vector<std::string> why((size_t)G);
    std::vector<std::thread> th;
    for (int r = 0; r < G; r++) {
        const int first = std::min(r * per, n), cnt = std::max(0, std::min(per, n - first));
        if (cnt == 0) continue;
        th.emplace_back([=, &err, &why]() {
            err[r] = mars_b200_detect_batch(g->models[r], cnt, (const uint8_t *)inputs + (size_t)first * in_stride, in_stride,
                                            dets + (size_t)first * maxd, counts + first, maxd, nms_thresh);
            if (err[r] != MARS_OK) why[r] = g_err; /* the diagnostic is thread-local: carry it to the caller's thread */
        });
    }
    for (auto &t : th) t.join();
    for (int r = 0; r < G; r++)
        if (err[r] != MARS_OK) { set_last_error("group: GPU %d: %s", g->devices[r], why[r].c_str()); return err[r]; }
    return MARS_OK;
}

} /* extern "C" */

#include "seam.inc"
