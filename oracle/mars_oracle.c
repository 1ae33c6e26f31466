/*
 * oracle/mars_oracle.c -- TEST INFRASTRUCTURE: CPU restatement of the reference's
 * `mars` hot path (loader/planner, layer executor, YOLO decode + NMS).
 *
 * This file is the parity CHECKER.  It is never linked into, imported by or
 * executed from the product library; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Parity pin: every function below is validated byte-for-byte (whole arena)
 * against oracle/_ref/libmars_ref.so -- the reference's own sources compiled by
 * oracle/build_ref.sh -- on all shipped models by tests/test_oracle.py,
 * and against the committed fixtures in tests/golden/ (generated from that
 * binary by tests/golden/make_golden.py).  The reference's only known-answer test
 * on this path (examples/mars_math_test.c:38-82) is restated in
 * tests/test_oracle.py::test_mars_math_known_answers.  Two functions have NO reference implementation and
 * are labelled "restatement, parity unpinned": mo_depthwise_* (the reference's
 * depthwise layer is a no-op, src/mars/mars_runtime.c:1168-1170) and
 * mo_decode_anchor_grid (C stub at examples/yolo_detect.cpp:184-205; formula from
 * mgk-decompiler/test_yolo_inference.py:136-202).
 *
 * All file:line citations are relative to /root/reference.
 * The structure deliberately differs from the reference (one flat arena addressed
 * by offsets, row-vectorised conv with an explicit hazard test, closed-form
 * exchange-sort passes) -- what must be equal is every byte it produces.
 */
#include "mars_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MO_ALIGN_UP(x, a) (((x) + (a)-1) & ~((size_t)(a)-1))

/* ------------------------------------------------------------------------- */
/* arithmetic contracts                                                        */
/* ------------------------------------------------------------------------- */

/* (int32_t)f as x86-64 `cvttss2si` performs it: out-of-range and NaN give INT_MIN
 * (observable in shipped models).  Written out explicitly so
 * the oracle does not depend on the host compiler's UB behaviour. */
static inline int32_t mo_f2i_x86(float v) {
    if (!(v > -2147483904.0f && v < 2147483648.0f)) return INT32_MIN; /* NaN falls here too */
    return (int32_t)v;
}

static inline int8_t mo_clamp_i8(int32_t r) { return (int8_t)(r > 127 ? 127 : (r < -128 ? -128 : r)); }

/* conv requantisation: src/mars/mxu_conv.c:663-666 (NCHW), :750-753 (NHWC) */
static inline int8_t mo_requant_conv(int32_t acc, float cs) {
    volatile float scaled = (float)acc * cs; /* volatile: one rounding per operation, no contraction */
    volatile float biased = scaled + (scaled >= 0 ? 0.5f : -0.5f);
    return mo_clamp_i8(mo_f2i_x86(biased));
}

/* eltwise requantisation "trunc(y*inv + 0.5)": src/mars/mars_runtime.c:831,898 */
static inline int8_t mo_requant_mul_inv(float y, float inv) {
    volatile float t = y * inv;
    volatile float u = t + 0.5f;
    return mo_clamp_i8(mo_f2i_x86(u));
}

/* eltwise requantisation "trunc(y/scale + 0.5)": src/mars/mars_runtime.c:764,1147 */
static inline int8_t mo_requant_div(float y, float scale) {
    volatile float t = y / scale;
    volatile float u = t + 0.5f;
    return mo_clamp_i8(mo_f2i_x86(u));
}

/* ------------------------------------------------------------------------- */
/* model object                                                                */
/* ------------------------------------------------------------------------- */

struct mo_model {
    mars_header_t header;
    mars_tensor_t *tensors;
    mars_layer_t *layers;
    size_t *toff;    /* arena offset of each tensor's base address */
    size_t *talloc;  /* alloc_size the reference would report */
    uint8_t *arena;  /* [weights | buf0 | buf1 | (buf2)] */
    size_t arena_size;
    size_t weights_size;
    size_t buffer_size;
    int num_buffers;
    int depthwise_mode; /* 0 = reference (no-op), 1 = restated depthwise */
};

/* src/mars/mars_runtime.c:80-124 */
size_t mo_tensor_byte_size(const mars_tensor_t *t) {
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

/* planner: src/mars/mars_runtime.c:248-337 */
int mo_load(const void *blob, size_t size, size_t arena_bytes, mo_model_t **out) {
    if (!blob || !out || size < sizeof(mars_header_t)) return MO_ERR_INVALID_FILE;
    mars_header_t h;
    memcpy(&h, blob, sizeof h);
    if (h.magic != MARS_MAGIC) return MO_ERR_INVALID_MAGIC;
    if (h.version_major != MARS_VERSION_MAJOR) return MO_ERR_VERSION_MISMATCH;
    size_t need = sizeof h + (size_t)h.num_tensors * sizeof(mars_tensor_t) + (size_t)h.num_layers * sizeof(mars_layer_t);
    if (size < need) return MO_ERR_INVALID_FILE;

    mo_model_t *m = (mo_model_t *)calloc(1, sizeof *m);
    if (!m) return MO_ERR_ALLOC_FAILED;
    m->header = h;
    m->tensors = (mars_tensor_t *)calloc(h.num_tensors ? h.num_tensors : 1, sizeof(mars_tensor_t));
    m->layers = (mars_layer_t *)calloc(h.num_layers ? h.num_layers : 1, sizeof(mars_layer_t));
    m->toff = (size_t *)calloc(h.num_tensors ? h.num_tensors : 1, sizeof(size_t));
    m->talloc = (size_t *)calloc(h.num_tensors ? h.num_tensors : 1, sizeof(size_t));
    const uint8_t *p = (const uint8_t *)blob + sizeof h;
    memcpy(m->tensors, p, (size_t)h.num_tensors * sizeof(mars_tensor_t));
    p += (size_t)h.num_tensors * sizeof(mars_tensor_t);
    memcpy(m->layers, p, (size_t)h.num_layers * sizeof(mars_layer_t));

    m->arena_size = arena_bytes ? arena_bytes : ((size_t)8 << 20);
    if (h.weights_size > m->arena_size || h.weights_offset + h.weights_size > size) {
        mo_free(m);
        return MO_ERR_ALLOC_FAILED;
    }
    /* slack as in oracle/ref_shim.c so identical over-reads stay in mapped memory */
    m->arena = (uint8_t *)calloc(1, m->arena_size + 4096);
    if (!m->arena) { mo_free(m); return MO_ERR_ALLOC_FAILED; }
    m->weights_size = h.weights_size;
    memcpy(m->arena, (const uint8_t *)blob + h.weights_offset, h.weights_size);

    size_t remaining = m->arena_size - m->weights_size, maxsz = 0;
    for (uint32_t i = 0; i < h.num_tensors; i++)
        if (m->tensors[i].data_size == 0) {
            size_t sz = MO_ALIGN_UP(mo_tensor_byte_size(&m->tensors[i]), 64);
            if (sz > maxsz) maxsz = sz;
        }
    size_t nb = 3, bs = maxsz;
    if (bs * nb > remaining) nb = 2;
    if (bs * nb > remaining) {
        bs = (remaining / 2) & ~(size_t)63;
        if (bs < 65536) { mo_free(m); return MO_ERR_ALLOC_FAILED; }
    }
    m->num_buffers = (int)nb;
    m->buffer_size = bs;
    uint32_t k = 0;
    for (uint32_t i = 0; i < h.num_tensors; i++) {
        if (m->tensors[i].data_size > 0) {
            m->toff[i] = (size_t)m->tensors[i].data_offset;
            m->talloc[i] = (size_t)m->tensors[i].data_size;
        } else {
            m->toff[i] = m->weights_size + (size_t)(k % nb) * bs;
            m->talloc[i] = bs;
            k++;
        }
    }
    *out = m;
    return MO_OK;
}

void mo_free(mo_model_t *m) {
    if (!m) return;
    free(m->tensors); free(m->layers); free(m->toff); free(m->talloc); free(m->arena);
    free(m);
}

uint8_t *mo_arena(mo_model_t *m) { return m->arena; }
size_t mo_arena_size(const mo_model_t *m) { return m->arena_size; }
size_t mo_weights_size(const mo_model_t *m) { return m->weights_size; }
size_t mo_buffer_size(const mo_model_t *m) { return m->buffer_size; }
int mo_num_buffers(const mo_model_t *m) { return m->num_buffers; }
uint32_t mo_num_layers(const mo_model_t *m) { return m->header.num_layers; }
uint32_t mo_num_tensors(const mo_model_t *m) { return m->header.num_tensors; }
const mars_tensor_t *mo_tensor_desc(const mo_model_t *m, uint32_t idx) { return idx < m->header.num_tensors ? &m->tensors[idx] : NULL; }
const mars_layer_t *mo_layer_desc(const mo_model_t *m, uint32_t idx) { return idx < m->header.num_layers ? &m->layers[idx] : NULL; }
size_t mo_tensor_offset(const mo_model_t *m, uint32_t idx) { return idx < m->header.num_tensors ? m->toff[idx] : (size_t)-1; }
size_t mo_tensor_alloc(const mo_model_t *m, uint32_t idx) { return idx < m->header.num_tensors ? m->talloc[idx] : 0; }
void mo_set_depthwise_mode(mo_model_t *m, int mode) { m->depthwise_mode = mode; }
int mo_input_index(const mo_model_t *m, int i) {
    if (i < 0 || (uint32_t)i >= m->header.num_inputs) return -1;
    uint32_t tid = m->header.input_tensor_ids[i];
    return tid < m->header.num_tensors ? (int)tid : -1; /* index, not id: src/mars/mars_runtime.c:399-401 */
}
int mo_output_index(const mo_model_t *m, int i) {
    if (i < 0 || (uint32_t)i >= m->header.num_outputs) return -1;
    uint32_t tid = m->header.output_tensor_ids[i];
    return tid < m->header.num_tensors ? (int)tid : -1;
}

/* first table entry whose desc.id matches: src/mars/mars_runtime.c:713-721 */
static int mo_find(const mo_model_t *m, uint32_t id) {
    if (id == 0xFFFFFFFFu) return -1;
    for (uint32_t i = 0; i < m->header.num_tensors; i++)
        if (m->tensors[i].id == id) return (int)i;
    return -1;
}

static size_t mo_numel(const mars_tensor_t *t) {
    size_t n = 1;
    for (uint32_t i = 0; i < t->ndims && i < MARS_MAX_DIMS; i++) n *= (size_t)(t->shape[i] < 0 ? 0 : t->shape[i]);
    return n;
}

static int ranges_overlap(size_t a0, size_t a1, size_t b0, size_t b1) { return a0 < b1 && b0 < a1 && a0 < a1 && b0 < b1; }

/* ------------------------------------------------------------------------- */
/* convolution                                                                 */
/* ------------------------------------------------------------------------- */

typedef struct {
    int in_c, in_h, in_w, out_c, out_h, out_w, kh, kw, sh, sw, pt, pl;
} mo_conv_dims;

/* literal loop nest of src/mars/mxu_conv.c:642-669; used whenever the output
 * range overlaps the input range (in-place layers) */
static void conv_i8_nchw_literal(uint8_t *A, size_t in, size_t w, const int32_t *bias, size_t out,
                                 const mo_conv_dims *d, float cs) {
    const int8_t *input = (const int8_t *)(A + in), *weight = (const int8_t *)(A + w);
    int8_t *output = (int8_t *)(A + out);
    int wpo = d->in_c * d->kh * d->kw;
    for (int oc = 0; oc < d->out_c; oc++) {
        const int8_t *wo = weight + (size_t)oc * wpo;
        int32_t b;
        if (bias) memcpy(&b, (const uint8_t *)bias + 4 * (size_t)oc, 4); else b = 0;
        for (int oh = 0; oh < d->out_h; oh++)
            for (int ow = 0; ow < d->out_w; ow++) {
                uint32_t sum = (uint32_t)b; /* two's-complement wrap, as x86 executes it */
                int wi = 0;
                for (int ic = 0; ic < d->in_c; ic++) {
                    const int8_t *ch = input + (size_t)ic * d->in_h * d->in_w;
                    for (int y = 0; y < d->kh; y++) {
                        int ih = oh * d->sh - d->pt + y;
                        for (int x = 0; x < d->kw; x++, wi++) {
                            int iw = ow * d->sw - d->pl + x;
                            if (ih >= 0 && ih < d->in_h && iw >= 0 && iw < d->in_w)
                                sum += (uint32_t)((int32_t)ch[ih * d->in_w + iw] * (int32_t)wo[wi]);
                        }
                    }
                }
                output[(size_t)oc * d->out_h * d->out_w + (size_t)oh * d->out_w + ow] = mo_requant_conv((int32_t)sum, cs);
            }
    }
}

/* same result, one output row at a time (vectorisable); legal only when no byte
 * written by this layer is read by it */
static void conv_i8_nchw_rows(uint8_t *A, size_t in, size_t w, const int32_t *bias, size_t out,
                              const mo_conv_dims *d, float cs) {
    const int8_t *input = (const int8_t *)(A + in), *weight = (const int8_t *)(A + w);
    int8_t *output = (int8_t *)(A + out);
    int wpo = d->in_c * d->kh * d->kw;
#pragma omp parallel for schedule(dynamic, 1)
    for (int oc = 0; oc < d->out_c; oc++) {
        uint32_t *acc = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(d->out_w > 0 ? d->out_w : 1));
        const int8_t *wo = weight + (size_t)oc * wpo;
        int32_t b;
        if (bias) memcpy(&b, (const uint8_t *)bias + 4 * (size_t)oc, 4); else b = 0;
        for (int oh = 0; oh < d->out_h; oh++) {
            for (int ow = 0; ow < d->out_w; ow++) acc[ow] = (uint32_t)b;
            for (int ic = 0; ic < d->in_c; ic++) {
                const int8_t *ch = input + (size_t)ic * d->in_h * d->in_w;
                for (int y = 0; y < d->kh; y++) {
                    int ih = oh * d->sh - d->pt + y;
                    if (ih < 0 || ih >= d->in_h) continue;
                    const int8_t *row = ch + (size_t)ih * d->in_w;
                    for (int x = 0; x < d->kw; x++) {
                        int32_t wv = wo[(ic * d->kh + y) * d->kw + x];
                        if (wv == 0) continue;
                        /* ow range with 0 <= ow*sw - pl + x < in_w */
                        int off = x - d->pl;
                        int lo = off >= 0 ? 0 : (-off + d->sw - 1) / d->sw;
                        int hi = (d->in_w - 1 - off) >= 0 ? (d->in_w - 1 - off) / d->sw + 1 : 0;
                        if (hi > d->out_w) hi = d->out_w;
                        if (d->sw == 1) {
                            const int8_t *r = row + off;
                            for (int ow = lo; ow < hi; ow++) acc[ow] += (uint32_t)((int32_t)r[ow] * wv);
                        } else {
                            for (int ow = lo; ow < hi; ow++) acc[ow] += (uint32_t)((int32_t)row[ow * d->sw + off] * wv);
                        }
                    }
                }
            }
            int8_t *o = output + (size_t)oc * d->out_h * d->out_w + (size_t)oh * d->out_w;
            for (int ow = 0; ow < d->out_w; ow++) o[ow] = mo_requant_conv((int32_t)acc[ow], cs);
        }
        free(acc);
    }
}

/* src/mars/mxu_conv.c:726-756 (loop order oh, ow, oc; OHWI weights) */
static void conv_i8_nhwc(uint8_t *A, size_t in, size_t w, const int32_t *bias, size_t out,
                         const mo_conv_dims *d, float cs, int parallel) {
    const int8_t *input = (const int8_t *)(A + in), *weight = (const int8_t *)(A + w);
    int8_t *output = (int8_t *)(A + out);
    int wpo = d->kh * d->kw * d->in_c;
    int row_stride = d->in_w * d->in_c;
#pragma omp parallel for schedule(static) if (parallel)
    for (int oh = 0; oh < d->out_h; oh++)
        for (int ow = 0; ow < d->out_w; ow++) {
            int8_t *op = output + ((size_t)oh * d->out_w + ow) * d->out_c;
            for (int oc = 0; oc < d->out_c; oc++) {
                const int8_t *wo = weight + (size_t)oc * wpo;
                int32_t b;
                if (bias) memcpy(&b, (const uint8_t *)bias + 4 * (size_t)oc, 4); else b = 0;
                uint32_t sum = (uint32_t)b;
                for (int y = 0; y < d->kh; y++) {
                    int ih = oh * d->sh - d->pt + y;
                    if (ih < 0 || ih >= d->in_h) continue;
                    for (int x = 0; x < d->kw; x++) {
                        int iw = ow * d->sw - d->pl + x;
                        if (iw < 0 || iw >= d->in_w) continue;
                        const int8_t *ip = input + (size_t)ih * row_stride + (size_t)iw * d->in_c;
                        const int8_t *wp = wo + (y * d->kw + x) * d->in_c;
                        int32_t s = 0;
                        for (int ic = 0; ic < d->in_c; ic++) s += (int32_t)ip[ic] * (int32_t)wp[ic];
                        sum += (uint32_t)s;
                    }
                }
                op[oc] = mo_requant_conv((int32_t)sum, cs);
            }
        }
}

/* src/mars/mxu_conv.c:685-709: strictly sequential fp32 accumulation ic -> kh -> kw */
static void conv_f32_nchw(uint8_t *A, size_t in, size_t w, size_t bias, int has_bias, size_t out,
                          const mo_conv_dims *d, int parallel) {
    const float *input = (const float *)(A + in), *weight = (const float *)(A + w);
    const float *bp = (const float *)(A + bias);
    float *output = (float *)(A + out);
    int wpo = d->in_c * d->kh * d->kw;
#pragma omp parallel for schedule(dynamic, 1) if (parallel)
    for (int oc = 0; oc < d->out_c; oc++) {
        const float *wo = weight + (size_t)oc * wpo;
        for (int oh = 0; oh < d->out_h; oh++)
            for (int ow = 0; ow < d->out_w; ow++) {
                float sum = has_bias ? bp[oc] : 0.0f; /* built with -ffp-contract=off: one rounding per op */
                int wi = 0;
                for (int ic = 0; ic < d->in_c; ic++) {
                    const float *ch = input + (size_t)ic * d->in_h * d->in_w;
                    for (int y = 0; y < d->kh; y++) {
                        int ih = oh * d->sh - d->pt + y;
                        for (int x = 0; x < d->kw; x++, wi++) {
                            int iw = ow * d->sw - d->pl + x;
                            if (ih >= 0 && ih < d->in_h && iw >= 0 && iw < d->in_w)
                                sum += ch[ih * d->in_w + iw] * wo[wi];
                        }
                    }
                }
                output[(size_t)oc * d->out_h * d->out_w + (size_t)oh * d->out_w + ow] = sum;
            }
    }
}

/* src/mars/mars_runtime.c:511-710 */
static int run_conv(mo_model_t *m, const mars_layer_t *L) {
    const mars_conv_params_t *p = &L->params.conv;
    int ii = mo_find(m, L->input_tensor_ids[0]), oi = mo_find(m, L->output_tensor_ids[0]);
    int wi = mo_find(m, p->weight_tensor_id), bi = mo_find(m, p->bias_tensor_id);
    if (ii < 0 || oi < 0 || wi < 0) return MO_ERR_INVALID_TENSOR;
    const mars_tensor_t *it = &m->tensors[ii], *ot = &m->tensors[oi], *wt = &m->tensors[wi];
    int in_nhwc = it->format == MARS_FORMAT_NHWC, out_nhwc = ot->format == MARS_FORMAT_NHWC;
    mo_conv_dims d;
    if (in_nhwc) { d.in_h = it->shape[1]; d.in_w = it->shape[2]; d.in_c = it->shape[3]; }
    else { d.in_c = it->shape[1]; d.in_h = it->shape[2]; d.in_w = it->shape[3]; }
    if (out_nhwc) { d.out_h = ot->shape[1]; d.out_w = ot->shape[2]; d.out_c = ot->shape[3]; }
    else { d.out_c = ot->shape[1]; d.out_h = ot->shape[2]; d.out_w = ot->shape[3]; }
    d.kh = (int)p->kernel_h; d.kw = (int)p->kernel_w; d.sh = (int)p->stride_h; d.sw = (int)p->stride_w;
    d.pt = d.pl = 0;
    if (p->padding == MARS_PAD_SAME) { /* :591-598, explicit pads are ignored */
        d.pt = ((d.out_h - 1) * d.sh + d.kh - d.in_h) / 2;
        d.pl = ((d.out_w - 1) * d.sw + d.kw - d.in_w) / 2;
    }
    int is_float = it->dtype == MARS_DTYPE_FLOAT32;
    size_t in = m->toff[ii], out = m->toff[oi], w = m->toff[wi];
    size_t es = is_float ? 4 : 1;
    size_t npx_o = (d.out_c > 0 && d.out_h > 0 && d.out_w > 0) ? (size_t)d.out_c * d.out_h * d.out_w : 0;
    size_t npx_i = (d.in_c > 0 && d.in_h > 0 && d.in_w > 0) ? (size_t)d.in_c * d.in_h * d.in_w : 0;
    size_t wbytes = (size_t)(d.out_c > 0 ? d.out_c : 0) * (size_t)(d.in_c > 0 ? d.in_c : 0) * d.kh * d.kw * es;
    int hazard = ranges_overlap(out, out + npx_o * es, in, in + npx_i * es) ||
                 ranges_overlap(out, out + npx_o * es, w, w + wbytes) ||
                 (bi >= 0 && ranges_overlap(out, out + npx_o * es, m->toff[bi], m->toff[bi] + 4 * (size_t)(d.out_c > 0 ? d.out_c : 0)));
    if (npx_o) {
        if (is_float) {
            conv_f32_nchw(m->arena, in, w, bi >= 0 ? m->toff[bi] : 0, bi >= 0, out, &d, !hazard);
        } else {
            volatile float prod = it->scale * wt->scale; /* mxu_conv.c:639 */
            float cs = prod / ot->scale;
            const int32_t *bias = bi >= 0 ? (const int32_t *)(m->arena + m->toff[bi]) : NULL;
            if (in_nhwc) conv_i8_nhwc(m->arena, in, w, bias, out, &d, cs, !hazard);
            else if (hazard) conv_i8_nchw_literal(m->arena, in, w, bias, out, &d, cs);
            else conv_i8_nchw_rows(m->arena, in, w, bias, out, &d, cs);
        }
    }
    if (p->activation == MARS_ACT_RELU) { /* :700-707, byte-wise whatever the dtype */
        int total = d.out_h * d.out_w * d.out_c;
        int8_t *o = (int8_t *)(m->arena + out);
        for (int i = 0; i < total; i++) if (o[i] < 0) o[i] = 0;
    }
    return MO_OK;
}

/* RESTATEMENT, PARITY UNPINNED: the reference executes DEPTHWISE_CONV2D as a
 * no-op.  mode 1 follows conv2d_int8_mxu's accumulate/requant convention with one
 * input channel per output channel (groups == Ci == Co, weights [C,1,kh,kw] or
 * OHWI [C,kh,kw,1], same bytes). */
static int run_depthwise(mo_model_t *m, const mars_layer_t *L) {
    if (!m->depthwise_mode) return MO_OK; /* src/mars/mars_runtime.c:1168-1170 */
    const mars_conv_params_t *p = &L->params.conv;
    int ii = mo_find(m, L->input_tensor_ids[0]), oi = mo_find(m, L->output_tensor_ids[0]);
    int wi = mo_find(m, p->weight_tensor_id), bi = mo_find(m, p->bias_tensor_id);
    if (ii < 0 || oi < 0 || wi < 0) return MO_ERR_INVALID_TENSOR;
    const mars_tensor_t *it = &m->tensors[ii], *ot = &m->tensors[oi], *wt = &m->tensors[wi];
    if (it->dtype == MARS_DTYPE_FLOAT32) return MO_ERR_INVALID_LAYER;
    int nhwc = it->format == MARS_FORMAT_NHWC;
    int C, ih_, iw_, oh_, ow_;
    if (nhwc) { ih_ = it->shape[1]; iw_ = it->shape[2]; C = it->shape[3]; oh_ = ot->shape[1]; ow_ = ot->shape[2]; }
    else { C = it->shape[1]; ih_ = it->shape[2]; iw_ = it->shape[3]; oh_ = ot->shape[2]; ow_ = ot->shape[3]; }
    int kh = (int)p->kernel_h, kw = (int)p->kernel_w, sh = (int)p->stride_h, sw = (int)p->stride_w, pt = 0, pl = 0;
    if (p->padding == MARS_PAD_SAME) { pt = ((oh_ - 1) * sh + kh - ih_) / 2; pl = ((ow_ - 1) * sw + kw - iw_) / 2; }
    else if (p->padding == MARS_PAD_EXPLICIT) { pt = (int)p->pad_top; pl = (int)p->pad_left; }
    volatile float prod = it->scale * wt->scale;
    float cs = prod / ot->scale;
    const int8_t *in = (const int8_t *)(m->arena + m->toff[ii]), *w = (const int8_t *)(m->arena + m->toff[wi]);
    const uint8_t *bias = bi >= 0 ? m->arena + m->toff[bi] : NULL;
    int8_t *tmp = (int8_t *)malloc((size_t)C * oh_ * ow_ + 1);
    for (int c = 0; c < C; c++) {
        int32_t b = 0;
        if (bias) memcpy(&b, bias + 4 * (size_t)c, 4);
        for (int oh = 0; oh < oh_; oh++)
            for (int ow = 0; ow < ow_; ow++) {
                uint32_t sum = (uint32_t)b;
                for (int y = 0; y < kh; y++)
                    for (int x = 0; x < kw; x++) {
                        int ih = oh * sh - pt + y, iw = ow * sw - pl + x;
                        if (ih < 0 || ih >= ih_ || iw < 0 || iw >= iw_) continue;
                        int8_t v = nhwc ? in[((size_t)ih * iw_ + iw) * C + c] : in[((size_t)c * ih_ + ih) * iw_ + iw];
                        sum += (uint32_t)((int32_t)v * (int32_t)w[((size_t)c * kh + y) * kw + x]);
                    }
                size_t o = nhwc ? ((size_t)oh * ow_ + ow) * C + c : ((size_t)c * oh_ + oh) * ow_ + ow;
                tmp[o] = mo_requant_conv((int32_t)sum, cs);
            }
    }
    memcpy(m->arena + m->toff[oi], tmp, (size_t)C * oh_ * ow_);
    free(tmp);
    return MO_OK;
}

/* ------------------------------------------------------------------------- */
/* element-wise, pooling, concat, upsample, batchnorm                         */
/* ------------------------------------------------------------------------- */

/* src/mars/mars_runtime.c:724-771 */
static int run_sigmoid(mo_model_t *m, const mars_layer_t *L) {
    int ii = mo_find(m, L->input_tensor_ids[0]), oi = mo_find(m, L->output_tensor_ids[0]);
    if (ii < 0 || oi < 0) return MO_ERR_INVALID_TENSOR;
    const mars_tensor_t *it = &m->tensors[ii], *ot = &m->tensors[oi];
    size_t n = mo_numel(it);
    if (it->dtype == MARS_DTYPE_FLOAT32) {
        const float *in = (const float *)(m->arena + m->toff[ii]);
        float *out = (float *)(m->arena + m->toff[oi]);
        for (size_t i = 0; i < n; i++) out[i] = 1.0f / (1.0f + expf(-in[i]));
        return MO_OK;
    }
    const int8_t *in = (const int8_t *)(m->arena + m->toff[ii]);
    int8_t *out = (int8_t *)(m->arena + m->toff[oi]);
    float is = it->scale, os = ot->scale > 0 ? ot->scale : 1.0f;
    int8_t lut[256]; /* only 256 distinct inputs: tabulate with the host libm */
    for (int v = -128; v < 128; v++) {
        volatile float x = (float)v * is;
        volatile float e = expf(-x);
        volatile float den = 1.0f + e;
        volatile float y = 1.0f / den;
        lut[v + 128] = mo_requant_div(y, os);
    }
    for (size_t i = 0; i < n; i++) out[i] = lut[in[i] + 128];
    return MO_OK;
}

/* src/mars/mars_runtime.c:774-838 (mul), :841-905 (add); numel from input A only */
static int run_binary(mo_model_t *m, const mars_layer_t *L, int is_add) {
    int ai = mo_find(m, L->input_tensor_ids[0]), bi = mo_find(m, L->input_tensor_ids[1]);
    int oi = mo_find(m, L->output_tensor_ids[0]);
    if (ai < 0 || bi < 0 || oi < 0) return MO_ERR_INVALID_TENSOR;
    const mars_tensor_t *at = &m->tensors[ai], *bt = &m->tensors[bi], *ot = &m->tensors[oi];
    size_t n = mo_numel(at);
    if (at->dtype == MARS_DTYPE_FLOAT32) {
        const float *a = (const float *)(m->arena + m->toff[ai]), *b = (const float *)(m->arena + m->toff[bi]);
        float *o = (float *)(m->arena + m->toff[oi]);
        for (size_t i = 0; i < n; i++) { volatile float r = is_add ? a[i] + b[i] : a[i] * b[i]; o[i] = r; }
        return MO_OK;
    }
    const int8_t *a = (const int8_t *)(m->arena + m->toff[ai]), *b = (const int8_t *)(m->arena + m->toff[bi]);
    int8_t *o = (int8_t *)(m->arena + m->toff[oi]);
    float sa = at->scale, sb = bt->scale, so = ot->scale > 0 ? ot->scale : 1.0f;
    volatile float inv = 1.0f / so;
    for (size_t i = 0; i < n; i++) {
        volatile float va = (float)a[i] * sa;
        volatile float vb = (float)b[i] * sb;
        volatile float y = is_add ? va + vb : va * vb;
        o[i] = mo_requant_mul_inv(y, inv);
    }
    return MO_OK;
}

/* src/mars/mars_runtime.c:1047-1089: RELU, RELU6 (no upper clamp), LEAKY (alpha fixed 0.01) */
static int run_relu(mo_model_t *m, const mars_layer_t *L) {
    int ii = mo_find(m, L->input_tensor_ids[0]), oi = mo_find(m, L->output_tensor_ids[0]);
    if (ii < 0 || oi < 0) return MO_ERR_INVALID_TENSOR;
    const mars_tensor_t *it = &m->tensors[ii];
    size_t n = mo_numel(it);
    int leaky = L->type == MARS_LAYER_LEAKY_RELU;
    float alpha = leaky ? 0.01f : 0.0f;
    if (it->dtype == MARS_DTYPE_FLOAT32) {
        const float *in = (const float *)(m->arena + m->toff[ii]);
        float *out = (float *)(m->arena + m->toff[oi]);
        for (size_t i = 0; i < n; i++) { volatile float r = in[i] > 0.0f ? in[i] : in[i] * alpha; out[i] = r; }
        return MO_OK;
    }
    const int8_t *in = (const int8_t *)(m->arena + m->toff[ii]);
    int8_t *out = (int8_t *)(m->arena + m->toff[oi]);
    for (size_t i = 0; i < n; i++) {
        if (in[i] > 0) out[i] = in[i];
        else if (leaky) {
            volatile float t = (float)in[i] * alpha;
            int32_t v = mo_f2i_x86(t);
            out[i] = (int8_t)(v < -128 ? -128 : v);
        } else out[i] = 0;
    }
    return MO_OK;
}

/* src/mars/mars_runtime.c:1092-1158 */
static int run_batchnorm(mo_model_t *m, const mars_layer_t *L) {
    int ii = mo_find(m, L->input_tensor_ids[0]), si = mo_find(m, L->input_tensor_ids[1]);
    int bi = mo_find(m, L->input_tensor_ids[2]), oi = mo_find(m, L->output_tensor_ids[0]);
    if (ii < 0 || oi < 0) return MO_ERR_INVALID_TENSOR;
    const mars_tensor_t *it = &m->tensors[ii], *ot = &m->tensors[oi];
    int n = it->shape[0] > 0 ? it->shape[0] : 1, c = it->shape[1] > 0 ? it->shape[1] : 1;
    int h = it->shape[2] > 0 ? it->shape[2] : 1, w = it->shape[3] > 0 ? it->shape[3] : 1;
    const float *s = si >= 0 ? (const float *)(m->arena + m->toff[si]) : NULL;
    const float *b = bi >= 0 ? (const float *)(m->arena + m->toff[bi]) : NULL;
    size_t plane = (size_t)h * w;
    if (it->dtype == MARS_DTYPE_FLOAT32) {
        const float *in = (const float *)(m->arena + m->toff[ii]);
        float *out = (float *)(m->arena + m->toff[oi]);
        for (int ni = 0; ni < n; ni++)
            for (int ci = 0; ci < c; ci++) {
                float sc = s ? s[ci] : 1.0f, bb = b ? b[ci] : 0.0f;
                size_t base = ((size_t)ni * c + ci) * plane;
                for (size_t k = 0; k < plane; k++) { volatile float t = in[base + k] * sc; volatile float y = t + bb; out[base + k] = y; }
            }
        return MO_OK;
    }
    const int8_t *in = (const int8_t *)(m->arena + m->toff[ii]);
    int8_t *out = (int8_t *)(m->arena + m->toff[oi]);
    float is = it->scale > 0 ? it->scale : 1.0f, os = ot->scale > 0 ? ot->scale : 1.0f;
    for (int ni = 0; ni < n; ni++)
        for (int ci = 0; ci < c; ci++) {
            float sc = s ? s[ci] : 1.0f, bb = b ? b[ci] : 0.0f;
            size_t base = ((size_t)ni * c + ci) * plane;
            for (size_t k = 0; k < plane; k++) {
                volatile float x = (float)in[base + k] * is;
                volatile float t = x * sc;
                volatile float y = t + bb;
                out[base + k] = mo_requant_div(y, os);
            }
        }
    return MO_OK;
}

/* src/mars/mars_runtime.c:908-960: NHWC indexing of shape[1..3], pads ignored */
static int run_maxpool(mo_model_t *m, const mars_layer_t *L) {
    const mars_pool_params_t *p = &L->params.pool;
    int ii = mo_find(m, L->input_tensor_ids[0]), oi = mo_find(m, L->output_tensor_ids[0]);
    if (ii < 0 || oi < 0) return MO_ERR_INVALID_TENSOR;
    const mars_tensor_t *it = &m->tensors[ii], *ot = &m->tensors[oi];
    int ih_ = it->shape[1], iw_ = it->shape[2], C = it->shape[3], oh_ = ot->shape[1], ow_ = ot->shape[2];
    int kh = (int)p->kernel_h, kw = (int)p->kernel_w, sh = (int)p->stride_h, sw = (int)p->stride_w;
    const int8_t *in = (const int8_t *)(m->arena + m->toff[ii]);
    int8_t *out = (int8_t *)(m->arena + m->toff[oi]);
    for (int c = 0; c < C; c++)
        for (int oh = 0; oh < oh_; oh++)
            for (int ow = 0; ow < ow_; ow++) {
                int8_t mx = -128;
                for (int y = 0; y < kh; y++)
                    for (int x = 0; x < kw; x++) {
                        int ih = oh * sh + y, iw = ow * sw + x;
                        if (ih < ih_ && iw < iw_) {
                            int8_t v = in[ih * iw_ * C + iw * C + c];
                            if (v > mx) mx = v;
                        }
                    }
                out[oh * ow_ * C + ow * C + c] = mx;
            }
    return MO_OK;
}

/* src/mars/mars_runtime.c:963-1000: byte loop in ascending order (overlap-visible) */
static int run_concat(mo_model_t *m, const mars_layer_t *L) {
    int oi = mo_find(m, L->output_tensor_ids[0]);
    if (oi < 0) return MO_ERR_INVALID_TENSOR;
    const mars_tensor_t *ot = &m->tensors[oi];
    int oh_ = ot->shape[1], ow_ = ot->shape[2], oc_ = ot->shape[3];
    int8_t *out = (int8_t *)(m->arena + m->toff[oi]);
    int coff = 0;
    uint32_t nin = L->num_inputs > 4 ? 4 : L->num_inputs;
    for (uint32_t n = 0; n < nin; n++) {
        int ii = mo_find(m, L->input_tensor_ids[n]);
        if (ii < 0) continue;
        int ic = m->tensors[ii].shape[3];
        const int8_t *in = (const int8_t *)(m->arena + m->toff[ii]);
        for (int h = 0; h < oh_; h++)
            for (int w = 0; w < ow_; w++)
                for (int c = 0; c < ic; c++)
                    out[h * ow_ * oc_ + w * oc_ + (coff + c)] = in[h * ow_ * ic + w * ic + c];
        coff += ic;
    }
    return MO_OK;
}

/* src/mars/mars_runtime.c:1003-1044 */
static int run_upsample(mo_model_t *m, const mars_layer_t *L) {
    const mars_upsample_params_t *p = &L->params.upsample;
    int ii = mo_find(m, L->input_tensor_ids[0]), oi = mo_find(m, L->output_tensor_ids[0]);
    if (ii < 0 || oi < 0) return MO_ERR_INVALID_TENSOR;
    const mars_tensor_t *it = &m->tensors[ii], *ot = &m->tensors[oi];
    int ih_ = it->shape[1], iw_ = it->shape[2], C = it->shape[3], oh_ = ot->shape[1], ow_ = ot->shape[2];
    if (oh_ <= 0 || ow_ <= 0 || C <= 0) return MO_OK;
    if ((p->scale_h == 0 && ih_ == 0) || (p->scale_w == 0 && iw_ == 0)) return MO_ERR_INVALID_LAYER; /* reference would divide by zero */
    int sh = p->scale_h > 0 ? (int)p->scale_h : oh_ / ih_, sw = p->scale_w > 0 ? (int)p->scale_w : ow_ / iw_;
    if (sh == 0 || sw == 0) return MO_ERR_INVALID_LAYER;
    const int8_t *in = (const int8_t *)(m->arena + m->toff[ii]);
    int8_t *out = (int8_t *)(m->arena + m->toff[oi]);
    for (int oh = 0; oh < oh_; oh++) {
        int ih = oh / sh; if (ih >= ih_) ih = ih_ - 1;
        for (int ow = 0; ow < ow_; ow++) {
            int iw = ow / sw; if (iw >= iw_) iw = iw_ - 1;
            for (int c = 0; c < C; c++) out[oh * ow_ * C + ow * C + c] = in[ih * iw_ * C + iw * C + c];
        }
    }
    return MO_OK;
}

/* dispatcher: src/mars/mars_runtime.c:1161-1224 */
int mo_run_layer(mo_model_t *m, uint32_t i) {
    if (!m || i >= m->header.num_layers) return MO_ERR_INVALID_LAYER;
    const mars_layer_t *L = &m->layers[i];
    switch ((int)L->type) {
        case MARS_LAYER_CONV2D: return run_conv(m, L);
        case MARS_LAYER_DEPTHWISE_CONV2D: return run_depthwise(m, L);
        case MARS_LAYER_MAXPOOL: return run_maxpool(m, L);
        case MARS_LAYER_RELU: case MARS_LAYER_RELU6: case MARS_LAYER_LEAKY_RELU: return run_relu(m, L);
        case MARS_LAYER_SIGMOID: return run_sigmoid(m, L);
        case MARS_LAYER_CONCAT: return run_concat(m, L);
        case MARS_LAYER_ADD: return run_binary(m, L, 1);
        case MARS_LAYER_MUL: return run_binary(m, L, 0);
        case MARS_LAYER_UPSAMPLE: return run_upsample(m, L);
        case MARS_LAYER_BATCHNORM: return run_batchnorm(m, L);
        case MARS_LAYER_AVGPOOL: case MARS_LAYER_SILU: case MARS_LAYER_RESHAPE:
        case MARS_LAYER_TRANSPOSE: case MARS_LAYER_SOFTMAX: return MO_OK; /* no-ops in the reference */
        default: return MO_ERR_INVALID_LAYER; /* GLOBAL_AVGPOOL(4), FC(16), unknown */
    }
}

/* src/mars/mars_runtime.c:439-459 */
int mo_run(mo_model_t *m) {
    if (!m) return MO_ERR_INVALID_FILE;
    for (uint32_t i = 0; i < m->header.num_layers; i++) {
        int e = mo_run_layer(m, i);
        if (e != MO_OK) return e;
    }
    return MO_OK;
}

/* ------------------------------------------------------------------------- */
/* post-process                                                                */
/* ------------------------------------------------------------------------- */

/* src/mars/mars_yolo_test.c:80-104 */
int mo_parse_output(const int8_t *data, int npred, float scale, mo_det_t *dets, int maxd) {
    int cnt = 0;
    for (int i = 0; i < npred && cnt < maxd; i++) {
        const int8_t *p = data + (size_t)i * 85;
        volatile float a = -(float)p[4] * scale;
        volatile float e = expf(a);
        volatile float den = 1.0f + e;
        volatile float obj = 1.0f / den;
        if (obj < 0.25f) continue;
        int best_c = 0;
        float best_s = -1e9f;
        for (int c = 0; c < 80; c++) {
            volatile float s = (float)p[5 + c] * scale;
            if (s > best_s) { best_s = s; best_c = c; }
        }
        volatile float e2 = expf(-best_s);
        volatile float den2 = 1.0f + e2;
        volatile float conf = obj / den2;
        if (conf < 0.25f) continue;
        dets[cnt].x = (float)p[0] * scale; dets[cnt].y = (float)p[1] * scale;
        dets[cnt].w = (float)p[2] * scale; dets[cnt].h = (float)p[3] * scale;
        dets[cnt].conf = conf; dets[cnt].cls = best_c;
        cnt++;
    }
    return cnt;
}

/* One pass of the reference's exchange sort (src/mars/mars_yolo_test.c:108-110) in
 * closed form: walking j upward, the element at i is replaced at every strict
 * prefix-maximum "record"; each record position receives the previous record's
 * element. */
static void exchange_pass(mo_det_t *d, int i, int n) {
    mo_det_t carry = d[i];
    for (int j = i + 1; j < n; j++)
        if (d[j].conf > carry.conf) { mo_det_t t = d[j]; d[j] = carry; carry = t; }
    d[i] = carry;
}

static float mo_iou_center(const mo_det_t *a, const mo_det_t *b) {
    volatile float ax1 = a->x - a->w / 2, bx1 = b->x - b->w / 2, ay1 = a->y - a->h / 2, by1 = b->y - b->h / 2;
    volatile float ax2 = a->x + a->w / 2, bx2 = b->x + b->w / 2, ay2 = a->y + a->h / 2, by2 = b->y + b->h / 2;
    float x1 = fmaxf(ax1, bx1), y1 = fmaxf(ay1, by1), x2 = fminf(ax2, bx2), y2 = fminf(ay2, by2);
    volatile float dx = x2 - x1, dy = y2 - y1;
    volatile float inter = fmaxf(0, dx) * fmaxf(0, dy);
    volatile float aa = a->w * a->h, ab = b->w * b->h;
    volatile float u = aa + ab;
    volatile float u2 = u - inter;
    volatile float u3 = u2 + 1e-6f;
    volatile float r = inter / u3;
    return r;
}

/* src/mars/mars_yolo_test.c:107-130 */
int mo_nms(mo_det_t *d, int n, float thresh) {
    for (int i = 0; i < n - 1; i++) exchange_pass(d, i, n);
    uint8_t *sup = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
    for (int i = 0; i < n; i++) {
        if (sup[i]) continue;
        for (int j = i + 1; j < n; j++) {
            if (sup[j] || d[i].cls != d[j].cls) continue;
            if (mo_iou_center(&d[i], &d[j]) > thresh) sup[j] = 1;
        }
    }
    int out = 0;
    for (int i = 0; i < n; i++) if (!sup[i]) d[out++] = d[i];
    free(sup);
    return out;
}

/* examples/yolo_detect.cpp:138-149 */
float mo_iou_corner(const mo_box_t *a, const mo_box_t *b) {
    float x0 = fmaxf(a->x0, b->x0), y0 = fmaxf(a->y0, b->y0), x1 = fminf(a->x1, b->x1), y1 = fminf(a->y1, b->y1);
    volatile float dx = x1 - x0, dy = y1 - y0;
    volatile float inter = fmaxf(0, dx) * fmaxf(0, dy);
    volatile float aw = a->x1 - a->x0, ah = a->y1 - a->y0, bw = b->x1 - b->x0, bh = b->y1 - b->y0;
    volatile float aa = aw * ah, ab = bw * bh;
    volatile float u = aa + ab;
    volatile float u2 = u - inter;
    volatile float u3 = u2 + 1e-6f;
    volatile float r = inter / u3;
    return r;
}

/* examples/yolo_detect.cpp:152-173.  std::sort's order among equal confidences is
 * implementation-defined, so this restatement (stable insertion order among ties)
 * is pinned against the reference on tie-free inputs only. */
int mo_nms_corner(mo_box_t *d, int n, float thresh) {
    for (int i = 1; i < n; i++) { /* stable insertion sort, descending */
        mo_box_t t = d[i];
        int j = i - 1;
        while (j >= 0 && d[j].confidence < t.confidence) { d[j + 1] = d[j]; j--; }
        d[j + 1] = t;
    }
    uint8_t *sup = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
    int out = 0;
    mo_box_t *res = (mo_box_t *)malloc(sizeof(mo_box_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        if (sup[i]) continue;
        res[out++] = d[i];
        for (int j = i + 1; j < n; j++) {
            if (sup[j]) continue;
            if (d[i].class_id == d[j].class_id && mo_iou_corner(&d[i], &d[j]) > thresh) sup[j] = 1;
        }
    }
    memcpy(d, res, sizeof(mo_box_t) * (size_t)out);
    free(res); free(sup);
    return out;
}

/* examples/yolo_detect.cpp:208-227 */
void mo_scale_detections(mo_box_t *d, int n, int orig_w, int orig_h, int net_w, int net_h) {
    float scale = fminf((float)net_w / orig_w, (float)net_h / orig_h);
    volatile float sw = orig_w * scale, sh = orig_h * scale;
    volatile float px = (net_w - sw) / 2, py = (net_h - sh) / 2;
    for (int i = 0; i < n; i++) {
        volatile float a = (d[i].x0 - px) / scale, b = (d[i].y0 - py) / scale;
        volatile float c = (d[i].x1 - px) / scale, e = (d[i].y1 - py) / scale;
        d[i].x0 = fmaxf(0, fminf(a, (float)orig_w - 1)); d[i].y0 = fmaxf(0, fminf(b, (float)orig_h - 1));
        d[i].x1 = fmaxf(0, fminf(c, (float)orig_w - 1)); d[i].y1 = fmaxf(0, fminf(e, (float)orig_h - 1));
    }
}

/* RESTATEMENT, PARITY UNPINNED IN C (stub at examples/yolo_detect.cpp:184-205):
 * anchor-grid decode following mgk-decompiler/test_yolo_inference.py:136-202 in
 * fp32, anchors/strides from examples/yolo_detect.cpp:176-181.  head: int8
 * [3, gh, gw, 85] dequantised with `scale`; appends to dets (corner boxes) up to
 * maxd, order = anchor, y, x. */
static const float MO_ANCHORS[3][6] = {{10, 13, 16, 30, 33, 23}, {30, 61, 62, 45, 59, 119}, {116, 90, 156, 198, 373, 326}};
static const int MO_STRIDES[3] = {8, 16, 32};

static inline float mo_sigmoidf(float x) { volatile float e = expf(-x); volatile float d = 1.0f + e; volatile float r = 1.0f / d; return r; }

int mo_decode_anchor_grid(const int8_t *head, int gh, int gw, float scale, int level, float conf_thresh,
                          mo_box_t *dets, int cnt, int maxd) {
    float stride = (float)MO_STRIDES[level];
    for (int a = 0; a < 3; a++)
        for (int y = 0; y < gh; y++)
            for (int x = 0; x < gw; x++) {
                if (cnt >= maxd) return cnt;
                const int8_t *p = head + (((size_t)a * gh + y) * gw + x) * 85;
                float obj = mo_sigmoidf((float)p[4] * scale);
                if (obj < conf_thresh) continue;
                int best = 0; int8_t bv = p[5];
                for (int c = 1; c < 80; c++) if (p[5 + c] > bv) { bv = p[5 + c]; best = c; } /* argmax of a monotone map (scale > 0) */
                volatile float conf = obj * mo_sigmoidf((float)bv * scale);
                if (conf < conf_thresh) continue;
                volatile float sx = mo_sigmoidf((float)p[0] * scale) * 2.0f, sy = mo_sigmoidf((float)p[1] * scale) * 2.0f;
                volatile float cx0 = sx - 0.5f, cy0 = sy - 0.5f;
                volatile float cx1 = cx0 + (float)x, cy1 = cy0 + (float)y;
                volatile float cx = cx1 * stride, cy = cy1 * stride;
                volatile float tw = mo_sigmoidf((float)p[2] * scale) * 2.0f, th = mo_sigmoidf((float)p[3] * scale) * 2.0f;
                volatile float tw2 = tw * tw, th2 = th * th;
                volatile float w = tw2 * MO_ANCHORS[level][2 * a], h = th2 * MO_ANCHORS[level][2 * a + 1];
                volatile float hw = w / 2, hh = h / 2;
                dets[cnt].x0 = cx - hw; dets[cnt].y0 = cy - hh; dets[cnt].x1 = cx + hw; dets[cnt].y1 = cy + hh;
                dets[cnt].confidence = conf; dets[cnt].class_id = best;
                cnt++;
            }
    return cnt;
}

/* ------------------------------------------------------------------------- */
/* mars_math.h / mxu_ops.h helpers                                             */
/* ------------------------------------------------------------------------- */

/* src/mars/mars_math.c:14-55, src/mars/mxu_ops.c:144-164 */
void mo_vec_add_f32(float *dst, const float *a, const float *b, size_t n) { for (size_t i = 0; i < n; i++) { volatile float r = a[i] + b[i]; dst[i] = r; } }
void mo_vec_sub_f32(float *dst, const float *a, const float *b, size_t n) { for (size_t i = 0; i < n; i++) { volatile float r = a[i] - b[i]; dst[i] = r; } }
void mo_vec_mul_f32(float *dst, const float *a, const float *b, size_t n) { for (size_t i = 0; i < n; i++) { volatile float r = a[i] * b[i]; dst[i] = r; } }
void mo_vec_relu_f32(float *dst, const float *a, size_t n) { for (size_t i = 0; i < n; i++) dst[i] = a[i] > 0.0f ? a[i] : 0.0f; }
float mo_vec_dot_f32(const float *a, const float *b, size_t n) {
    volatile float acc = 0.0f;
    for (size_t i = 0; i < n; i++) { volatile float p = a[i] * b[i]; acc = acc + p; }
    return acc;
}
void mo_matmul_f32(float *Cm, const float *A, const float *B, size_t M, size_t K, size_t N) {
    for (size_t m = 0; m < M; m++)
        for (size_t n = 0; n < N; n++) {
            volatile float acc = 0.0f;
            for (size_t k = 0; k < K; k++) { volatile float p = A[m * K + k] * B[k * N + n]; acc = acc + p; }
            Cm[m * N + n] = acc;
        }
}
