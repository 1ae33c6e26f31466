/*
 * mars_internal.h -- private types of libmars_b200.so.
 *
 * Data layout in HBM:
 *   d_weights : the model's weight blob, one copy per GPU, shared by every image;
 *   d_slots   : `capacity` image slots, each = the reference's work buffers
 *               [buf0 | buf1 | (buf2)] exactly as the planner of reference
 *               src/mars/mars_runtime.c:248-337 lays them out behind the weights.
 * Every tensor is addressed by its byte offset in the reference's arena
 * ("arena offset"): offsets below weights_size resolve into d_weights, the rest
 * into the image's slot.  A layer compiles into one or more `Op`s over arena offsets.
 */
#ifndef MARS_INTERNAL_H
#define MARS_INTERNAL_H

#include <stddef.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/mars_b200.h"

namespace marsb200 {

enum OpKind : int {
    OP_NOP = 0,
    OP_CONV_I8_NCHW,   /* reference src/mars/mxu_conv.c:630-670 */
    OP_CONV_I8_NHWC,   /* reference src/mars/mxu_conv.c:713-757 */
    OP_CONV_F32_NCHW,  /* reference src/mars/mxu_conv.c:673-710 */
    OP_DW_I8,          /* restated depthwise (parity unpinned) */
    OP_BYTE_RELU,      /* reference src/mars/mars_runtime.c:700-707 */
    OP_SIGMOID_I8,     /* reference src/mars/mars_runtime.c:752-768 (as a 256-entry table) */
    OP_SIGMOID_F32,    /* :742-749 */
    OP_MUL_I8, OP_ADD_I8,   /* :818-835, :885-902 */
    OP_MUL_F32, OP_ADD_F32, /* :807-815, :874-882 */
    OP_RELU_I8,        /* :1072-1085 (relu / relu6 / leaky as a 256-entry table) */
    OP_RELU_F32,       /* :1066-1071 */
    OP_BN_I8, OP_BN_F32, /* :1092-1158 */
    OP_MAXPOOL,        /* :908-960 */
    OP_CONCAT,         /* :963-1000, one op per input, in order */
    OP_CONCAT_PERIODIC,/* in-place concat input with constant shift */
    OP_UPSAMPLE,       /* :1003-1044 */
    OP_LUT_I8,         /* fused chain of int8 unary layers as one table */
    OP_KIND_COUNT
};

enum ExecMode : int {
    EXEC_PARALLEL = 0,  /* one thread (or tile) per output element, any order */
    EXEC_OC_PASSES,     /* NCHW conv whose output aliases its input: one launch per output channel */
    EXEC_OC_PASSES_SCRATCH, /* same, each pass staged through a scratch plane */
    EXEC_PIXEL_SERIAL,  /* NHWC conv in place: thread per pixel, output channels in order */
    EXEC_SERIAL         /* literal reference loop order on one thread per image */
};

enum ConvImpl : int {
    CONV_DIRECT = 0,    /* exact direct kernel (any shape) */
    CONV_TC_NCHW,       /* tcgen05 implicit GEMM (conv_tc.cu): NCHW / OIHW layers and, despite the name, NHWC / OHWI ones */
    CONV_TC_F32         /* tcgen05 kind::tf32 implicit GEMM for float32 layers (conv_tf32.cu) */
};

/* epilogue applied to an int8 value produced by an op (fused following layers) */
struct Epilogue {
    int lut_off;        /* >=0: offset of a 256-byte table in the const pool applied to the int8 result */
    int64_t extra_out[2]; /* additional arena offsets the *pre-table* stages are also written to, -1 = none */
    int extra_lut[2];     /* table applied for each extra output (-1 = identity) */
};

struct Op {
    int kind = OP_NOP;
    int mode = EXEC_PARALLEL;
    int impl = CONV_DIRECT;
    int layer = -1;        /* originating layer index */
    int fused_layers = 0;  /* how many following layers were folded into this op */
    bool xlat = false;     /* some operand straddles the weights/slot boundary: per-access translation */
    int64_t in0 = -1, in1 = -1, in2 = -1, out = -1, w = -1, bias = -1; /* arena offsets */
    int ic = 0, ih = 0, iw = 0, oc = 0, oh = 0, ow = 0;
    int kh = 1, kw = 1, sh = 1, sw = 1, pt = 0, pl = 0;
    int coff = 0;          /* concat channel offset / periodic shift */
    int in_ctot = 0;       /* concat: channels of this input */
    float f0 = 0, f1 = 0, f2 = 0; /* scales: conv combined scale; eltwise sa, sb, inv/so */
    uint64_t n = 0;        /* element count for flat ops */
    int lut = -1;          /* const-pool offset of a 256-byte table (sigmoid/relu/fused) */
    int post_relu = 0;     /* conv: byte-ReLU folded into the epilogue */
    int post_lut = -1;     /* conv: fused following unary chain */
    /* conv with the following SIGMOID and MUL layers folded into its epilogue: y = conv output
     * byte, S = lut_s[y] stored at out_s, Z = lut_z[y] stored at out_z; store_y=false when a
     * later stage of the same chain overwrites Y's bytes anyway */
    bool store_y = true, store_z = true;
    int64_t out_s = -1, out_z = -1;
    int lut_s = -1, lut_z = -1;
    /* producer -> consumer link: a kxk tensor-core conv reads a channel-innermost, zero-padded copy of
     * its input; when that input is exactly one output stream of an earlier tensor-core conv, the producer's
     * epilogue writes the copy itself (nhwc_consumer / nhwc_stream: 0 = Z, 1 = S, 2 = Y) and the consumer skips its
     * own layout pre-pass (copy_from = producer op, copy_off = byte offset of its region in the per-image link area) */
    int copy_from = -1, nhwc_consumer = -1, nhwc_stream = -1;
    int64_t copy_off = 0;
    /* 1x1 tensor-core conv whose fused outputs overwrite its own input buffer while other CTAs (a second N tile) still read it:
     * the kernel reads a private device copy of the input made just before the launch (round-robin work buffers) */
    bool private_in = false;
    /* concat forwarding (opt level 3): the op's stream fwd_stream (0 = Z, 1 = S, 2 = Y) is ALSO written at arena offset fwd_out --
     * its place in the output of the concat that would copy it there later (program.cpp forward_concat_inputs) */
    int64_t fwd_out = -1;
    int fwd_stream = -1;
    /* write/read extents for hazard analysis and bounds checks */
    int64_t wlo = 0, whi = 0;
    std::string note;
};

struct Program {
    std::vector<Op> ops;
    std::vector<uint8_t> const_pool; /* 256-byte tables, uploaded once */
    size_t scratch_bytes = 0;        /* per-image scratch for EXEC_OC_PASSES_SCRATCH */
    size_t linked_bytes = 0;         /* per-image bytes of producer-written conv input copies */
    size_t max_extent = 0;           /* highest arena offset (exclusive) any op reads or writes: an image slot must reach that far */
};

/* device-side view of the arena; passed by value to kernels */
struct ArenaView {
    const uint8_t *wbase; /* d_weights */
    uint8_t *sbase;       /* slot 0 of the launch */
    uint64_t W;           /* weights_size: arena offsets >= W live in the slot */
    uint64_t slot_stride;
    const uint8_t *cpool; /* const pool */
    uint8_t *scratch;     /* per-image scratch base */
    uint64_t scratch_stride;
};

struct Model; /* defined in runtime.cu */

#ifdef __CUDACC__
/* ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------------
 * The ~110 kernels of a step run back to back on one stream; each is launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization, so the next kernel's CTAs may be scheduled (and run their prologue:
 * barrier init, TMEM allocation, table fill) while the tail of the current one drains, instead of after it.  Every such kernel
 * releases its dependents at its very start and executes griddepcontrol.wait before its first access to memory another
 * kernel of the step writes or reads -- the wait returns once the preceding grid has completed and its writes are visible.
 * Opt-in (MARS_PDL=1): measured slower than plain stream order on this workload, see pdl_enabled() in runtime.cu; without the
 * attribute the wait is a no-op. */
__device__ __forceinline__ void pdl_release_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_prior_grid() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_begin() { pdl_release_dependents(); pdl_wait_prior_grid(); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

/* ---- host-side helpers shared between translation units ---- */
size_t tensor_byte_size(const mars_tensor_t *t);
int find_tensor(const mars_header_t &h, const mars_runtime_tensor_t *tensors, uint32_t id);
void set_last_error(const char *fmt, ...);
/* once per CUDA device: true the first time it is called with this mask on the current device (function attributes and
 * SM counts are per device; the API supports several devices in one process) */
bool first_time_on_device(unsigned long long *mask);

/* compile the layer table into ops (program.cpp) */
mars_error_t compile_program(const mars_header_t &h, const mars_runtime_tensor_t *tensors,
                             const mars_runtime_layer_t *layers, const std::vector<size_t> &toff,
                             size_t weights_size, size_t arena_size, int opt_level, int depthwise_mode,
                             Program *out, int f32_mode = 0);
std::string describe_program(const Program &p);

} // namespace marsb200

#endif
