/*
 * mars.h -- on-disk layout of a `.mars` model file (B200 build).
 *
 * ABI contract: byte-for-byte the packed little-endian layout the reference
 * declares in /root/reference include/mars.h:103-221 (header 76 B, tensor
 * descriptor 124 B, layer descriptor 112 B; the "64/64/128 bytes" remarks in the
 * reference header are stale, the real sizes are pinned by
 * mars-compiler/src/mars_format.rs:14-19 and tools/mars_gen_test.py:8-11).
 * Type, field and enumerator names are kept so that callers written against the
 * reference header compile unchanged.  The sizes are enforced below with
 * static assertions.
 */
#ifndef MARS_H
#define MARS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MARS_MAGIC 0x5352414Du /* bytes 'M','A','R','S' */
#define MARS_VERSION_MAJOR 1
#define MARS_VERSION_MINOR 0

#define MARS_MAX_DIMS 6
#define MARS_MAX_NAME_LEN 64
#define MARS_MAX_LAYERS 256
#define MARS_MAX_TENSORS 512

/* element types (reference include/mars.h:35-42) */
typedef enum {
    MARS_DTYPE_FLOAT32 = 0,
    MARS_DTYPE_INT32 = 1,
    MARS_DTYPE_INT16 = 2,
    MARS_DTYPE_INT8 = 3,
    MARS_DTYPE_UINT8 = 4,
    MARS_DTYPE_UINT4 = 5
} mars_dtype_t;

/* layout tags (reference include/mars.h:46-56).  The executor only ever tests
 * `== MARS_FORMAT_NHWC` (conv) and the two NNA-native tags (byte-size rule). */
typedef enum {
    MARS_FORMAT_NCHW = 0,
    MARS_FORMAT_NDHWC32 = 1,
    MARS_FORMAT_HWIO = 2,
    MARS_FORMAT_NMHWSOIB2 = 3,
    MARS_FORMAT_NMC32 = 4,
    MARS_FORMAT_D1 = 5,
    MARS_FORMAT_OHWI = 6,
    MARS_FORMAT_NHWC = 7,
    MARS_FORMAT_OIHW = 8
} mars_format_t;

/* layer opcodes (reference include/mars.h:59-79) */
typedef enum {
    MARS_LAYER_CONV2D = 0,
    MARS_LAYER_DEPTHWISE_CONV2D = 1,
    MARS_LAYER_MAXPOOL = 2,
    MARS_LAYER_AVGPOOL = 3,
    MARS_LAYER_GLOBAL_AVGPOOL = 4,
    MARS_LAYER_RELU = 5,
    MARS_LAYER_RELU6 = 6,
    MARS_LAYER_LEAKY_RELU = 7,
    MARS_LAYER_SILU = 8,
    MARS_LAYER_SIGMOID = 9,
    MARS_LAYER_CONCAT = 10,
    MARS_LAYER_ADD = 11,
    MARS_LAYER_MUL = 12,
    MARS_LAYER_UPSAMPLE = 13,
    MARS_LAYER_RESHAPE = 14,
    MARS_LAYER_SOFTMAX = 15,
    MARS_LAYER_FC = 16,
    MARS_LAYER_TRANSPOSE = 17,
    MARS_LAYER_BATCHNORM = 18
} mars_layer_type_t;

/* fused activation selector of conv/fc params (reference include/mars.h:82-91) */
typedef enum {
    MARS_ACT_NONE = 0,
    MARS_ACT_RELU = 1,
    MARS_ACT_RELU6 = 2,
    MARS_ACT_LEAKY_RELU = 3,
    MARS_ACT_SILU = 4,
    MARS_ACT_SIGMOID = 5,
    MARS_ACT_TANH = 6,
    MARS_ACT_HARD_SWISH = 7
} mars_activation_t;

/* padding selector (reference include/mars.h:94-98) */
typedef enum {
    MARS_PAD_VALID = 0,
    MARS_PAD_SAME = 1,
    MARS_PAD_EXPLICIT = 2
} mars_padding_t;

#define MARS_PACKED __attribute__((packed))

/* file header, 76 bytes (reference include/mars.h:103-116) */
typedef struct MARS_PACKED {
    uint32_t magic;
    uint16_t version_major;
    uint16_t version_minor;
    uint32_t flags;
    uint32_t num_layers;
    uint32_t num_tensors;
    uint32_t num_inputs;
    uint32_t num_outputs;
    uint64_t weights_offset; /* file offset of the weight blob */
    uint64_t weights_size;   /* bytes in the weight blob */
    uint32_t input_tensor_ids[4];
    uint32_t output_tensor_ids[4];
} mars_header_t;

/* tensor descriptor, 124 bytes (reference include/mars.h:121-132) */
typedef struct MARS_PACKED {
    uint32_t id;
    char name[MARS_MAX_NAME_LEN - 4];
    mars_dtype_t dtype;
    mars_format_t format;
    uint32_t ndims;
    int32_t shape[MARS_MAX_DIMS];
    uint64_t data_offset; /* offset inside the weight blob; meaningless for runtime tensors */
    uint64_t data_size;   /* 0 => runtime (activation) tensor */
    float scale;
    int32_t zero_point;
} mars_tensor_t;

/* conv / depthwise parameters, 60 bytes (reference include/mars.h:139-155) */
typedef struct MARS_PACKED {
    uint32_t kernel_h, kernel_w;
    uint32_t stride_h, stride_w;
    uint32_t dilation_h, dilation_w;
    mars_padding_t padding;
    uint32_t pad_top, pad_bottom, pad_left, pad_right;
    uint32_t groups;
    mars_activation_t activation;
    uint32_t weight_tensor_id;
    uint32_t bias_tensor_id; /* 0xFFFFFFFF => no bias */
} mars_conv_params_t;

/* pooling parameters (reference include/mars.h:158-168) */
typedef struct MARS_PACKED {
    uint32_t kernel_h, kernel_w;
    uint32_t stride_h, stride_w;
    mars_padding_t padding;
    uint32_t pad_top, pad_bottom, pad_left, pad_right;
} mars_pool_params_t;

typedef struct MARS_PACKED { float alpha; } mars_act_params_t;
typedef struct MARS_PACKED { uint32_t axis; uint32_t num_inputs; } mars_concat_params_t;
typedef struct MARS_PACKED { uint32_t scale_h, scale_w, mode; } mars_upsample_params_t;
typedef struct MARS_PACKED { int32_t new_shape[MARS_MAX_DIMS]; uint32_t ndims; } mars_reshape_params_t;
typedef struct MARS_PACKED {
    uint32_t weight_tensor_id;
    uint32_t bias_tensor_id;
    mars_activation_t activation;
} mars_fc_params_t;

/* layer descriptor, 112 bytes (reference include/mars.h:204-221) */
typedef struct MARS_PACKED {
    uint32_t id;
    mars_layer_type_t type;
    uint32_t num_inputs;
    uint32_t num_outputs;
    uint32_t input_tensor_ids[4];
    uint32_t output_tensor_ids[4];
    union {
        mars_conv_params_t conv;
        mars_pool_params_t pool;
        mars_act_params_t act;
        mars_concat_params_t concat;
        mars_upsample_params_t upsample;
        mars_reshape_params_t reshape;
        mars_fc_params_t fc;
        uint8_t raw[64];
    } params;
} mars_layer_t;

#if defined(__cplusplus)
static_assert(sizeof(mars_header_t) == 76, "mars_header_t ABI");
static_assert(sizeof(mars_tensor_t) == 124, "mars_tensor_t ABI");
static_assert(sizeof(mars_layer_t) == 112, "mars_layer_t ABI");
#else
_Static_assert(sizeof(mars_header_t) == 76, "mars_header_t ABI");
_Static_assert(sizeof(mars_tensor_t) == 124, "mars_tensor_t ABI");
_Static_assert(sizeof(mars_layer_t) == 112, "mars_layer_t ABI");
#endif

/* File = header | tensor descriptors | layer descriptors | weight blob
 * (weights_offset = 76 + 124*T + 112*L in compiler output). */

#ifdef __cplusplus
}
#endif
#endif /* MARS_H */
