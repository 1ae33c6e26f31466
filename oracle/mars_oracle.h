/*
 * oracle/mars_oracle.h -- TEST INFRASTRUCTURE (see mars_oracle.c header).
 * CPU restatement of the reference mars hot path used only as the parity checker.
 */
#ifndef MARS_ORACLE_H
#define MARS_ORACLE_H

#include "../include/mars.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { /* numerically equal to mars_error_t (reference include/mars_runtime.h:19-29) */
    MO_OK = 0, MO_ERR_INVALID_MAGIC = -1, MO_ERR_VERSION_MISMATCH = -2, MO_ERR_ALLOC_FAILED = -3,
    MO_ERR_INVALID_FILE = -4, MO_ERR_LAYER_FAILED = -6, MO_ERR_INVALID_TENSOR = -7, MO_ERR_INVALID_LAYER = -8
};

typedef struct mo_model mo_model_t;

/* det_t of reference src/mars/mars_yolo_test.c:37 (centre boxes) */
typedef struct { float x, y, w, h, conf; int32_t cls; } mo_det_t;
/* Detection of reference examples/yolo_detect.cpp:39-43 (corner boxes) */
typedef struct { float x0, y0, x1, y1, confidence; int32_t class_id; } mo_box_t;

int mo_load(const void *blob, size_t size, size_t arena_bytes, mo_model_t **out);
void mo_free(mo_model_t *m);
int mo_run(mo_model_t *m);
int mo_run_layer(mo_model_t *m, uint32_t i);
void mo_set_depthwise_mode(mo_model_t *m, int mode);

uint8_t *mo_arena(mo_model_t *m);
size_t mo_arena_size(const mo_model_t *m);
size_t mo_weights_size(const mo_model_t *m);
size_t mo_buffer_size(const mo_model_t *m);
int mo_num_buffers(const mo_model_t *m);
uint32_t mo_num_layers(const mo_model_t *m);
uint32_t mo_num_tensors(const mo_model_t *m);
const mars_tensor_t *mo_tensor_desc(const mo_model_t *m, uint32_t idx);
const mars_layer_t *mo_layer_desc(const mo_model_t *m, uint32_t idx);
size_t mo_tensor_offset(const mo_model_t *m, uint32_t idx);
size_t mo_tensor_alloc(const mo_model_t *m, uint32_t idx);
int mo_input_index(const mo_model_t *m, int i);
int mo_output_index(const mo_model_t *m, int i);
size_t mo_tensor_byte_size(const mars_tensor_t *t);

int mo_parse_output(const int8_t *data, int npred, float scale, mo_det_t *dets, int maxd);
int mo_nms(mo_det_t *d, int n, float thresh);
float mo_iou_corner(const mo_box_t *a, const mo_box_t *b);
int mo_nms_corner(mo_box_t *d, int n, float thresh);
void mo_scale_detections(mo_box_t *d, int n, int orig_w, int orig_h, int net_w, int net_h);
int mo_decode_anchor_grid(const int8_t *head, int gh, int gw, float scale, int level, float conf_thresh,
                          mo_box_t *dets, int cnt, int maxd);

void mo_vec_add_f32(float *dst, const float *a, const float *b, size_t n);
void mo_vec_sub_f32(float *dst, const float *a, const float *b, size_t n);
void mo_vec_mul_f32(float *dst, const float *a, const float *b, size_t n);
void mo_vec_relu_f32(float *dst, const float *a, size_t n);
float mo_vec_dot_f32(const float *a, const float *b, size_t n);
void mo_matmul_f32(float *C, const float *A, const float *B, size_t M, size_t K, size_t N);

#ifdef __cplusplus
}
#endif
#endif
