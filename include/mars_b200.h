/*
 * mars_b200.h -- additive C-ABI entry points of libmars_b200.so (B200 build).
 *
 * Nothing here changes a signature of the reference API (include/mars_runtime.h,
 * include/mars_math.h, include/nna.h); these functions expose what the reference
 * does not have: an image batch sharded over HBM "slots", device-resident runs,
 * and the YOLO post-process (reference file-static functions) as library calls.
 * Plain pointers and sizes only -- no torch / CUDA types cross this boundary.
 */
#ifndef MARS_B200_H
#define MARS_B200_H

#include "mars_runtime.h"

#ifdef __cplusplus
extern "C" {
#endif

/* det_t of reference src/mars/mars_yolo_test.c:37 (centre box, 24 bytes) */
typedef struct { float x, y, w, h, conf; int32_t cls; } mars_det_t;
/* Detection of reference examples/yolo_detect.cpp:39-43 (corner box, 24 bytes) */
typedef struct { float x0, y0, x1, y1, confidence; int32_t class_id; } mars_box_t;

/* ---- library / device ---------------------------------------------------- */
const char *mars_b200_version(void);
/* number of visible CUDA devices, or <0 (MARS_ERR_NNA_INIT_FAILED) without a driver */
int mars_b200_device_count(void);
/* select the device used by subsequent mars_load_* calls (default: MARS_DEVICE env or 0) */
int mars_b200_set_device(int ordinal);
/* arena bytes for subsequent loads; 0 = the reference's 8 MiB literal
 * (src/mars/mars_runtime.c:209).  Also settable with MARS_ARENA_BYTES. */
void mars_b200_set_arena_bytes(size_t bytes);
/* last diagnostic of the calling thread's most recent failing call */
const char *mars_b200_last_error(void);

/* ---- whole-arena mirror control (strict drop-in / tests) ---------------- */
/* copy the host mirror (ddr_base) of image slot `slot` to the device arena / back.
 * mars_run() itself only moves the work buffers holding model inputs / outputs. */
mars_error_t mars_b200_arena_upload(mars_model_t *m, int slot);
mars_error_t mars_b200_arena_download(mars_model_t *m, int slot, void *host_dst, size_t bytes);
/* zero every work buffer of every slot (the reference never clears them; parity
 * protocol = zero at load, then compare run k against run k) */
mars_error_t mars_b200_arena_clear(mars_model_t *m);
/* run a single layer on slot 0 (per-layer parity tests; mirrors oracle_ref_run_layer) */
mars_error_t mars_b200_run_layer(mars_model_t *m, uint32_t layer);
/* per-layer schedule summary as text: kernel kind, hazard class, fusion decision */
size_t mars_b200_describe(mars_model_t *m, char *dst, size_t cap);
/* 0 = exact direct kernels only; 1 = + tcgen05 convolutions and vectorised memory-bound kernels,
 * one device op per layer (per-layer semantics kept); 2 = + SIGMOID/MUL folded into conv epilogues (whole arena
 * still byte-identical to the reference's); 3 = + stores that no later layer, no later run and no output
 * buffer can observe are dropped from fused epilogues (default; every model OUTPUT stays bit-exact) */
void mars_b200_set_opt_level(mars_model_t *m, int level);
/* 0 = reference semantics (DEPTHWISE_CONV2D is a no-op, src/mars/mars_runtime.c:1168-1170),
 * 1 = restated depthwise convolution (parity unpinned) */
void mars_b200_set_depthwise_mode(mars_model_t *m, int mode);
/* strict mode (also MARS_STRICT=1): when a convolution cannot be planned on the tensor-core kernel the library normally
 * recompiles the whole model on the exact direct kernels (correct, ~100x slower) and says so on stderr; in strict mode the
 * call that compiles the model (mars_load_*, mars_b200_set_batch, ...) fails with MARS_ERR_LAYER_FAILED instead */
void mars_b200_set_strict(int on);
/* float32 convolutions (reference conv2d_float32_mxu, src/mars/mxu_conv.c:673-710): 2 (default; also MARS_F32_MODE) = tcgen05
 * kind::tf32 with a hi/lo operand split, three MMAs per k-step (fp32-grade products, results within ~1e-6 relative of the
 * reference's sequential fp32 sum); 1 = plain tf32 (operands rounded to 10 mantissa bits, ~5e-4 per product); 0 = the
 * exact-order fp32 kernel, bit-identical to the reference (the control) */
void mars_b200_set_f32_mode(mars_model_t *m, int mode);

/* ---- image batch ---------------------------------------------------------- */
/* allocate `capacity` image slots (each = one set of work buffers; weights shared) */
mars_error_t mars_b200_set_batch(mars_model_t *m, int capacity);
int mars_b200_get_batch(mars_model_t *m);
/* bytes of one image's input tensor 0 / output tensor 0 (numel * elem size) */
size_t mars_b200_input_bytes(mars_model_t *m);
size_t mars_b200_output_bytes(mars_model_t *m);
/* host -> slots [first, first+n): n inputs, `stride` bytes apart in host memory */
mars_error_t mars_b200_upload_inputs(mars_model_t *m, int first, int n, const void *host, size_t stride);
/* Pre-processing in front of mars_run (what the reference's caller does on the CPU, src/mars/mars_yolo_test.c:40-77 and
 * :157-165): n RGB8 frames of w x h pixels in host memory, frame_stride bytes apart, are letterboxed on the GPU into input
 * tensor 0 of slots [first, first+n): stbir_resize_uint8 (include/stb/stb_image_resize.h) semantics bit for bit, gray border
 * -17, px - 128, NCHW or NHWC as the tensor's format field says.  Replaces load_image + mars_b200_upload_inputs. */
mars_error_t mars_b200_preprocess_batch(mars_model_t *m, int first, int n, const uint8_t *frames, size_t frame_stride, int w, int h);
/* slots -> host: output tensor 0 of n images */
mars_error_t mars_b200_download_outputs(mars_model_t *m, int first, int n, void *host, size_t stride);
/* run all layers for slots [first, first+n) with inputs already resident in HBM */
mars_error_t mars_b200_run_resident(mars_model_t *m, int first, int n);
/* decode + NMS on the device for slots [first, first+n); results stay resident */
mars_error_t mars_b200_detect_resident(mars_model_t *m, int first, int n, float nms_thresh);
/* copy the detections of slots [first, first+n) to the host:
 * dets[n][maxd] (maxd <= 1000), counts[n] */
mars_error_t mars_b200_download_detections(mars_model_t *m, int first, int n, mars_det_t *dets, int32_t *counts, int maxd);
/* end to end from host buffers: H2D, all layers, decode, NMS, D2H (pipelined in chunks) */
mars_error_t mars_b200_detect_batch(mars_model_t *m, int n, const void *inputs, size_t in_stride,
                                    mars_det_t *dets, int32_t *counts, int maxd, float nms_thresh);
/* asynchronous form: queue a batch of up to capacity/2 images on half `pool` (0 or 1) of the slot pool and return;
 * mars_b200_wait_batch(pool) blocks until its detections are in dets/counts.  Submitting to one half while the other
 * computes overlaps the host->device copy and the read-back with the kernels.  Buffers must stay valid (pinned memory,
 * e.g. nna_malloc, for true overlap) until the wait returns. */
mars_error_t mars_b200_submit_batch(mars_model_t *m, int pool, int n, const void *inputs, size_t in_stride,
                                    mars_det_t *dets, int32_t *counts, int maxd, float nms_thresh);
mars_error_t mars_b200_wait_batch(mars_model_t *m, int pool);
/* the same without the YOLO post-process: output tensor 0 of every image is read back (out_stride bytes apart); waited for
 * with mars_b200_wait_batch like a detection batch */
mars_error_t mars_b200_submit_run_batch(mars_model_t *m, int pool, int n, const void *inputs, size_t in_stride, void *outputs, size_t out_stride);
/* end to end from host buffers returning raw output tensor 0 per image */
mars_error_t mars_b200_run_batch(mars_model_t *m, int n, const void *inputs, size_t in_stride,
                                 void *outputs, size_t out_stride);
/* run all layers, then (with_detect) decode + NMS, as ONE device-timed region */
mars_error_t mars_b200_step_resident(mars_model_t *m, int first, int n, float nms_thresh, int with_detect);
/* the same step without the host synchronisation: returns once the work is enqueued on the model's compute stream
 * (mars_b200_compute_stream); the caller orders its own work behind it there -- e.g. the NCCL gather of the records -- and waits itself */
mars_error_t mars_b200_enqueue_step_resident(mars_model_t *m, int first, int n, float nms_thresh, int with_detect);
/* device addresses of the resident detection records, for a device-side gather (NCCL):
 * dets = mars_det_t[capacity][*stride_dets], counts = int32[capacity] */
void mars_b200_detections_device(mars_model_t *m, void **dets, void **counts, int *stride_dets);

/* ---- several GPUs of one box behind one call (batch sharded over the devices, no inter-GPU traffic on the path) ---------- */
typedef struct mars_b200_group mars_b200_group_t;
/* one replica of the model per device (devices = NULL: ordinals 0 .. n_devices-1), `per_gpu_batch` image slots each */
mars_error_t mars_b200_group_load(const void *data, size_t size, const int *devices, int n_devices, int per_gpu_batch, mars_b200_group_t **out);
void mars_b200_group_free(mars_b200_group_t *g);
int mars_b200_group_size(mars_b200_group_t *g);
/* the replica on the i-th device of the group (for per-device settings: opt level, f32 mode, ...) */
mars_model_t *mars_b200_group_model(mars_b200_group_t *g, int i);
/* mars_b200_detect_batch over the whole group: image b runs on GPU b / ceil(n / G), one host thread per GPU; dets[n][maxd] and
 * counts[n] are filled directly by every GPU's read-back (the single-process form of the detection gather) */
mars_error_t mars_b200_group_detect_batch(mars_b200_group_t *g, int n, const void *inputs, size_t in_stride, mars_det_t *dets, int32_t *counts,
                                          int maxd, float nms_thresh);

/* ---- introspection (tests, bench) ----------------------------------------- */
/* record CUDA events around every device op of full passes; read back with mars_b200_op_info */
void mars_b200_set_profile(mars_model_t *m, int on);
int mars_b200_num_ops(mars_model_t *m);
/* info[0..12] = kind, layer, impl, mode, ic, oc, oh, ow, kh, kw, fused_layers, ih, iw */
int mars_b200_op_info(mars_model_t *m, int op, int32_t *info, double *ms, uint64_t *calls, uint64_t *flat_n);
/* out5 = weights_size, buffer_size, num_buffers, slot_stride, arena_size */
void mars_b200_geometry(mars_model_t *m, size_t *out5);
/* arena offset of tensor table entry `index` (reference planner, src/mars/mars_runtime.c:248-337) */
size_t mars_b200_tensor_offset(mars_model_t *m, uint32_t index);

/* the CUDA stream (cudaStream_t) the model's kernels run on; work enqueued there by the caller (e.g. an NCCL gather of the
 * resident detection records) is ordered between the batches */
void *mars_b200_compute_stream(mars_model_t *m);
/* kernels launched by this model since load (the bench's gpu_launches claim) */
uint64_t mars_b200_launch_count(mars_model_t *m);
/* CUDA-event time of the last mars_b200_run_resident / detect_resident, milliseconds */
float mars_b200_last_gpu_ms(mars_model_t *m);

/* ---- YOLO pre-process on host buffers ------------------------------------ */
/* load_image of reference src/mars/mars_yolo_test.c:40-77 without the file decode: rgb = h x w x 3 bytes in,
 * out = tw x th x 3 int8 (three planes, or interleaved when nhwc != 0).  Returns 0 on success, -1 on failure. */
int mars_b200_letterbox(const uint8_t *rgb, int w, int h, int tw, int th, int nhwc, int8_t *out);
/* load_and_preprocess_image of reference examples/yolo_detect.cpp:72-130 without the file decode: the same resize into an
 * RGBA uint8 frame, out = tw x th x 4 bytes (the reference fixes 640 x 640): 114 everywhere outside the image (alpha
 * included), R, G, B, 0 inside.  Returns 0 on success, -1 on failure. */
int mars_b200_letterbox_rgba(const uint8_t *rgb, int w, int h, int tw, int th, uint8_t *out);
/* host half of the above (no GPU needed): the (source sample, coefficient) taps each of the out_size output samples of one
 * axis sums, in the order stbir_resize_uint8 adds them.  start[out_size + 1]; returns the tap count (> cap: nothing copied). */
int mars_b200_resize_taps(int in_size, int out_size, int32_t *start, int32_t *src, float *w, int cap);

/* ---- planner (host half, no GPU needed) ----------------------------------- */
/* What the library would do with a .mars blob: the compiled op list at `opt_level` (0 exact kernels ... 3 default), one line per
 * device op with its kernel choice and the fused / linked / forwarded / elided output streams, written to dst (NUL terminated,
 * truncated to cap).  arena_bytes = 0: the reference's 8 MiB (src/mars/mars_runtime.c:209).  Returns the full length, 0 on error. */
size_t mars_b200_plan_describe(const void *data, size_t size, size_t arena_bytes, int opt_level, char *dst, size_t cap);

/* ---- conv requantisation in integers (host half, no GPU needed) ----------- */
/* The tcgen05 conv epilogue replaces the float requantisation of reference src/mars/mxu_conv.c:663-666,
 * r = clamp((int32)(fl((float)t * scale) +- 0.5f)), by clamp(floor((t * m + c) / 2^(32 + s))) when integers (m, s, c) exist that
 * reproduce it for EVERY |t| <= tmax (t = accumulator + bias).  Returns 1 and fills m / s / c, or 0 when no such triple exists
 * (the layer then keeps a float variant).  Exposed so that the CPU test suite can check the fit by brute force. */
int mars_b200_requant_fit(float scale, long long tmax, int *m, int *s, long long *c);
/* the reference's requantisation of one value, restated (what the fit is checked against) */
int mars_b200_requant_ref(int t, float scale);

/* ---- NNA-native tensor layouts -------------------------------- */
/* The layouts behind the format tags MARS_FORMAT_NMHWSOIB2 / MARS_FORMAT_NDHWC32 (reference include/mars.h:46-56), which the
 * reference's compiler can emit and its runtime only sizes (src/mars/mars_runtime.c:93-110):
 *   NMHWSOIB2 weights  [ceil(Co/32)][ceil(Ci/32)][KH][KW][32 out][32 in], 1024-byte blocks, absent channels zero
 *                      (packer: mars-compiler/src/mars_format.rs:436-470; unpackers: mgk-decompiler/mgk_decompiler.py:530-540,
 *                      mgk-decompiler/scripts/extract_weights_nmhwsoib2.py:52-80)
 *   NDHWC32 features   [N][ceil(C/32)][H][W][32], absent channels zero (mars-compiler/src/mars_format.rs:490-531)
 * Byte permutations on the device; host-buffer forms (copy in, convert, copy out) and device-resident forms (`stream` is a
 * cudaStream_t or NULL; the NDHWC32 buffer must be 16-byte aligned).  All return 0 on success, -1 on error
 * (mars_b200_last_error); the size helpers return 0 for non-positive dimensions. */
size_t mars_b200_nmhwsoib2_size(int out_ch, int in_ch, int kh, int kw);            /* mars_format.rs:472-476 */
size_t mars_b200_ndhwc32_size(int batch, int channels, int height, int width);     /* mars_format.rs:485-488 */
int mars_b200_pack_weights_nmhwsoib2(const int8_t *oihw, int out_ch, int in_ch, int kh, int kw, uint8_t *packed);
int mars_b200_unpack_weights_nmhwsoib2(const uint8_t *packed, int out_ch, int in_ch, int kh, int kw, int8_t *oihw);
int mars_b200_nchw_to_ndhwc32(const uint8_t *nchw, int batch, int channels, int height, int width, uint8_t *out);
int mars_b200_ndhwc32_to_nchw(const uint8_t *native, int batch, int channels, int height, int width, uint8_t *nchw);
int mars_b200_pack_weights_nmhwsoib2_device(const int8_t *d_oihw, int out_ch, int in_ch, int kh, int kw, uint8_t *d_packed, void *stream);
int mars_b200_unpack_weights_nmhwsoib2_device(const uint8_t *d_packed, int out_ch, int in_ch, int kh, int kw, int8_t *d_oihw, void *stream);
int mars_b200_nchw_to_ndhwc32_device(const uint8_t *d_nchw, int batch, int channels, int height, int width, uint8_t *d_out, void *stream);
int mars_b200_ndhwc32_to_nchw_device(const uint8_t *d_native, int batch, int channels, int height, int width, uint8_t *d_nchw, void *stream);

/* ---- YOLO post-process on host buffers ----------------------------------- */
/* parse_output of reference src/mars/mars_yolo_test.c:80-104 (conf threshold 0.25) */
int mars_yolo_parse_output(const int8_t *data, int npred, float scale, mars_det_t *dets, int maxd);
/* nms of reference src/mars/mars_yolo_test.c:107-130 (exchange sort + greedy, centre boxes) */
int mars_yolo_nms(mars_det_t *dets, int n, float thresh);
/* nms of reference examples/yolo_detect.cpp:152-173 (corner boxes; ties keep input order) */
int mars_yolo_nms_boxes(mars_box_t *dets, int n, float thresh);
/* scale_detections of reference examples/yolo_detect.cpp:208-227 */
void mars_yolo_scale_detections(mars_box_t *dets, int n, int orig_w, int orig_h, int net_w, int net_h);
/* anchor-grid decode (formula: mgk-decompiler/test_yolo_inference.py:136-202, anchors
 * examples/yolo_detect.cpp:176-181).  head = int8 [3,gh,gw,85]; appends after `cnt`. */
int mars_yolo_decode_anchor_grid(const int8_t *head, int gh, int gw, float scale, int level,
                                 float conf_thresh, mars_box_t *dets, int cnt, int maxd);

#ifdef __cplusplus
}
#endif
#endif /* MARS_B200_H */
