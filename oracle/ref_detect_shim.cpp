/*
 * oracle/ref_detect_shim.cpp -- TEST INFRASTRUCTURE.
 * Reaches the file-static helpers of reference examples/yolo_detect.cpp
 * (sigmoid :133-135, iou :138-149, nms :152-173, ANCHORS/STRIDES :176-181,
 * scale_detections :208-227, load_and_preprocess_image :72-130) by including that TU with main() renamed.
 */
#define main static __attribute__((unused)) ref_yolo_detect_main
#include "examples/yolo_detect.cpp"
#undef main

extern "C" {
int oracle_ref_cpp_det_size(void) { return (int)sizeof(Detection); }
float oracle_ref_cpp_sigmoid(float x) { return sigmoid(x); }
float oracle_ref_cpp_iou(const Detection *a, const Detection *b) { return iou(*a, *b); }
int oracle_ref_cpp_nms(Detection *d, int n, float thresh) {
    std::vector<Detection> v(d, d + n);
    nms(v, thresh);
    for (size_t i = 0; i < v.size(); i++) d[i] = v[i];
    return (int)v.size();
}
void oracle_ref_cpp_scale_detections(Detection *d, int n, int orig_w, int orig_h) {
    std::vector<Detection> v(d, d + n);
    scale_detections(v, orig_w, orig_h);
    for (int i = 0; i < n; i++) d[i] = v[i];
}
/* the reference's own load_and_preprocess_image() (:72-130): stb decode of the file, stbir letterbox resize into a
 * 640x640 RGBA frame, gray 114 border.  out = 640*640*4 bytes. */
int oracle_ref_cpp_preprocess(const char *path, unsigned char *out, int *ow, int *oh) {
    uint8_t *p = load_and_preprocess_image(path, ow, oh);
    if (!p) return -1;
    memcpy(out, p, (size_t)YOLO_INPUT_WIDTH * YOLO_INPUT_HEIGHT * 4);
    free(p);
    return 0;
}
void oracle_ref_cpp_anchors(float *anchors18, int *strides3) {
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 6; j++) anchors18[i * 6 + j] = ANCHORS[i][j];
        strides3[i] = STRIDES[i];
    }
}
}
