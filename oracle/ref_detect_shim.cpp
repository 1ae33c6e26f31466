/*
 * oracle/ref_detect_shim.cpp -- TEST INFRASTRUCTURE.
 * Reaches the file-static helpers of reference examples/yolo_detect.cpp
 * (sigmoid :133-135, iou :138-149, nms :152-173, ANCHORS/STRIDES :176-181,
 * scale_detections :208-227) by including that TU with main() renamed.
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
void oracle_ref_cpp_anchors(float *anchors18, int *strides3) {
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 6; j++) anchors18[i * 6 + j] = ANCHORS[i][j];
        strides3[i] = STRIDES[i];
    }
}
}
