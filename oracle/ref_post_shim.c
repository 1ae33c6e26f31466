/*
 * oracle/ref_post_shim.c -- TEST INFRASTRUCTURE.
 * Pulls in the reference caller TU (src/mars/mars_yolo_test.c) with its main()
 * renamed so that its file-static parse_output() (:80-104) and nms() (:107-130)
 * -- the only working C decode/NMS in the reference -- can be called as the
 * post-process oracle.  The arithmetic executed is the reference's own.
 */
#define main static __attribute__((unused)) ref_yolo_test_main
#include "src/mars/mars_yolo_test.c"
#undef main

#include "mars_runtime.h"

int oracle_ref_parse_output(const int8_t *data, int npred, float scale, void *dets, int maxd) {
    return parse_output(data, npred, scale, (det_t *)dets, maxd);
}
int oracle_ref_nms(void *dets, int n, float thresh) { return nms((det_t *)dets, n, thresh); }
int oracle_ref_det_size(void) { return (int)sizeof(det_t); }

/* the reference's own load_image() (:40-77): stb_image decode of the file (tests write a binary PPM), stbir letterbox
 * resize, int8 packing.  out = tw*th*3 bytes. */
int oracle_ref_load_image(const char *path, int tw, int th, int nhwc, int8_t *out, int *ow, int *oh) {
    int8_t *p = load_image(path, tw, th, nhwc, ow, oh);
    if (!p) return -1;
    memcpy(out, p, (size_t)tw * th * 3);
    free(p);
    return 0;
}

/* run one layer through the reference's own mars_run(): a shallow copy of the
 * model whose layer table starts at layer i and has length 1. */
int oracle_ref_run_layer(mars_model_t *m, uint32_t i) {
    if (!m || i >= m->header.num_layers) return MARS_ERR_INVALID_LAYER;
    mars_model_t one = *m;
    one.layers = m->layers + i;
    one.header.num_layers = 1;
    return (int)mars_run(&one);
}
