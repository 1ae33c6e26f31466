/*
 * conv_tc.h -- host interface of the tcgen05 int8 implicit-GEMM convolution (conv_tc.cu).
 * A TcPlan holds what is fixed per (layer, arena): the TMA tensor maps over the NCHW
 * activation planes of every image slot and over the repacked weights, the tile shape and
 * the fused-epilogue description.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "mars_internal.h"

namespace marsb200 {

struct ArenaGeom {
    uint8_t *d_weights;
    uint8_t *d_slots;
    size_t W, slot_stride;
    int capacity;
};

struct TcPlan {
    bool valid = false;
    void *impl = nullptr; /* TcPlanImpl*, owned */
};

/* true when the op can run on the tensor-core kernel; fills `plan` */
bool tc_plan(const Op &o, const ArenaGeom &g, const uint8_t *d_cpool, TcPlan *plan);
bool tc_launch(const TcPlan &plan, int first, int n, cudaStream_t s);
void tc_release(std::vector<TcPlan> &plans);

} // namespace marsb200
