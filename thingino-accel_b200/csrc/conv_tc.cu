/* conv_tc.cu -- placeholder until the tcgen05 kernel lands (every op stays on the direct path) */
#include "conv_tc.h"
namespace marsb200 {
bool tc_plan(const Op &, const ArenaGeom &, const uint8_t *, TcPlan *) { return false; }
bool tc_launch(const TcPlan &, int, int, cudaStream_t) { return false; }
void tc_release(std::vector<TcPlan> &plans) { plans.clear(); }
} // namespace marsb200
