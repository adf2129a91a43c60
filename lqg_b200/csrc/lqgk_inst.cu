// lqgk_inst.cu -- instantiates every kernel and the launch sequence for ONE dimension tuple, selected with
// -DLQGK_INST_X/B/U/Y/D (one object file per line of lqgk_dims.h; __graft_entry__.build() compiles them in parallel).
#include "lqgk_run.cuh"

#if !defined(LQGK_INST_X)
#error "compile with -DLQGK_INST_X=.. -DLQGK_INST_B=.. -DLQGK_INST_U=.. -DLQGK_INST_Y=.. -DLQGK_INST_D=.."
#endif

namespace lqgk {
using DMI = Dims<LQGK_INST_X, LQGK_INST_B, LQGK_INST_U, LQGK_INST_Y, LQGK_INST_D>;
template <>
int Runner<LQGK_INST_X, LQGK_INST_B, LQGK_INST_U, LQGK_INST_Y, LQGK_INST_D>::run_f32(const Call& c) { return run<DMI, float>(c); }
template <>
int Runner<LQGK_INST_X, LQGK_INST_B, LQGK_INST_U, LQGK_INST_Y, LQGK_INST_D>::run_f64(const Call& c) { return run<DMI, double>(c); }
template <>
size_t Runner<LQGK_INST_X, LQGK_INST_B, LQGK_INST_U, LQGK_INST_Y, LQGK_INST_D>::plan_bytes(const LqgkDims& d, int mode, int32_t max_chunk) {
  size_t Sc = choose_chunk<DMI>(d, mode, false, (size_t)-1, max_chunk);
  return make_plan<DMI>(d, mode, false, Sc).bytes;
}
}  // namespace lqgk
