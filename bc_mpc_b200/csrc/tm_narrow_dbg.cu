// Narrow-net tcgen05 rollout kernel (hidden <= 256, two tiles per CTA), debug flavour
#include "rollout_narrow.cuh"

namespace bcmpc {
namespace tmx {
template <bool LN>
static cudaError_t nw_pick(int nc, const TmArgs& a, dim3 grid, cudaStream_t st) {
  switch (nc) {
    case 1: return nw_launch_t<1, LN, true>(a, grid, st);
    case 2: return nw_launch_t<2, LN, true>(a, grid, st);
    case 3: return nw_launch_t<3, LN, true>(a, grid, st);
    case 4: return nw_launch_t<4, LN, true>(a, grid, st);
  }
  return cudaErrorInvalidValue;
}
cudaError_t nw_launch_dbg(int nc, bool ln, const TmArgs& a, dim3 grid, cudaStream_t st) {
  return ln ? nw_pick<true>(nc, a, grid, st) : nw_pick<false>(nc, a, grid, st);
}
}  // namespace tmx
}  // namespace bcmpc
