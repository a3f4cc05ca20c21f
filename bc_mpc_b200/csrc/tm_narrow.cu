// Narrow-net tcgen05 rollout kernel (hidden <= 256, two tiles per CTA), production flavour
#include "rollout_narrow.cuh"

namespace bcmpc {
namespace tmx {
template <bool LN>
static cudaError_t nw_pick(int nc, const TmArgs& a, dim3 grid, cudaStream_t st) {
  switch (nc) {
    case 1: return nw_launch_t<1, LN, false>(a, grid, st);
    case 2: return nw_launch_t<2, LN, false>(a, grid, st);
    case 3: return nw_launch_t<3, LN, false>(a, grid, st);
    case 4: return nw_launch_t<4, LN, false>(a, grid, st);
  }
  return cudaErrorInvalidValue;
}
cudaError_t nw_launch_prod(int nc, bool ln, const TmArgs& a, dim3 grid, cudaStream_t st) {
  return ln ? nw_pick<true>(nc, a, grid, st) : nw_pick<false>(nc, a, grid, st);
}
}  // namespace tmx
}  // namespace bcmpc
