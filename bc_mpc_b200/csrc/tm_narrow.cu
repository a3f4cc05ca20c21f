// Narrow-net tcgen05 rollout kernel (hidden <= 256, two tiles per CTA), production flavour
#include "rollout_narrow.cuh"

namespace bcmpc {
namespace tmx {
template <bool LN, bool POL>
static cudaError_t nw_pick(int nc, const TmArgs& a, dim3 grid, cudaStream_t st) {
  switch (nc) {
    case 1: return nw_launch_t<1, LN, POL, false>(a, grid, st);
    case 2: return nw_launch_t<2, LN, POL, false>(a, grid, st);
    case 3: return nw_launch_t<3, LN, POL, false>(a, grid, st);
    case 4: return nw_launch_t<4, LN, POL, false>(a, grid, st);
  }
  return cudaErrorInvalidValue;
}
cudaError_t nw_launch_prod(int nc, bool ln, bool pol, const TmArgs& a, dim3 grid, cudaStream_t st) {
  if (pol) return ln ? nw_pick<true, true>(nc, a, grid, st) : nw_pick<false, true>(nc, a, grid, st);
  return ln ? nw_pick<true, false>(nc, a, grid, st) : nw_pick<false, false>(nc, a, grid, st);
}
}  // namespace tmx
}  // namespace bcmpc
