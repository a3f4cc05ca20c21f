// Narrow-net tcgen05 rollout kernel (hidden <= 256, two tiles per CTA), production flavour
#include "rollout_narrow.cuh"

namespace bcmpc {
namespace tmx {
cudaError_t nw_launch_prod(int nc, const TmArgs& a, dim3 grid, cudaStream_t st) {
  switch (nc) {
    case 1: return nw_launch_t<1, false>(a, grid, st);
    case 2: return nw_launch_t<2, false>(a, grid, st);
    case 3: return nw_launch_t<3, false>(a, grid, st);
    case 4: return nw_launch_t<4, false>(a, grid, st);
  }
  return cudaErrorInvalidValue;
}
}  // namespace tmx
}  // namespace bcmpc
