// tcgen05 rollout kernel, variant 6 (bit flags: 1 = proposal network, 2 = LayerNorm dynamics, 4 = learned-reward model)
#include "rollout_tmem.cuh"

namespace bcmpc {
namespace tmx {
cudaError_t tm_launch_var6(int cg, int nc, const TmArgs& a, dim3 grid, cudaStream_t st) {
  return cg == 2 ? tm_launch_48<2, 6>(nc, a, grid, st) : tm_launch_48<1, 6>(nc, a, grid, st);
}
}  // namespace tmx
}  // namespace bcmpc
