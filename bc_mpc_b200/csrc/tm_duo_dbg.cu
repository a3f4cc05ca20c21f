// Two-tiles-per-CTA tcgen05 rollout kernel for wide nets (hidden 257..510, rollout_duo.cuh), debug flavour
#include "rollout_duo.cuh"

namespace bcmpc {
namespace tmx {
template <int VAR>
static cudaError_t du_pick(int nc, const TmArgs& a, dim3 grid, cudaStream_t st) {
  switch (nc) {
    case 5: return du_launch_t<5, VAR, true>(a, grid, st);
    case 6: return du_launch_t<6, VAR, true>(a, grid, st);
    case 7: return du_launch_t<7, VAR, true>(a, grid, st);
    case 8: return du_launch_t<8, VAR, true>(a, grid, st);
  }
  return cudaErrorInvalidValue;
}
cudaError_t du_launch_dbg(int nc, int var, const TmArgs& a, dim3 grid, cudaStream_t st) {
  if (var == 0) return du_pick<0>(nc, a, grid, st);
  if (var == (kVarBias | kVarTanh)) return du_pick<kVarBias | kVarTanh>(nc, a, grid, st);
  return cudaErrorInvalidValue;
}
}  // namespace tmx
}  // namespace bcmpc
