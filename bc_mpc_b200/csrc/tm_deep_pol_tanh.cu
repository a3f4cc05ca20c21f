// Cluster-split tcgen05 rollout kernel with the proposal network, tanh dynamics (the reference builder's default activation), production flavour
#include "rollout_deep.cuh"

namespace bcmpc {
namespace dpx {
BCMPC_DEEP_DEFINE_POL(deep_launch_pol_tanh_prod, false, true)
}  // namespace dpx
}  // namespace bcmpc
