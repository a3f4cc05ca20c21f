// Cluster-split tcgen05 rollout kernel with the proposal network, tanh dynamics, debug flavour
#include "rollout_deep.cuh"

namespace bcmpc {
namespace dpx {
BCMPC_DEEP_DEFINE_POL(deep_launch_pol_tanh_dbg, true, true)
}  // namespace dpx
}  // namespace bcmpc
