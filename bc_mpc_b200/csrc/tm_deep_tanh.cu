// Cluster-split tcgen05 rollout kernel (hidden layers up to 4 x 1024: BASELINE.json configs[4]), tanh hidden layers, production flavour
#include "rollout_deep.cuh"

namespace bcmpc {
namespace dpx {
BCMPC_DEEP_DEFINE(deep_launch_tanh_prod, false, true)
}  // namespace dpx
}  // namespace bcmpc
