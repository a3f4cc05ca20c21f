// Cluster-split tcgen05 rollout kernel (hidden layers up to 4 x 1024: BASELINE.json configs[4]), tanh hidden layers, debug flavour
#include "rollout_deep.cuh"

namespace bcmpc {
namespace dpx {
BCMPC_DEEP_DEFINE(deep_launch_tanh_dbg, true, true)
}  // namespace dpx
}  // namespace bcmpc
