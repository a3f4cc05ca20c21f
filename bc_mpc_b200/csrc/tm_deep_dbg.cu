// Cluster-split tcgen05 rollout kernel (hidden layers up to 4 x 1024: BASELINE.json configs[4]), debug flavour
#include "rollout_deep.cuh"

namespace bcmpc {
namespace dpx {
BCMPC_DEEP_DEFINE(deep_launch_dbg, true, false)
}  // namespace dpx
}  // namespace bcmpc
