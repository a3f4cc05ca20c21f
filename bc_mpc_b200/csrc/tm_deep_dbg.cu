// Cluster-split tcgen05 rollout kernel, debug flavour (injected randomness, per-path outputs: tests, bcmpc_rollout_debug, bcmpc_predict)
#include "rollout_deep.cuh"

namespace bcmpc {
namespace dpx {
BCMPC_DEEP_DEFINE(deep_launch_dbg, true)
}  // namespace dpx
}  // namespace bcmpc
