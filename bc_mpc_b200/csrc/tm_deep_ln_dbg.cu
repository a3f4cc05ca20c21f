// Cluster-split tcgen05 rollout kernel, LayerNorm dynamics (FLAGS.LAYER_NORM with n_layers > 2 or size > 510), debug flavour
#include "rollout_deep.cuh"

namespace bcmpc {
namespace dpx {
BCMPC_DEEP_DEFINE_LN(deep_launch_ln_dbg, true, false, true)
}  // namespace dpx
}  // namespace bcmpc
