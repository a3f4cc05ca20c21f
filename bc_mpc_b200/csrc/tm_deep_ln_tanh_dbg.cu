// Cluster-split tcgen05 rollout kernel, tanh dynamics with LayerNorm (the reference builder's default activation + FLAGS.LAYER_NORM), debug flavour
#include "rollout_deep.cuh"

namespace bcmpc {
namespace dpx {
BCMPC_DEEP_DEFINE_LN(deep_launch_ln_tanh_dbg, true, true, true)
}  // namespace dpx
}  // namespace bcmpc
