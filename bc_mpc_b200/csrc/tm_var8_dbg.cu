// tcgen05 rollout kernel: variant 8 (bit flags: 1 = proposal network, 2 = LayerNorm dynamics, 4 = learned-reward model,
// 8 = explicit biases), debug flavour (injected randomness, per-path outputs)
#include "rollout_tmem.cuh"

namespace bcmpc {
namespace tmx {
BCMPC_TM_DEFINE_VARIANT(8, true)
}  // namespace tmx
}  // namespace bcmpc
