// tcgen05 rollout kernel: variant 27 (bit flags: 1 = proposal network, 2 = LayerNorm dynamics, 4 = learned-reward model,
// 8 = explicit biases, 16 = tanh, 32 = proposal network of up to 256 units per hidden layer), debug flavour (injected randomness, per-path outputs)
#include "rollout_tmem.cuh"

namespace bcmpc {
namespace tmx {
BCMPC_TM_DEFINE_VARIANT(27, true)
}  // namespace tmx
}  // namespace bcmpc
