// tcgen05 rollout kernel: variant 3 (bit flags: 1 = proposal network, 2 = LayerNorm dynamics, 4 = learned-reward model),
// debug flavour (injected randomness, per-path outputs)
#include "rollout_tmem.cuh"

namespace bcmpc {
namespace tmx {
BCMPC_TM_DEFINE_VARIANT(3, true)
}  // namespace tmx
}  // namespace bcmpc
