// tcgen05 rollout kernel: variant 25 (bit flags: 1 = proposal network, 2 = LayerNorm dynamics, 4 = learned-reward model,
// 8 = explicit biases, 16 = plain tanh dynamics), production flavour
#include "rollout_tmem.cuh"

namespace bcmpc {
namespace tmx {
BCMPC_TM_DEFINE_VARIANT(25, false)
}  // namespace tmx
}  // namespace bcmpc
