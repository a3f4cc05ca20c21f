// Cluster-split tcgen05 rollout kernel with the proposal network (policy-guided shooting on nets of more than two hidden layers), production flavour
#include "rollout_deep.cuh"

namespace bcmpc {
namespace dpx {
BCMPC_DEEP_DEFINE_POL(deep_launch_pol_prod, false, false)
}  // namespace dpx
}  // namespace bcmpc
