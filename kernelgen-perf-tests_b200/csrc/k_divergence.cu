// k_divergence.cu -- instantiates the divergence stencil (float + double) of the tile-streaming engine.
#include "b200_launch.cuh"
#include "b200_ops3d.cuh"

namespace b200 {
B200_DEFINE_OP_TILED(divergence, DivergenceOp, 12)
}  // namespace b200
