// k_jacobi.cu -- instantiates the jacobi stencil (float + double) of the tile-streaming engine.
#include "b200_launch.cuh"
#include "b200_ops2d.cuh"

namespace b200 {
B200_DEFINE_OP(jacobi, JacobiOp)
B200_DEFINE_OP(jacobi2, Jacobi2Op)       // two sweeps per pass (temporal blocking)
}  // namespace b200
