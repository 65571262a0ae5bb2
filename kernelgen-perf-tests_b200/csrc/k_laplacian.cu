// k_laplacian.cu -- instantiates the laplacian stencil (float + double) of the tile-streaming engine.
#include "b200_launch.cuh"
#include "b200_ops3d.cuh"

namespace b200 {
B200_DEFINE_OP(laplacian, LaplacianOp)
}  // namespace b200
