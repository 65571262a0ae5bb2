// k_uxx1.cu -- instantiates the uxx1 stencil (float + double) of the tile-streaming engine.
#include "b200_launch.cuh"
#include "b200_ops3d.cuh"

namespace b200 {
B200_DEFINE_OP(uxx1, Uxx1Op)
}  // namespace b200
