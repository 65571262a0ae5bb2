// k_gaussblur.cu -- instantiates the gaussblur stencil (float + double) of the tile-streaming engine.
#include "b200_launch.cuh"
#include "b200_ops2d.cuh"

namespace b200 {
B200_DEFINE_OP(gaussblur, GaussblurOp)
B200_DEFINE_OP(gaussblur2, Gaussblur2Op)       // two sweeps per pass (temporal blocking)
}  // namespace b200
