// k_wave13pt.cu -- instantiates the wave13pt stencil (float + double) of the tile-streaming engine.
#include "b200_launch.cuh"
#include "b200_ops3d.cuh"

namespace b200 {
B200_DEFINE_OP_TILED(wave13pt, Wave13ptOp, 12)
}  // namespace b200
