// k_lapgsrb.cu -- instantiates the lapgsrb stencil (float + double) of the tile-streaming engine.
#include "b200_launch.cuh"
#include "b200_ops3d.cuh"

namespace b200 {
B200_DEFINE_OP_TILED(lapgsrb, LapgsrbOp, 12)
}  // namespace b200
