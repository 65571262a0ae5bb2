// k_gameoflife.cu -- instantiates the gameoflife stencil (float + double) of the tile-streaming engine.
#include "b200_launch.cuh"
#include "b200_ops2d.cuh"

namespace b200 {
B200_DEFINE_OP(gameoflife, GameoflifeOp)
B200_DEFINE_OP(gameoflife2, Gameoflife2Op)       // two sweeps per pass (temporal blocking)
}  // namespace b200
