// Block adjoint gridding kernels for blocks of 4 x 1 x 1 grid points (kbblocks.cuh).
#include "kbblocks.cuh"
namespace ib200 {
IB200_BLOCKS_INSTANTIATE(1, 1);
}
