// Block adjoint gridding kernels for blocks of 4 x 2 x 1 grid points (kbblocks.cuh).
#include "kbblocks.cuh"
namespace ib200 {
IB200_BLOCKS_INSTANTIATE(2, 1);
}
