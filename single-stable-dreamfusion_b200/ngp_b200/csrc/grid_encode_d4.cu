// Instantiates the grid-encoding kernels for input dimension D = 4 (see grid_encode.cuh).
#include "grid_encode.cuh"
namespace ngp {
namespace grid {
NGP_GRID_INSTANTIATE_DIM(4)
}  // namespace grid
}  // namespace ngp
