// Instantiates the grid-encoding kernels for input dimension D = 5 (see grid_encode.cuh).
#include "grid_encode.cuh"
namespace ngp {
namespace grid {
NGP_GRID_INSTANTIATE_DIM(5)
}  // namespace grid
}  // namespace ngp
