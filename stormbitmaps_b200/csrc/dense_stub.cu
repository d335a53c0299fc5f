// TEMPORARY: placeholder until dense_csa.cu lands.
#include "common.cuh"
namespace storm {
TileShape csa_tile_shape() { return {128, 128}; }
int launch_dense_csa(const DenseJob&, cudaStream_t) { set_error("CSA kernel not built"); return STORM_B200_EINVAL; }
}
