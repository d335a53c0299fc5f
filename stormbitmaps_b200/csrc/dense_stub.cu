// TEMPORARY: placeholders until dense_csa.cu / dense_umma.cu land.
#include "common.cuh"
namespace storm {
TileShape csa_tile_shape() { return {128, 128}; }
int launch_dense_csa(const DenseJob&, cudaStream_t) { set_error("CSA kernel not built"); return STORM_B200_EINVAL; }
TileShape umma_tile_shape() { return {256, 256}; }
int launch_dense_umma(const DenseJob&, cudaStream_t) { set_error("UMMA kernel not built"); return STORM_B200_EINVAL; }
bool umma_supports(const DenseJob&) { return false; }
}
