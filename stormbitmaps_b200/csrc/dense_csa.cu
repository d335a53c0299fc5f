// dense_csa.cu -- CUDA-core tile kernel, carry-save form: a 7:3 compressor in registers
// feeds POPC (what the reference's Harley-Seal loop does in zmm registers,
// libalgebra.h:2684-2744).  Kernel body and documentation: dense_tile.cuh.
#include "dense_tile.cuh"

namespace storm {

TileShape csa_tile_shape() { return {tile::TM, tile::TN}; }
int launch_dense_csa(const DenseJob& job, cudaStream_t stream) { return tile::launch_dense_tile<tile::CSA>(job, stream); }

}  // namespace storm
