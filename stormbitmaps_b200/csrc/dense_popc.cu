// dense_popc.cu -- CUDA-core tile kernel, direct form: LOP3 + POPC per 32-bit word.
// Kernel body and documentation: dense_tile.cuh.  Replaces storm.c:1165-1169 /
// 1199-1238 + libalgebra.h:2872-2890.
#include "dense_tile.cuh"

namespace storm {

TileShape popc_tile_shape() { return {tile::TM, tile::TN}; }
int launch_dense_popc(const DenseJob& job, cudaStream_t stream) { return tile::launch_dense_tile<tile::DIRECT>(job, stream); }

}  // namespace storm
