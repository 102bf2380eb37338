// One translation unit per stream count: nvcc -DDISCO_N=<N>
#include "disco_bvp.cuh"

#define DISCO_CAT2(a, b) a##b
#define DISCO_CAT(a, b) DISCO_CAT2(a, b)

namespace disco {
void DISCO_CAT(launch_bvp_n, DISCO_N)(const ChunkView& V, cudaStream_t s) { launch_bvp_n<DISCO_N>(V, s); }
void DISCO_CAT(launch_bvp_adjoint_n, DISCO_N)(const ChunkView& V, cudaStream_t s) { launch_bvp_adjoint_n<DISCO_N>(V, s); }
void DISCO_CAT(launch_bvp_multi_n, DISCO_N)(const ChunkView& V, cudaStream_t s) { launch_bvp_multi_n<DISCO_N>(V, s); }
}  // namespace disco
