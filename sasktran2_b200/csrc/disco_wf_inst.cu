// One translation unit per stream count: nvcc -DDISCO_N=<N>
#include "disco_wf.cuh"

#define DISCO_CAT2(a, b) a##b
#define DISCO_CAT(a, b) DISCO_CAT2(a, b)

namespace disco {
void DISCO_CAT(launch_wf_layer_n, DISCO_N)(const ChunkView& V, cudaStream_t s) { launch_wf_layer_n<DISCO_N>(V, s); }
}  // namespace disco
