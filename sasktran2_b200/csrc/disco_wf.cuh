// Weighting-function kernels (thin wrappers around disco_wf_body.h), instantiated per stream count in
// disco_wf_inst.cu.
#pragma once
#include "disco_kernels.cuh"
#include "disco_wf_body.h"

namespace disco {

template <int N, int G>
__global__ void __launch_bounds__(64) k_wf_layer(ChunkView V) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (V.wf_bottom_only) {   // only the layer on the ground (kernel-based BRDF next to the register-resident kernel)
        if (idx >= (long long)V.nw * V.M) return;
        idx = idx * V.T.L + (V.T.L - 1);
    } else if (idx >= (long long)V.nw * V.M * V.T.L) {
        return;
    }
    wf_layer_body<N, G>(V, idx);
}

template <int N>
static void launch_wf_layer_n(const ChunkView& V, cudaStream_t s) {
    const long long n = (long long)V.nw * V.M * (V.wf_bottom_only ? 1 : V.T.L);
    const unsigned blocks = (unsigned)((n + 63) / 64);
    switch (V.ngroups) {
        case 0: k_wf_layer<N, 0><<<blocks, 64, 0, s>>>(V); break;
        case 1: k_wf_layer<N, 1><<<blocks, 64, 0, s>>>(V); break;
        case 2: k_wf_layer<N, 2><<<blocks, 64, 0, s>>>(V); break;
        default: break;
    }
}

}  // namespace disco
