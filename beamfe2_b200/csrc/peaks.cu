// peaks.cu -- measurement utility (NOT part of the product path): FP64 FMA peak of the device, the
// denominator of the FP64 roofline fraction bench.py reports (MEASURED_PEAKS.json has no FP64 entry).
#include <cuda_runtime.h>
#include <stdint.h>

namespace {
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
}  // namespace

extern "C" {
// Launches `blocks` x 256 threads, each doing iters*128 dependent-chain-of-8 DFMAs.
// Returns the number of FMAs issued (flops = 2x) or a negative CUDA error.
long long bfe2_peaks_dfma(double* d_out, int blocks, int iters, void* stream) {
    k_dfma_peak<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_out, iters, 1.0);
    if (cudaGetLastError() != cudaSuccess) return -1;
    return (long long)blocks * 256 * (long long)iters * 128;
}
}
