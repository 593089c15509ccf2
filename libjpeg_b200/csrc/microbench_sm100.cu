// microbench_sm100.cu -- measures the int32 issue rate of this GPU so that the reconstruction kernel's
// "fraction of int32 roofline" has a MEASURED denominator (MEASURED_PEAKS.json only carries HBM and bf16).
// Three loops of independent register chains: IMAD only (fma pipe), IADD3/LOP3 only (alu pipe), and a 1:1 mix.
// Reported in Gop/s with SURVEY.md 8d's counting: an integer multiply-add counts 2, everything else 1.
#include <cuda_runtime.h>

#include <cstdint>

#include "b200jpg.h"

namespace {

template <int MODE>
__global__ void __launch_bounds__(256) int_rate_kernel(int *out, int iters, int a, int b) {
    int x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (MODE == 0) {  // 8 IMAD
                x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
                x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
            } else if (MODE == 1) {  // 8 ALU (LOP3 / IADD3 alternating so nothing folds)
                x0 = (x0 ^ a) + b; x1 = (x1 ^ a) + b; x2 = (x2 ^ a) + b; x3 = (x3 ^ a) + b;
                x4 = (x4 ^ a) + b; x5 = (x5 ^ a) + b; x6 = (x6 ^ a) + b; x7 = (x7 ^ a) + b;
            } else {  // 4 IMAD + 4 LOP3 + 4 IADD3
                x0 = x0 * a + b; x1 = (x1 ^ a) + b; x2 = x2 * a + b; x3 = (x3 ^ a) + b;
                x4 = x4 * a + b; x5 = (x5 ^ a) + b; x6 = x6 * a + b; x7 = (x7 ^ a) + b;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <int MODE>
float run(int *buf, int grid, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int_rate_kernel<MODE><<<grid, 256>>>(buf, iters / 8, 3, 7);  // warm-up
    cudaEventRecord(e0);
    int_rate_kernel<MODE><<<grid, 256>>>(buf, iters, 3, 7);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms;
}

}  // namespace

extern "C" B200JPG_API int b200jpg_microbench_int32(int device, float *imad_gops, float *alu_gops, float *mix_gops) {
    if (cudaSetDevice(device < 0 ? 0 : device) != cudaSuccess) return B200JPG_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device < 0 ? 0 : device) != cudaSuccess) return B200JPG_ERR_NO_DEVICE;
    const int grid = prop.multiProcessorCount * 8, iters = 4096;
    int *buf = nullptr;
    if (cudaMalloc(&buf, sizeof(int) * (size_t)grid * 256) != cudaSuccess) return B200JPG_ERR_OUT_OF_MEMORY;
    const double threads = (double)grid * 256, per_iter = 16.0 * 8.0;
    float t0 = run<0>(buf, grid, iters), t1 = run<1>(buf, grid, iters), t2 = run<2>(buf, grid, iters);
    cudaFree(buf);
    if (cudaGetLastError() != cudaSuccess) return B200JPG_ERR_CUDA;
    // ops per thread-iteration: mode 0: 128 IMAD = 256 ops; mode 1: 128 (LOP3 + IADD3) = 256 ops;
    // mode 2: 64 IMAD (128 ops) + 64 (LOP3 + IADD3) (128 ops)
    if (imad_gops) *imad_gops = (float)(threads * iters * per_iter * 2.0 / (t0 * 1e-3) / 1e9);
    if (alu_gops) *alu_gops = (float)(threads * iters * per_iter * 2.0 / (t1 * 1e-3) / 1e9);
    if (mix_gops) *mix_gops = (float)(threads * iters * per_iter * 2.0 / (t2 * 1e-3) / 1e9);
    return B200JPG_OK;
}
