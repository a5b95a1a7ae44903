// FP64 issue-rate probe: cycles per DFMA warp instruction per SM, for 1..32 resident warps with 8 independent chains each.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp64_probe tools/fp64_probe.cu && gpurun_out/fp64_probe
#include <cstdio>
#include <cuda_runtime.h>
template <typename T>
__global__ void probe(T* out, long long* cyc, int iters) {
    T a[8];
    for (int i = 0; i < 8; ++i) a[i] = (T)(threadIdx.x + i) * (T)1e-3;
    const T b = (T)1.0000001, c = (T)1e-9;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = a[i] * b + c;
    }
    const long long t1 = clock64();
    T s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    for (int warps = 1; warps <= 32; warps *= 2) {
        long long h[148];
        probe<double><<<148, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        const double dp = (double)h[0] / (iters * 8.0 * warps);
        probe<float><<<148, warps * 32>>>((float*)out, cyc, iters); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        const double sp = (double)h[0] / (iters * 8.0 * warps);
        printf("warps/SM %2d: DFMA %.2f cycles per warp instruction per SM, FFMA %.3f\n", warps, dp, sp);
    }
    return 0;
}
