// epi_probe.cu -- cost of K2's arg-max reduction over TMEM columns in isolation (no MMAs, no other warps).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/epi_probe tools/epi_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../template-speech-recognition_b200/csrc/common.cuh"
using namespace tsr;
namespace tsr { void set_error(const char*, ...) {} }

template <int MODE>
__global__ void __launch_bounds__(256, 1) epi_kernel(int reps, int nwarps, long long* out, int* sink) {
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc<512>(&tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;
    if (warp < nwarps) {
        const uint32_t taddr = tmem_base + (warp >= 4 ? 128 : 0) + ((uint32_t)((warp & 3) * 32) << 16);
        long long t_ld = 0, t_red = 0;
        int acc = 0;
        for (int r = 0; r < reps; ++r) {
            float best = -INFINITY; int bestk = 0;
            float bv[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            int bk[4] = {0, 1, 2, 3};
            for (int nb = 0; nb < 128; nb += 64) {
                uint32_t v[64];
                const long long t0 = clock64();
#pragma unroll
                for (int j = 0; j < 4; ++j) tmem_ld16(taddr + nb + 16 * j, *reinterpret_cast<uint32_t(*)[16]>(&v[16 * j]));
                tmem_ld_wait();
                const long long t1 = clock64();
                if (MODE == 0) {
#pragma unroll
                    for (int i = 0; i < 64; ++i) {
                        const float sc = __uint_as_float(v[i]);
                        if (sc > bv[i & 3]) { bv[i & 3] = sc; bk[i & 3] = nb + i; }
                    }
                } else if (MODE == 1) {
                    float val[32]; int idx[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float a = __uint_as_float(v[2 * i]), b = __uint_as_float(v[2 * i + 1]);
                        const bool p = b > a;
                        val[i] = p ? b : a; idx[i] = p ? 2 * i + 1 : 2 * i;
                    }
#pragma unroll
                    for (int w = 16; w >= 1; w >>= 1)
#pragma unroll
                        for (int i = 0; i < w; ++i) {
                            const bool p = val[2 * i + 1] > val[2 * i];
                            val[i] = p ? val[2 * i + 1] : val[2 * i];
                            idx[i] = p ? idx[2 * i + 1] : idx[2 * i];
                        }
                    if (val[0] > best) { best = val[0]; bestk = nb + idx[0]; }
                } else {   // MODE 2: max only (FMNMX), then find the index by equality in a second pass
                    float m = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 64; ++i) m = fmaxf(m, __uint_as_float(v[i]));
                    if (m > best) {
                        best = m;
                        int k = 63;
#pragma unroll
                        for (int i = 62; i >= 0; --i) if (__uint_as_float(v[i]) == m) k = i;
                        bestk = nb + k;
                    }
                }
                const long long t2 = clock64();
                t_ld += t1 - t0; t_red += t2 - t1;
            }
            if (MODE == 0) {
                best = bv[0]; bestk = bk[0];
#pragma unroll
                for (int c = 1; c < 4; ++c) if (bv[c] > best || (bv[c] == best && bk[c] < bestk)) { best = bv[c]; bestk = bk[c]; }
            }
            acc += bestk;
        }
        sink[threadIdx.x] = acc;
        if (lane == 0) { out[warp * 2] = t_ld / (reps * 2); out[warp * 2 + 1] = t_red / (reps * 2); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

int main() {
    long long* d_out; int* d_sink;
    cudaMalloc(&d_out, 64 * sizeof(long long)); cudaMalloc(&d_sink, 256 * sizeof(int));
    const char* names[3] = {"4 running chains", "tournament", "fmax + equality scan"};
    for (int mode = 0; mode < 3; ++mode)
        for (int nw : {4, 8}) {
            if (mode == 0) epi_kernel<0><<<1, 256>>>(200, nw, d_out, d_sink);
            else if (mode == 1) epi_kernel<1><<<1, 256>>>(200, nw, d_out, d_sink);
            else epi_kernel<2><<<1, 256>>>(200, nw, d_out, d_sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[16]; cudaMemcpy(h, d_out, 16 * sizeof(long long), cudaMemcpyDeviceToHost);
            printf("%-22s %d warps: per 64-column block: load+wait %lld cycles, reduce %lld cycles\n", names[mode], nw, h[0], h[1]);
        }
    return 0;
}
