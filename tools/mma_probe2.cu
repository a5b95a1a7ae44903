// mma_probe2.cu -- how many tcgen05.mma per cycle one SM sustains for the code_parts tile shape (M=128, N=112) as a
// function of kind (f16 K=16 / i8 K=32), issuing warps and accumulators in flight.  Every warp issues chains of
// `chain` dependent MMAs (same accumulator) and commits; `nacc` accumulators per warp are used round-robin.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_probe2 tools/mma_probe2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../template-speech-recognition_b200/csrc/common.cuh"

using namespace tsr;
namespace tsr { void set_error(const char*, ...) {} }

struct Case { const char* name; int kind, N, nwarps, nacc, chain, wait_each; };

__global__ void __launch_bounds__(128, 1) probe_kernel(Case c, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = c.kind ? 0x01010101u : 0x3C003C00u;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_ptr);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;
    long long t0 = 0, t1 = 0;
    if (warp < c.nwarps) {
        const uint32_t abase = smem_u32(smem);
        const uint32_t bbase = smem_u32(smem + 64 * 1024);
        const uint32_t idesc = c.kind == 0 ? umma_idesc_f32acc(TSR_FMT_FP16, 128, c.N)
                                           : ((2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24));
        uint64_t ad4[4], bd4[4];
        for (int ks = 0; ks < 4; ++ks) {
            ad4[ks] = umma_desc_kmajor_noswz(abase + warp * 8192 + (ks * 2 + ks * 40) * 16, 16, 128);     // K2-style: rows 16 B apart, LBO 16 B
            bd4[ks] = umma_desc_kmajor_noswz(bbase + ks * 2 * c.N * 16, c.N * 16, 128);
        }
        uint32_t phase = 0;
        const int per_warp_acc = c.nacc;                     // accumulators of this warp: columns (warp * nacc + a) * 128
        for (int r = 0; r < reps + 1; ++r) {
            if (r == 1) t0 = clock64();
            for (int a = 0; a < per_warp_acc; ++a) {
                const uint32_t dt = tmem_base + (uint32_t)(((warp * per_warp_acc + a) & 3) * (c.N > 128 ? 256 : 128));
                for (int i = 0; i < c.chain; i += 4) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        if (c.kind == 0) umma_f16_elect(dt, ad4[ks], bd4[ks], idesc, (i + ks) > 0);
                        else umma_i8_elect(dt, ad4[ks], bd4[ks], idesc, (i + ks) > 0);
                    }
                }
                if (c.wait_each) { umma_commit_elect(&bar[warp]); mbar_wait(&bar[warp], phase); phase ^= 1; }
            }
            if (!c.wait_each) { umma_commit_elect(&bar[warp]); mbar_wait(&bar[warp], phase); phase ^= 1; }
        }
        t1 = clock64();
        if ((threadIdx.x & 31) == 0) out[blockIdx.x * 4 + warp] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

int main() {
    Case cases[] = {
        // kind, N, issuing warps, accumulators per warp, dependent chain length, wait for completion after every chain
        {"f16 N=112  1 warp  x 1 acc, chain 28, wait each", 0, 112, 1, 1, 28, 1},
        {"i8  N=112  1 warp  x 1 acc, chain 28, wait each", 1, 112, 1, 1, 28, 1},
        {"f16 N=112  1 warp  x 1 acc, chain 12, wait each", 0, 112, 1, 1, 12, 1},
        {"i8  N=112  1 warp  x 1 acc, chain 12, wait each", 1, 112, 1, 1, 12, 1},
        {"f16 N=112  1 warp  x 4 acc, chain 28           ", 0, 112, 1, 4, 28, 0},
        {"i8  N=112  1 warp  x 4 acc, chain 12           ", 1, 112, 1, 4, 12, 0},
        {"f16 N=112  2 warps x 2 acc, chain 28           ", 0, 112, 2, 2, 28, 0},
        {"i8  N=112  2 warps x 2 acc, chain 12           ", 1, 112, 2, 2, 12, 0},
        {"f16 N=112  4 warps x 1 acc, chain 28           ", 0, 112, 4, 1, 28, 0},
        {"i8  N=112  4 warps x 1 acc, chain 12           ", 1, 112, 4, 1, 12, 0},
        {"i8  N=112  4 warps x 1 acc, chain 12, wait each", 1, 112, 4, 1, 12, 1},
        {"i8  N=224  2 warps x 1 acc, chain 8            ", 1, 224, 2, 1, 8, 0},
        {"i8  N=256  2 warps x 1 acc, chain 12           ", 1, 256, 2, 1, 12, 0},
        {"i8  N=64   4 warps x 1 acc, chain 12           ", 1, 64, 4, 1, 12, 0},
        {"f16 N=64   4 warps x 1 acc, chain 28           ", 0, 64, 4, 1, 28, 0},
    };
    long long* d_out;
    cudaMalloc(&d_out, 148 * 4 * sizeof(long long));
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int reps = 100;
    for (int grid : {1, 148}) {
        printf("grid = %d CTAs (one per SM)\n", grid);
        for (auto& c : cases) {
            cudaMemset(d_out, 0, 148 * 4 * sizeof(long long));
            probe_kernel<<<grid, 128, 200 * 1024>>>(c, reps, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
            long long h[148 * 4];
            cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
            double mx = 0; for (int i = 0; i < grid * 4; ++i) mx = h[i] > mx ? h[i] : mx;
            const double nm = (double)reps * c.chain * c.nacc * c.nwarps;
            const double ideal = c.N / 2.0;
            printf("  %s : %7.1f cycles per MMA on the SM (tensor time %5.1f: %3.0f%%), %7.1f per MMA of one chain\n", c.name, mx / nm, ideal,
                   100.0 * ideal / (mx / nm), mx / ((double)reps * c.chain * c.nacc));
        }
    }
    return 0;
}
