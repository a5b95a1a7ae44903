// mma_probe.cu -- micro-benchmark of tcgen05.mma issue/execute cost versus operand layout.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_probe tools/mma_probe.cu
// Run:    gpurun_out/mma_probe            (prints cycles per MMA for each case)
// Not part of the product; used to choose tile shapes / layouts.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../template-speech-recognition_b200/csrc/common.cuh"

using namespace tsr;
namespace tsr { void set_error(const char*, ...) {} }

struct Case {
    const char* name;
    int kind;        // 0 = f16 (K=16), 1 = i8 (K=32)
    int N;
    int a_mode;      // 0 = canonical no-swizzle (LBO 2048, SBO 128), 1 = K2-style (LBO 16, start + 16*k), 2 = SW128
    int b_mode;      // 0 = canonical no-swizzle (LBO N*16, SBO 128), 2 = SW128
    int nmma;        // MMAs per batch (distinct K slices cycled over 4)
    int nacc;        // 1: all MMAs accumulate into one TMEM tile; 2: alternate between two tiles
};

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    // K-major SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (1)
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void umma_i8_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}

__global__ void __launch_bounds__(128, 1) probe_kernel(Case c, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5;
    // fill smem with small finite values
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = c.kind ? 0x01010101u : 0x3C003C00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_ptr);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;
    if (warp == 0) {
        const uint32_t abase = smem_u32(smem);
        const uint32_t bbase = smem_u32(smem + 64 * 1024);
        uint32_t idesc;
        if (c.kind == 0) idesc = umma_idesc_f32acc(TSR_FMT_FP16, 128, c.N);
        else idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint64_t ad4[4], bd4[4];
        for (int ks = 0; ks < 4; ++ks) {
            if (c.a_mode == 0) ad4[ks] = umma_desc_kmajor_noswz(abase + ks * 4096, 2048, 128);
            else if (c.a_mode == 1) ad4[ks] = umma_desc_kmajor_noswz(abase + (ks * 2 + ks * 40) * 16, 16, 128);
            else ad4[ks] = desc_sw128(abase + ks * 32);
            if (c.b_mode == 0) bd4[ks] = umma_desc_kmajor_noswz(bbase + ks * 2 * c.N * 16, c.N * 16, 128);
            else bd4[ks] = desc_sw128(bbase + ks * 32);
        }
        uint32_t phase = 0;
        long long t0 = 0, t1 = 0;
        for (int r = 0; r < reps + 1; ++r) {
            if (r == 1) t0 = clock64();
            for (int i = 0; i < c.nmma; i += 4) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t dt = tmem_base + (c.nacc > 1 ? (ks & 1) * 256 : 0);
                    if (c.kind == 0) umma_f16_elect(dt, ad4[ks], bd4[ks], idesc, (i + ks) > 1);
                    else umma_i8_elect(dt, ad4[ks], bd4[ks], idesc, (i + ks) > 1);
                }
            }
            umma_commit_elect(&bar);
            mbar_wait(&bar, phase);
            phase ^= 1;
        }
        t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

int main() {
    Case cases[] = {
        {"f16 N=112 A canonical  B canonical", 0, 112, 0, 0, 24, 1},
        {"f16 N=112 A K2-style   B canonical", 0, 112, 1, 0, 24, 1},
        {"f16 N=128 A canonical  B canonical", 0, 128, 0, 0, 24, 1},
        {"f16 N=160 A canonical  B canonical", 0, 160, 0, 0, 24, 1},
        {"f16 N=240 A canonical  B canonical", 0, 240, 0, 0, 24, 1},
        {"f16 N=256 A canonical  B canonical", 0, 256, 0, 0, 24, 1},
        {"f16 N=112 A sw128      B sw128    ", 0, 112, 2, 2, 24, 1},
        {"f16 N=160 A sw128      B sw128    ", 0, 160, 2, 2, 24, 1},
        {"f16 N=256 A sw128      B sw128    ", 0, 256, 2, 2, 24, 1},
        {"f16 N=160 A canonical  B sw128    ", 0, 160, 0, 2, 24, 1},
        {"f16 N=160 A sw128      B canonical", 0, 160, 2, 0, 24, 1},
        {"i8  N=160 A canonical  B canonical", 1, 160, 0, 0, 24, 1},
        {"i8  N=240 A canonical  B canonical", 1, 240, 0, 0, 24, 1},
        {"i8  N=240 A sw128      B sw128    ", 1, 240, 2, 2, 24, 1},
        {"i8  N=256 A sw128      B sw128    ", 1, 256, 2, 2, 24, 1},
        {"f16 N=112 canonical, 2 accumulators", 0, 112, 0, 0, 24, 2},
        {"f16 N=160 canonical, 2 accumulators", 0, 160, 0, 0, 24, 2},
        {"f16 N=256 canonical, 2 accumulators", 0, 256, 0, 0, 24, 2},
        {"i8  N=240 canonical, 2 accumulators", 1, 240, 0, 0, 24, 2},
    };
    long long* d_out;
    cudaMalloc(&d_out, 148 * sizeof(long long));
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int reps = 200;
    for (int grid : {1, 148}) {
        printf("grid = %d CTAs (one per SM)\n", grid);
        for (auto& c : cases) {
            probe_kernel<<<grid, 128, 200 * 1024>>>(c, reps, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
            long long h[148];
            cudaMemcpy(h, d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
            double mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            double per = mx / (double)(reps * c.nmma);
            double ideal = 128.0 * c.N / 256.0;
            printf("  %s : %7.1f cycles/MMA (ideal %5.1f, %.0f%%)\n", c.name, per, ideal, 100.0 * ideal / per);
        }
    }
    return 0;
}
