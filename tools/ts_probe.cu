// ts_probe.cu -- does taking the A operand from TMEM (tcgen05.mma [d], [a_tmem], b_desc) relieve the shared-memory
// pipe of K2 (M=128, N=112, K=16, kind::f16), and how do concurrent tcgen05.ld accumulator reads interact?
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ts_probe tools/ts_probe.cu
// Warp 0 issues batches of 26 MMAs (two accumulators alternating per batch); warps 4-7 (optional) read the other
// accumulator with tcgen05.ld.x64 in a loop.  Prints cycles per MMA and reader bytes per cycle.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../template-speech-recognition_b200/csrc/common.cuh"

using namespace tsr;
namespace tsr { void set_error(const char*, ...) {} }

__device__ __forceinline__ void umma_f16_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}

struct Case { const char* name; int N; int ts; int readers; int cols; int width; int batches; int issuers; };

__global__ void __launch_bounds__(256, 1) probe_kernel(Case c, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar2[2];
    __shared__ uint32_t tmem_ptr;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
    if (threadIdx.x == 0) { mbar_init(&bar2[0], 1); mbar_init(&bar2[1], 1); mbar_fence_init(); stop = 0; }
    if (warp == 0) tmem_alloc<512>(&tmem_ptr);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;
    // TMEM columns: [0,128) acc0, [128,256) acc1, [256,368) A (104 columns used)
    if (warp < c.issuers) {
        uint64_t& bar = bar2[warp];
        const uint32_t abase = smem_u32(smem), bbase = smem_u32(smem + 64 * 1024);
        const uint32_t idesc = umma_idesc_f32acc(TSR_FMT_FP16, 128, c.N);
        uint32_t phase = 0;
        long long t0 = 0;
        for (int r = 0; r < reps + 1; ++r) {
            if (r == 1) t0 = clock64();
            const uint32_t dt = tmem_base + (c.issuers > 1 ? warp : (r & 1)) * 128;
            for (int bb = 0; bb < c.batches; ++bb)
#pragma unroll
            for (int s = 0; s < 26; ++s) {
                const uint64_t bd = umma_desc_kmajor_noswz(bbase + (s % 13) * 2 * c.N * 16, c.N * 16, 128);
                if (c.ts) umma_f16_ts_elect(dt, tmem_base + 256 + (s % 13) * 8, bd, idesc, s > 0);
                else umma_f16_elect(dt, umma_desc_kmajor_noswz(abase + (s % 13) * 4096, 2048, 128), bd, idesc, s > 0);
            }
            umma_commit_elect(&bar);
            mbar_wait(&bar, phase);
            phase ^= 1;
        }
        const long long t1 = clock64();
        if (lane == 0 && warp == 0) { out[blockIdx.x * 2] = t1 - t0; stop = 1; }
    } else if (warp >= 4 && c.readers) {
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        long long n = 0;
        const long long t0 = clock64();
        uint32_t sink = 0;
        while (!stop) {
            if (c.width == 16) {
                for (int col = 0; col < c.cols; col += 64) {
                    uint32_t v[64];
#pragma unroll
                    for (int j = 0; j < 4; ++j) tmem_ld16(taddr + col + 16 * j, *reinterpret_cast<uint32_t(*)[16]>(&v[16 * j]));
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 64; ++i) sink ^= v[i];
                }
            } else if (c.width == 32) {
                for (int col = 0; col < c.cols; col += 64) {
                    uint32_t v[64];
                    tmem_ld32(taddr + col, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                    tmem_ld32(taddr + col + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 64; ++i) sink ^= v[i];
                }
            } else if (c.width == 64) {
                for (int col = 0; col < c.cols; col += 64) {
                    uint32_t v[64];
                    tmem_ld64(taddr + col, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 64; ++i) sink ^= v[i];
                }
            } else {
                for (int col = 0; col < c.cols; col += 128) {
                    uint32_t v[128];
                    tmem_ld128(taddr + col, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 128; ++i) sink ^= v[i];
                }
            }
            n += c.cols * 128;    // bytes per warp pass: 32 lanes x cols x 4
        }
        const long long t1 = clock64();
        if (sink == 0x12345678u) out[500] = sink;
        if (lane == 0 && warp == 4) out[blockIdx.x * 2 + 1] = n * 4 * 1000 / (t1 - t0);   // CTA bytes per 1000 cycles (4 reader warps)
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

int main() {
    Case cases[] = {
        {"SS N=112 2 issuers                 ", 112, 0, 0, 128, 32, 16, 2},
        {"TS N=112 2 issuers                 ", 112, 1, 0, 128, 32, 16, 2},
        {"SS N=112 2 issuers + readers x16   ", 112, 0, 1, 128, 16, 16, 2},
        {"TS N=112 2 issuers + readers x16   ", 112, 1, 1, 128, 16, 16, 2},
        {"TS N=112 1 issuer                  ", 112, 1, 0, 128, 32, 16, 1},
    };
    long long* d_out;
    cudaMalloc(&d_out, 1024 * sizeof(long long));
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const int reps = 60;
    for (int grid : {1, 148}) {
        printf("grid = %d\n", grid);
        for (auto& c : cases) {
            cudaMemset(d_out, 0, 1024 * sizeof(long long));
            probe_kernel<<<grid, 256, 160 * 1024>>>(c, reps, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
            long long h[296];
            cudaMemcpy(h, d_out, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
            double mx = 0, rd = 0;
            for (int i = 0; i < grid; ++i) { mx = h[2 * i] > mx ? h[2 * i] : mx; rd += h[2 * i + 1]; }
            printf("  %s : %6.1f cycles/MMA (floor %4.1f)   readers %.1f B/clk per CTA\n", c.name, mx / (reps * 26.0 * c.batches * c.issuers),
                   128.0 * c.N / 256.0, rd / grid / 1000.0);
        }
    }
    return 0;
}
