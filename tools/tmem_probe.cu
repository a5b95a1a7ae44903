// tmem_probe.cu -- latency / throughput of tcgen05.ld shapes (not part of the product).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../template-speech-recognition_b200/csrc/common.cuh"
using namespace tsr;
namespace tsr { void set_error(const char*, ...) {} }

#define LD(NUM, ...) asm volatile("tcgen05.ld.sync.aligned.32x32b.x" #NUM ".b32 {" __VA_ARGS__ "}, [%" #NUM "];"

template <int X>
__device__ __forceinline__ uint32_t ld_sum(uint32_t taddr);

template <>
__device__ __forceinline__ uint32_t ld_sum<16>(uint32_t taddr) {
    uint32_t v[16];
    tmem_ld16(taddr, v);
    tmem_ld_wait();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    return s;
}
template <>
__device__ __forceinline__ uint32_t ld_sum<32>(uint32_t taddr) {
    uint32_t v[32];
    tmem_ld32(taddr, v);
    tmem_ld_wait();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += v[i];
    return s;
}
template <>
__device__ __forceinline__ uint32_t ld_sum<64>(uint32_t taddr) {
    uint32_t v[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
        "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
          "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
          "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]),
          "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]),
          "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr) : "memory");
    tmem_ld_wait();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 64; ++i) s += v[i];
    return s;
}

template <int X>
__global__ void __launch_bounds__(256, 1) probe(int nwarps, int reps, int back_to_back, long long* out, uint32_t* sink) {
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(&tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t s = 0;
    long long t0 = 0, t1 = 0;
    if (warp < nwarps) {
        for (int r = 0; r < reps + 1; ++r) {
            if (r == 1) t0 = clock64();
            if (back_to_back) {   // 4 loads in flight, one wait
                if (X == 16) {
                    uint32_t a[16], b[16], c[16], d[16];
                    tmem_ld16(base, a); tmem_ld16(base + 16, b); tmem_ld16(base + 32, c); tmem_ld16(base + 48, d);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) s += a[i] + b[i] + c[i] + d[i];
                }
            } else {
                s += ld_sum<X>(base + (r & 3) * X);
            }
        }
        t1 = clock64();
    }
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = s;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem_ptr); }
}

int main() {
    long long* d_out; uint32_t* d_sink;
    cudaMalloc(&d_out, 8); cudaMalloc(&d_sink, 1024 * 4);
    const int reps = 200;
    for (int nw : {1, 4, 8}) {
        long long h;
        probe<16><<<1, 256>>>(nw, reps, 0, d_out, d_sink); cudaDeviceSynchronize(); cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
        printf("warps=%d  x16 ld+wait       : %6.1f cycles each\n", nw, (double)h / reps);
        probe<32><<<1, 256>>>(nw, reps, 0, d_out, d_sink); cudaDeviceSynchronize(); cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
        printf("warps=%d  x32 ld+wait       : %6.1f cycles each\n", nw, (double)h / reps);
        probe<64><<<1, 256>>>(nw, reps, 0, d_out, d_sink); cudaDeviceSynchronize(); cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
        printf("warps=%d  x64 ld+wait       : %6.1f cycles each\n", nw, (double)h / reps);
        probe<16><<<1, 256>>>(nw, reps, 1, d_out, d_sink); cudaDeviceSynchronize(); cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
        printf("warps=%d  4 x (x16) one wait: %6.1f cycles per group\n", nw, (double)h / reps);
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
