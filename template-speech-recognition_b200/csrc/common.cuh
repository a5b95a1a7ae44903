// common.cuh -- shared device helpers for the sm_100a kernels: error plumbing, mbarrier /
// bulk-copy (TMA) / tcgen05 inline-PTX wrappers, bit -> bf16 expansion.
//
// Everything here targets sm_100a only (compile with -gencode arch=compute_100a,code=sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tsr_b200.h"

namespace tsr {

// ----------------------------------------------------------------------------- host errors
void set_error(const char* fmt, ...);
int  check_launch(const char* what);          // cudaGetLastError -> TSR_ERR_CUDA
int  sm_count();                              // multiProcessorCount of the current device
int  require_sm100(const char* what);         // TSR_ERR_UNSUPPORTED unless cc 10.x

#define TSR_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) {                                           \
            tsr::set_error(__VA_ARGS__);                         \
            return TSR_ERR_INVALID;                              \
        }                                                        \
    } while (0)

#define TSR_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            tsr::set_error("%s failed: %s", #call, cudaGetErrorString(e__));             \
            return TSR_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

// ----------------------------------------------------------------------------- device PTX
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Spin bound for every barrier wait: a lost arrival traps (CUDA error) instead of hanging
// the box.  ~2^26 polls with back-off is seconds of wall time, far above any legitimate wait.
#ifndef TSR_SPIN_LIMIT
#define TSR_SPIN_LIMIT (1u << 26)
#endif

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// try_wait with a suspend-time hint: the hardware parks the thread (no issue slots consumed) until
// the phase completes or the hint (ns) expires -- a plain poll loop in a high-priority warp starves
// the other warps of its SMSP (measured: 10x slower epilogues next to a spinning MMA warp).
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(100000u)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > TSR_SPIN_LIMIT) {
            printf("tsr_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}

// Same wait for warps that are NOT on the critical issue path (producers, epilogue): back off
// between polls so the spinning warp does not take issue slots from the MMA-issuing warp (the
// SMSP arbiter favours higher warp ids; a tight poll loop in a higher-id warp starves it).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(40);
        if (++spins > TSR_SPIN_LIMIT) {
            printf("tsr_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// 1-D bulk copy global -> shared (TMA engine, UBLKCP in SASS); completes on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- TMEM allocation (one full warp executes these) -------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
    static_assert(kCols == 32 || kCols == 64 || kCols == 128 || kCols == 256 || kCols == 512, "TMEM cols");
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// ---- CTA pair (cta_group::2) helpers: two CTAs of a cluster on one TPC share one MMA ------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// arrive on an mbarrier of any CTA of the cluster.  Default semantics (as CUTLASS ClusterBarrier::arrive):
// a .release.cluster arrive costs a cluster-wide memory barrier (~1000 cycles measured), which serialised
// the per-tile relay; the data it guards is already in the peer's shared memory (TMA complete_tx /
// fence.proxy.async + local arrive) when the relaying thread observes the local barrier.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait with cluster-scope acquire: pairs with remote arrivals
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0, ok = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(100000u)
            : "memory");
        if (!ok && ++spins > TSR_SPIN_LIMIT) {
            printf("tsr_b200: cluster mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    } while (!ok);
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result) {   // one warp of EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// commit of the leader's cta_group::2 MMAs, arriving on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- UMMA descriptors ------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleave") canonical layout:
// core matrix = 8 rows x 16 bytes with rows 16 B apart; SBO = byte distance between 8-row
// groups (M/N direction); LBO = byte distance between the two 8-element K groups of one
// K=16 MMA.  Bit layout: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// layout_type=0 [61,64).
__device__ __forceinline__ uint64_t umma_desc_kmajor_noswz(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// 16-bit operand formats of the weight planes (and of the expanded {0,1} bits): both run at the
// same kind::f16 tensor rate.  FP16 hi+lo carries 22 mantissa bits (fp32-exact for |w| < 2^10),
// BF16 hi+lo carries 16 and is the fallback when the weights exceed the fp16 range.
#define TSR_FMT_BF16 0u
#define TSR_FMT_FP16 1u
#define TSR_FP16_MAX_WEIGHT 1024.0f
__host__ __device__ constexpr uint32_t fmt_one(uint32_t fmt) { return fmt == TSR_FMT_FP16 ? 0x3C00u : 0x3F80u; }
// Instruction descriptor for kind::f16, A/B both `fmt` (K-major both), FP32 accumulate.
__host__ __device__ constexpr uint32_t umma_idesc_f32acc(uint32_t fmt, int M, int N) {
    return (1u << 4)                                        // c_format  = F32
           | ((fmt == TSR_FMT_FP16 ? 0u : 1u) << 7)         // a_format  (0 = F16, 1 = BF16)
           | ((fmt == TSR_FMT_FP16 ? 0u : 1u) << 10)        // b_format
           | ((uint32_t)(N >> 3) << 17)                     // n_dim
           | ((uint32_t)(M >> 4) << 24);                    // m_dim
}
// D[tmem] (+)= A[smem] * B[smem]; one thread issues for the CTA.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// Warp-collective variants: every lane of a converged warp executes the call, one elected lane
// issues.  Keeping the issuing loop warp-uniform avoids the divergent single-thread region (and the
// per-instruction ELECT / BRA.U.ANY loops ptxas wraps around tcgen05 ops in divergent code).
__device__ __forceinline__ void umma_f16_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
        : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// TMEM -> registers: 32 lanes x 16 consecutive fp32 columns; lane i of the warp reads TMEM
// lane (taddr.lane + i).  A warp may only touch the lane quadrant 32*(warp_id % 4).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- bit -> 16-bit float expansion ---------------------------------------------------------
// 8 bits -> 8 values in {0.0, 1.0} of a 16-bit float format whose 1.0 is `one` (0x3F80 bf16,
// 0x3C00 fp16); bit i -> element i.
__device__ __forceinline__ uint4 expand_byte(uint32_t b, uint32_t one) {
    uint4 r;
    // two bits per 32-bit word: bit0 -> low half, bit1 -> high half; (x & 0x00010001) * one has no carries
    r.x = (((b) | (b << 15)) & 0x00010001u) * one;
    b >>= 2;
    r.y = (((b) | (b << 15)) & 0x00010001u) * one;
    b >>= 2;
    r.z = (((b) | (b << 15)) & 0x00010001u) * one;
    b >>= 2;
    r.w = (((b) | (b << 15)) & 0x00010001u) * one;
    return r;
}

// hi + lo split of a weight into two 16-bit floats of format `fmt` (returned as raw bits)
__device__ __forceinline__ void split_hi_lo(double w, uint32_t fmt, uint16_t& hi, uint16_t& lo) {
    if (fmt == TSR_FMT_FP16) {
        __half h = __double2half(w);
        __half l = __double2half(w - (double)__half2float(h));
        hi = __half_as_ushort(h);
        lo = __half_as_ushort(l);
    } else {
        __nv_bfloat16 h = __double2bfloat16(w);
        __nv_bfloat16 l = __double2bfloat16(w - (double)__bfloat162float(h));
        hi = __bfloat16_as_ushort(h);
        lo = __bfloat16_as_ushort(l);
    }
}
// max |w| tracker: non-negative floats order like their bit patterns; NaN/Inf map above every finite value
__device__ __forceinline__ void atomic_max_abs(uint32_t* slot, float w) {
    atomicMax(slot, __float_as_uint(fabsf(w)));
}
__device__ __forceinline__ uint32_t pick_format(uint32_t maxabs_bits, int requested) {
    if (requested == 1) return TSR_FMT_BF16;
    if (requested == 2) return TSR_FMT_FP16;
    return maxabs_bits <= __float_as_uint(TSR_FP16_MAX_WEIGHT) ? TSR_FMT_FP16 : TSR_FMT_BF16;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

#endif  // __CUDACC__

}  // namespace tsr
