// em.cu -- K4: Bernoulli-mixture EM on bit-packed data
// (bernoulli_mixture.py:263-291 E-step, :177-193 M-step, :108-160 loop; bernoulli_em.py:22-49, 103-127).
//
// Both contractions of an EM iteration are the same product  C[m][r] = sum_k Wq[m][k] * bit[r][k]  of an int8
// matrix with a bit matrix, run by ONE tcgen05 kernel (bitgemm_kernel):
//   E-step:  r = data row n, k = feature d, Wq = the 3 balanced base-256 digit planes of
//            round((logT - logInvT) / scale[comp])                           (rows m = plane * KT + comp)
//   M-step:  r = feature d, k = data row n (bits transposed once per data set), Wq = the 3 digit planes of
//            round(aff * 8355711)
// The bit rows are the UMMA *N* operand (N = 256 rows per MMA, expanded to u8 {0,1} in shared memory), the int8
// planes the *M* operand (128 rows = up to 42 components x 3 planes; more components tile over M), so one MMA
// carries 128 cycles of tensor work whatever the number of components.  int32 accumulation is exact, which makes
// every sum order independent: the kernel is stream-K (each CTA owns an equal share of the (tile, chunk) space and
// adds its partial tiles to global int32 accumulators with coalesced reductions) and a rank's statistics are
// deterministic.  Posteriors are quantised per component with a power-of-two scale s_k >= max_n aff[n][k], so a
// component that holds 1e-30 of the mass keeps 24 significant bits (the reference forms its template from the
// RATIO of such sums, bernoulli_mixture.py:179-181; a global scale would silently kill it).
//
// Per iteration: bitgemm (E) -> em_posterior_kernel (soft-max, log-likelihood, column maxima) ->
// em_aff_planes_kernel (digit planes, responsibility sums) -> bitgemm (M) -> em_pack_kernel (float64 statistics)
// -> [allreduce by the caller] -> em_control_kernel (stopping rule on the device) -> em_params_kernel (divide,
// clip, logs, weight planes of the next E-step).
#include <math.h>
#include <string.h>

#include "common.cuh"

extern "C" int tsr_rowwords(int D);

namespace tsr {

#define EM_QMAX 8355711              // 127 * (1 + 256 + 65536): largest value of three balanced base-256 digits
#define EM_KT_MAX 42                 // mixture components per 128-row M tile (3 planes each)
#define EM_LL_SCALE 1048576.0        // 2^20: fixed-point unit in which a rank sums its rows' log-likelihoods (order independent)
#define EM_C_SCALE 68719476736.0      // 2^36: fixed-point unit of c[k] = sum_d logInvT[k][d] (integer sums are order independent)
#define EM_MAX_K 4095

#define BG_THREADS 448               // 14 warps: loader, 8 expanders, 4 epilogue, MMA issuer
#define BG_NSTRIP 4                  // expanded bit strips in flight (one per expander group)
#define BG_NW 4                      // ring of int8 weight tiles
#define BG_STRIP_BYTES 32768         // 256 rows x 128 k-positions of u8
#define BG_NRAW 4                    // ring of raw bit tiles (256 rows x 16 B) landed by TMA tensor loads
#define BG_RAW_BYTES 4096
#define BG_TBUF_WORDS (32 * 33)      // per-epilogue-warp transpose buffer

struct EmShape {
    int K, D, KT, Mtiles, mrows, mrp;   // mrows = 3*KT useful rows of an M tile, mrp = mrows rounded up to 16
    long long N, Npad;                  // Npad = N rounded up to 256
    int rwX, nchunkE, Dpad;             // words per data row, 128-feature chunks per row, D rounded up to 256
    long long nchunkM;                  // 128-row chunks of the M-step contraction
    size_t tile_bytes;                  // one int8 tile: 8 k-groups x mrp rows x 16 B
    size_t off_wblob, off_c, off_scale, off_accE, off_accM, off_ablob, off_small, off_scratch, off_aff, total;
};

static EmShape em_shape(long long N, int D, int K) {
    EmShape s;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    s.K = K; s.D = D; s.N = N;
    s.KT = K < EM_KT_MAX ? K : EM_KT_MAX;
    s.Mtiles = (K + s.KT - 1) / s.KT;
    s.mrows = 3 * s.KT;
    s.mrp = (s.mrows + 15) & ~15;     // also the UMMA N of the TS kernel (M = 128 needs N % 16 == 0)
    s.Npad = (N + 255) / 256 * 256;
    s.rwX = tsr_rowwords(D);
    s.nchunkE = (s.rwX + 3) / 4;
    s.Dpad = (D + 255) / 256 * 256;
    s.nchunkM = s.Npad / 128;
    s.tile_bytes = (size_t)8 * s.mrp * 16;
    s.off_wblob = 0;
    s.off_c = up((size_t)s.Mtiles * s.nchunkE * s.tile_bytes);
    s.off_scale = s.off_c + up((size_t)K * 8);
    s.off_accE = s.off_scale + up((size_t)K * 4);
    s.off_accM = s.off_accE + up((size_t)s.Mtiles * s.mrp * s.Npad * 4);
    s.off_ablob = s.off_accM + up((size_t)s.Mtiles * s.mrp * s.Dpad * 4);
    s.off_small = s.off_ablob + up((size_t)s.Mtiles * s.nchunkM * s.tile_bytes);
    s.off_scratch = s.off_small + up(sizeof(long long) * (32 * (size_t)K + 32 + (size_t)K) + sizeof(uint32_t) * 32 * (size_t)K + 64);
    s.off_aff = s.off_scratch + up((size_t)K * s.Npad * 4);
    s.total = s.off_aff + up((size_t)K * (size_t)N * 4);
    return s;
}

// Device-side state of the EM loop: the stopping rule is evaluated on the device so that the host can queue
// several iterations without reading the log-likelihood back (kernels of iterations past the stop exit at once).
struct EmCtl {
    int done, iterations, ntrace, mode;      // mode 0: relative change (bernoulli_mixture.py:144), 1: absolute gain (bernoulli_em.py:29)
    int min_iter, max_iter, cap, pad;
    double tol, ll_old;
    // double trace[cap] follows
};

// ------------------------------------------------------------------------------------------------
// bit x int8 GEMM, stream-K.
//   bits  [R][rw] uint32, row r, k-position order = bit order
//   wblob [Mtiles][nchunk][8 k-groups][mrp rows][16 B]: byte b of k-group g of chunk c holds k-position
//         128c + 32(g/2) + 8(b%4) + 4(g%2) + b/4  -- the order in which the expanders lay a 32-bit word out
//         (word >> s) & 0x01010101 gives the four bytes of shift s with two instructions
//   acc   [Mtiles][mrp][Rpad] int32 += products (Rpad = R rounded up to 256)
struct BgArgs {
    const uint32_t* tiled;          // the bit matrix in tile order [128-row tile][chunk][128 rows][16 B] (tsr_em_tile_bits), or null
    const uint32_t* bits;
    const uint8_t* wblob;
    int* acc;
    long long R, Rpad;
    int rw, nchunk, Mtiles, mrp, mrows, Rtiles;
    int nw;                         // TS kernel: weight tiles in flight
    const int* done;
};

// A CTA's share of the flattened (R tile, M tile, chunk) space, cut into segments at tile borders.
struct BgWalk {
    long long pos, end, tile;
    int c, nchunk, Mtiles;
    __device__ BgWalk(const BgArgs& A) {
        const long long total = (long long)A.Rtiles * A.Mtiles * A.nchunk;
        pos = total * blockIdx.x / gridDim.x;
        end = total * (blockIdx.x + 1) / gridDim.x;
        nchunk = A.nchunk; Mtiles = A.Mtiles;
        tile = pos / nchunk;
        c = (int)(pos - tile * nchunk);
    }
    // next segment: chunks [c0, c0 + n) of tile (rt, mt)
    __device__ bool next_seg(int& rt, int& mt, int& c0, int& n) {
        if (pos >= end) return false;
        rt = (int)(tile / Mtiles); mt = (int)(tile - (long long)rt * Mtiles);
        c0 = c;
        long long len = nchunk - c;
        if (len > end - pos) len = end - pos;
        n = (int)len;
        pos += len; c += n;
        if (c == nchunk) { c = 0; ++tile; }
        return true;
    }
};

__device__ __forceinline__ void umma_i8_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// RAW_TMA: the bit rows reach shared memory as bulk (TMA) copies of the pre-tiled matrix (tsr_em_tile_bits: the
// 256 rows x 16 bytes a chunk needs are two contiguous 2 KB blocks) and the expanders read them with conflict-free
// LDS.128.  Fetching 16 bytes per row straight from the row-major matrix -- per-thread LDG, or a 2-D tensor-map TMA
// box -- touches one 32-byte DRAM sector per row at a stride of the row pitch and ran at 1.2 TB/s (18 % of the HBM
// peak) whatever the rest of the kernel did; the per-thread loads also put as many wavefronts on the L1 data pipe
// as the expanded stores.  The LDG path remains for callers without a tiled copy.
template <bool RAW_TMA>
__global__ void __launch_bounds__(BG_THREADS, 1) bitgemm_kernel(const __grid_constant__ BgArgs A) {
    if (A.done && *A.done) return;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t tile_bytes = 8u * (uint32_t)A.mrp * 16u;
    uint8_t* wring = smem;                                            // BG_NW weight tiles; the M = 128 descriptors read past a
    uint8_t* strips = smem + BG_NW * tile_bytes;                      // short tile into what follows (harmless: rows >= mrows unused)
    uint8_t* raw = strips + BG_NSTRIP * BG_STRIP_BYTES;               // RAW_TMA: BG_NRAW raw bit tiles
    int* tbuf = reinterpret_cast<int*>(raw + (RAW_TMA ? BG_NRAW * BG_RAW_BYTES : 0));
    uint64_t* wfull = reinterpret_cast<uint64_t*>(tbuf + 4 * BG_TBUF_WORDS);
    uint64_t* wempty = wfull + BG_NW;
    uint64_t* afull = wempty + BG_NW;
    uint64_t* aempty = afull + BG_NSTRIP;
    uint64_t* tfull = aempty + BG_NSTRIP;
    uint64_t* tempty = tfull + 2;
    uint64_t* rfull = tempty + 2;                                     // [BG_NRAW]
    uint64_t* rempty = rfull + BG_NRAW;                               // [BG_NRAW]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(rempty + BG_NRAW);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < BG_NW; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
        // (one arrival per expander WARP: 256 per-thread arrivals on one barrier word serialise in the shared-memory
        // atomic unit -- with the work removed the pipeline skeleton alone took 85 % of the kernel time)
        for (int s = 0; s < BG_NSTRIP; ++s) { mbar_init(&afull[s], 8); mbar_init(&aempty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 128); }
        for (int s = 0; s < BG_NRAW; ++s) { mbar_init(&rfull[s], 1); mbar_init(&rempty[s], 8); }
        mbar_fence_init();
    }
    if (warp == 13) tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ======================= weight-tile (and raw bit tile) loader =======================
        if (lane == 0) {
            BgWalk w(A);
            int rt, mt, c0, n, stage = 0;
            uint32_t ph = 0, ic = 0;
            while (w.next_seg(rt, mt, c0, n)) {
                const uint8_t* src = A.wblob + ((size_t)mt * A.nchunk + c0) * tile_bytes;
                for (int i = 0; i < n; ++i, ++ic) {
                    if (RAW_TMA) {
                        const uint32_t rs = ic % BG_NRAW;
                        mbar_wait_relaxed(&rempty[rs], ((ic / BG_NRAW) & 1) ^ 1);
                        mbar_arrive_expect_tx(&rfull[rs], BG_RAW_BYTES);
                        const uint32_t* t0 = A.tiled + (((size_t)(2 * rt) * A.nchunk + (c0 + i)) << 9);       // 512 words per (tile, chunk)
                        bulk_g2s(raw + rs * BG_RAW_BYTES, t0, 2048, &rfull[rs]);
                        bulk_g2s(raw + rs * BG_RAW_BYTES + 2048, t0 + ((size_t)A.nchunk << 9), 2048, &rfull[rs]);
                    }
                    mbar_wait_relaxed(&wempty[stage], ph ^ 1);
                    mbar_arrive_expect_tx(&wfull[stage], tile_bytes);
                    bulk_g2s(wring + stage * tile_bytes, src + (size_t)i * tile_bytes, tile_bytes, &wfull[stage]);
                    if (++stage == BG_NW) { stage = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp >= 1 && warp <= 8) {
        // ================================== bit expanders ==================================
        // All 256 threads work on every strip, one row each (68 instructions), into strip buffer ic % 4.  A buffer's
        // cycle is MMA (512 cycles) -> wake-up -> expand -> proxy fence -> arrive -> MMA wake-up; with a quarter of the
        // threads per strip (one group of 64 per buffer, four rows per thread) the expansion alone took ~500 cycles and
        // the whole cycle ~4000, twice what four buffers can hide; one row per thread keeps it near 1000.
        const int tg = threadIdx.x - 32;                     // row of the strip
        BgWalk w(A);
        int rt, mt, c0, n;
        uint32_t ic = 0;
        auto expand = [&](uint8_t* strip, const uint4& x) {
            uint8_t* dst = strip + tg * 16;
            const uint32_t wv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int wi = 0; wi < 4; ++wi) {
                const uint32_t v = wv[wi];
                const uint4 a = make_uint4(v & 0x01010101u, (v >> 1) & 0x01010101u, (v >> 2) & 0x01010101u, (v >> 3) & 0x01010101u);
                const uint4 b = make_uint4((v >> 4) & 0x01010101u, (v >> 5) & 0x01010101u, (v >> 6) & 0x01010101u, (v >> 7) & 0x01010101u);
                *reinterpret_cast<uint4*>(dst + (2 * wi) * 4096) = a;
                *reinterpret_cast<uint4*>(dst + (2 * wi + 1) * 4096) = b;
            }
        };
        auto load = [&](int l_rt, int l_c) -> uint4 {        // LDG path: the row's four words of chunk l_c
            const long long row = (long long)l_rt * 256 + tg;
            uint2 lo = make_uint2(0u, 0u), hi = make_uint2(0u, 0u);
            if (row < A.R) {
                const uint2* p = reinterpret_cast<const uint2*>(A.bits + row * A.rw + 4 * l_c);
                lo = __ldg(p);
                if (4 * l_c + 2 < A.rw) hi = __ldg(p + 1);
            }
            return make_uint4(lo.x, lo.y, hi.x, hi.y);
        };
        if (RAW_TMA) {
            while (w.next_seg(rt, mt, c0, n)) {
                for (int i = 0; i < n; ++i, ++ic) {
                    const uint32_t buf = ic & 3, rs = ic % BG_NRAW;
                    mbar_wait_relaxed(&rfull[rs], (ic / BG_NRAW) & 1);
                    const uint4 x = *reinterpret_cast<const uint4*>(raw + rs * BG_RAW_BYTES + tg * 16);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&rempty[rs]);  // the warp's words are in registers: the tile may be refilled
                    mbar_wait_relaxed(&aempty[buf], ((ic >> 2) & 1) ^ 1);
                    expand(strips + buf * BG_STRIP_BYTES, x);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&afull[buf]);
                }
            }
        } else {
            // two chunks of look-ahead in registers
            int prt[3], pc[3];
            uint4 x[3];
            int have = 0;
            bool more = true;
            int i = 0;
            n = 0;
            auto fetch = [&]() {                             // queue the next chunk's words, if any
                if (i == n) {
                    if (!w.next_seg(rt, mt, c0, n)) { more = false; return; }
                    i = 0;
                }
                prt[have] = rt; pc[have] = c0 + i; x[have] = load(rt, c0 + i);
                ++i; ++have;
            };
            fetch(); if (more) fetch();
            while (have > 0) {
                if (more) fetch();
                const uint32_t buf = ic & 3;
                mbar_wait_relaxed(&aempty[buf], ((ic >> 2) & 1) ^ 1);
                expand(strips + buf * BG_STRIP_BYTES, x[0]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&afull[buf]);
                ++ic;
                x[0] = x[1]; x[1] = x[2];
                --have;
            }
        }
    } else if (warp == 13) {
        // ==================================== MMA issuer ====================================
        // M = 128 (int8 planes, signed) x N = 256 (bit rows, unsigned) x K = 32, S32 accumulate
        const uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t desc_hi = (uint32_t)(128 >> 4) | (1u << 14);                                 // SBO = 128 B, version 1
        const uint64_t a_hi = ((uint64_t)desc_hi << 32) | ((uint64_t)((uint32_t)(A.mrp * 16) >> 4) << 16);   // LBO = k-group stride of a weight tile
        const uint64_t b_hi = ((uint64_t)desc_hi << 32) | ((uint64_t)(4096 >> 4) << 16);                     // LBO = k-group stride of a strip
        const uint32_t kstep_a = (uint32_t)(2 * A.mrp * 16) >> 4, kstep_b = (uint32_t)(2 * 4096) >> 4;
        BgWalk w(A);
        int rt, mt, c0, n, stage = 0;
        uint32_t ph = 0, ic = 0, seg = 0;
        while (w.next_seg(rt, mt, c0, n)) {
            const uint32_t accb = seg & 1;
            mbar_wait(&tempty[accb], ((seg >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + accb * 256;
            for (int i = 0; i < n; ++i, ++ic) {
                const uint32_t buf = ic & 3;
                mbar_wait(&afull[buf], (ic >> 2) & 1);
                mbar_wait(&wfull[stage], ph);
                tc_fence_after();
                const uint32_t abase = smem_u32(wring + stage * tile_bytes) >> 4;
                const uint32_t bbase = smem_u32(strips + buf * BG_STRIP_BYTES) >> 4;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    umma_i8_ss(d_tmem, a_hi | (uint64_t)(abase + j * kstep_a), b_hi | (uint64_t)(bbase + j * kstep_b), idesc,
                               (i | j) != 0 ? 1u : 0u);
                umma_commit_elect(&aempty[buf]);
                umma_commit_elect(&wempty[stage]);
                if (++stage == BG_NW) { stage = 0; ph ^= 1; }
            }
            umma_commit_elect(&tfull[accb]);
            ++seg;
        }
    } else if (warp >= 9 && warp <= 12) {
        // ===================================== epilogue =====================================
        // TMEM lane = row m of the tile, column = bit row r.  A warp transposes 32 x 32 blocks through its private
        // shared-memory buffer so that one reduction instruction adds 32 consecutive r of one m (128 contiguous bytes).
        const int quad = warp & 3;
        int* tb = tbuf + (warp - 9) * BG_TBUF_WORDS;
        int live = A.mrows - quad * 32;                     // rows of this quadrant that carry data
        if (live > 32) live = 32;
        BgWalk w(A);
        int rt, mt, c0, n;
        uint32_t seg = 0;
        while (w.next_seg(rt, mt, c0, n)) {
            const uint32_t accb = seg & 1;
            mbar_wait_relaxed(&tfull[accb], (seg >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + accb * 256 + ((uint32_t)(quad * 32) << 16);
            int* out = A.acc + ((size_t)mt * A.mrp + quad * 32) * A.Rpad + (size_t)rt * 256 + lane;
            for (int cb = 0; cb < 8; ++cb) {
                uint32_t v[32];
                if (live > 0) {
                    tmem_ld32(taddr + cb * 32, v);
                    tmem_ld_wait();
                }
                if (cb == 7) {                              // accumulator read: hand it back to the MMA warp
                    tc_fence_before();
                    mbar_arrive(&tempty[accb]);
                }
                if (live > 0) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) tb[i * 33 + lane] = (int)v[i];
                    __syncwarp();
                    for (int mi = 0; mi < live; ++mi) {
                        const int val = tb[lane * 33 + mi];
                        if (val != 0) atomicAdd(out + (size_t)mi * A.Rpad + cb * 32, val);
                    }
                    __syncwarp();
                }
            }
            ++seg;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 13) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------
// bitgemm_ts_kernel: the same product with the operand roles swapped once more and the bit rows as the A operand
// *in tensor memory*:   C[m][r] += sum_k bit[r][k] * Wq[m][k]   as   D[128 bit rows][N = mrp planes] += A(TMEM) * B(smem).
// The SS kernel above is bound by the L1/shared-memory data pipe: every expanded byte is written to shared memory by
// the expanders and read back by the tensor core (ncu: lsu 42 % + tensor-operand 39 % of the pipe's peak, tensor pipe
// 47 %).  Here an expander thread owns TMEM lane = bit row, expands its words in registers and stores them with
// tcgen05.st (TMEM write port, 256 B/clk, no shared memory); the tensor core reads A from TMEM and only the small
// int8 weight tile (N rows x 32 B per MMA) from shared memory.  One MMA is M = 128 x N = mrp x K = 32, i.e. short
// (N / 2 cycles), so with mrp <= 64 two issuer warps take alternate chunks into their own accumulators.
// Used when all components fit one M tile (K <= 42) and the rows can be fetched by TMA; otherwise the SS kernel runs.
//   TMEM columns: [0, 2 * NIS * mrp) accumulators (segment parity, issuer); the rest a ring of A chunks x 32 columns.
#define TS_THREADS 768               // warp 0 raw-tile loader, 1-16 expanders, 17-20 epilogue, 21-22 MMA issuers, 23 weight-tile loader
#define TS_NA_MAX 16                 // A chunks in flight in TMEM: every 32-column slot the accumulators leave free (10 at K = 16)
#define TS_NRAW 16                   // raw bit tiles in flight: a chunk lasts ~200 cycles, a TMA round trip ~2000-3000
#define TS_RAW_BYTES 2048            // 128 rows x 16 B
#define TS_NW_MAX 16                 // weight tiles in flight (as many as fit 144 KB)

struct TsWalk {     // as BgWalk with 128-row tiles and one M tile
    long long pos, end, tile;
    int c, nchunk;
    __device__ TsWalk(int Rtiles, int nchunk_) {
        const long long total = (long long)Rtiles * nchunk_;
        pos = total * blockIdx.x / gridDim.x;
        end = total * (blockIdx.x + 1) / gridDim.x;
        nchunk = nchunk_;
        tile = pos / nchunk;
        c = (int)(pos - tile * nchunk);
    }
    __device__ bool next_seg(int& rt, int& c0, int& n) {
        if (pos >= end) return false;
        rt = (int)tile; c0 = c;
        long long len = nchunk - c;
        if (len > end - pos) len = end - pos;
        n = (int)len;
        pos += len; c += n;
        if (c == nchunk) { c = 0; ++tile; }
        return true;
    }
};

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <int NIS>
__global__ void __launch_bounds__(TS_THREADS, 1) bitgemm_ts_kernel(const __grid_constant__ BgArgs A) {
    if (A.done && *A.done) return;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t tile_bytes = 8u * (uint32_t)A.mrp * 16u;
    const uint32_t TS_NW = (uint32_t)A.nw;
    uint8_t* wring = smem;
    uint8_t* raw = smem + TS_NW * tile_bytes;
    uint64_t* wfull = reinterpret_cast<uint64_t*>(raw + TS_NRAW * TS_RAW_BYTES);
    uint64_t* wempty = wfull + TS_NW_MAX;
    uint64_t* rfull = wempty + TS_NW_MAX;
    uint64_t* rempty = rfull + TS_NRAW;
    uint64_t* afull = rempty + TS_NRAW;          // [TS_NA_MAX]
    uint64_t* aempty = afull + TS_NA_MAX;        // [TS_NA_MAX]
    uint64_t* tfull = aempty + TS_NA_MAX;        // [2]
    uint64_t* tempty = tfull + 2;                // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int N = A.mrp;
    const int Rtiles = (int)(A.Rpad / 128);

    if (threadIdx.x == 0) {
        for (int s = 0; s < TS_NW_MAX; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
        for (int s = 0; s < TS_NRAW; ++s) { mbar_init(&rfull[s], 1); mbar_init(&rempty[s], 4); }      // one arrival per warp of a set
        for (int s = 0; s < TS_NA_MAX; ++s) { mbar_init(&afull[s], 4); mbar_init(&aempty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], NIS); mbar_init(&tempty[s], 128); }
        mbar_fence_init();
    }
    if (warp == 21) tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t a_col0 = (uint32_t)(2 * NIS * N + 31) & ~31u;      // first column of the A ring
    const uint32_t TS_NA = (512u - a_col0) / 32u < TS_NA_MAX ? (512u - a_col0) / 32u : TS_NA_MAX;

    if (warp == 0) {
        // raw bit tiles: one 2 KB bulk copy (128 rows x 16 B, contiguous in the tiled matrix) per chunk
        if (lane == 0) {
            TsWalk w(Rtiles, A.nchunk);
            int rt, c0, n;
            uint32_t ic = 0;
            while (w.next_seg(rt, c0, n)) {
                for (int i = 0; i < n; ++i, ++ic) {
                    const uint32_t rs = ic % TS_NRAW;
                    mbar_wait_relaxed(&rempty[rs], ((ic / TS_NRAW) & 1) ^ 1);
                    mbar_arrive_expect_tx(&rfull[rs], TS_RAW_BYTES);
                    bulk_g2s(raw + rs * TS_RAW_BYTES, A.tiled + (((size_t)rt * A.nchunk + (c0 + i)) << 9), TS_RAW_BYTES, &rfull[rs]);
                }
            }
        }
    } else if (warp == 23) {
        // weight tiles: one bulk copy per chunk
        if (lane == 0) {
            TsWalk w(Rtiles, A.nchunk);
            int rt, c0, n;
            uint32_t ic = 0;
            while (w.next_seg(rt, c0, n)) {
                const uint8_t* src = A.wblob + (size_t)c0 * tile_bytes;
                for (int i = 0; i < n; ++i, ++ic) {
                    const uint32_t stage = ic % TS_NW;
                    mbar_wait_relaxed(&wempty[stage], ((ic / TS_NW) & 1) ^ 1);
                    mbar_arrive_expect_tx(&wfull[stage], tile_bytes);
                    bulk_g2s(wring + stage * tile_bytes, src + (size_t)i * tile_bytes, tile_bytes, &wfull[stage]);
                }
            }
        }
    } else if (warp >= 1 && warp <= 16) {
        // ================================== bit expanders ==================================
        // TMEM lane = bit row: a warp may only touch lanes 32 * (warp % 4) .. +31.  Four sets of four warps (one per lane
        // quadrant); set p expands the chunks ic % 4 == p, a thread its row's four words (128 bytes = 32 TMEM columns).
        // One chunk costs a warp ~100 instructions plus ~400 cycles of latencies it cannot overlap (barrier waits, the
        // shared-memory load, tcgen05.wait::st), so the sets work on four chunks concurrently.
        const int quad = warp & 3, set = (warp - 1) >> 2;
        const int row = quad * 32 + lane;
        TsWalk w(Rtiles, A.nchunk);
        int rt, c0, n;
        uint32_t ic = 0;
        while (w.next_seg(rt, c0, n)) {
            for (int i = 0; i < n; ++i, ++ic) {
                if ((int)(ic & 3) != set) continue;
                const uint32_t rs = ic % TS_NRAW, slot = ic % TS_NA;
                mbar_wait_relaxed(&rfull[rs], (ic / TS_NRAW) & 1);
                const uint4 x = *reinterpret_cast<const uint4*>(raw + rs * TS_RAW_BYTES + row * 16);
                __syncwarp();
                if (lane == 0) mbar_arrive(&rempty[rs]);
                mbar_wait_relaxed(&aempty[slot], ((ic / TS_NA) & 1) ^ 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + a_col0 + slot * 32 + ((uint32_t)(quad * 32) << 16);
                const uint32_t wv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int wi = 0; wi < 4; ++wi) {
                    uint32_t v[8];
#pragma unroll
                    for (int s8 = 0; s8 < 8; ++s8) v[s8] = (wv[wi] >> s8) & 0x01010101u;     // K index 4 s + jj  <->  bit 8 jj + s of the word
                    tmem_st8(taddr + 8 * wi, v);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&afull[slot]);
            }
        }
    } else if (warp == 21 || (NIS == 2 && warp == 22)) {
        // ==================================== MMA issuers ====================================
        // issuer `iss` takes the chunks ic % NIS == iss into accumulator (segment parity, iss)
        const int iss = warp - 21;
        const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // S32 += U8 * S8
        const uint32_t desc_hi = (uint32_t)(128 >> 4) | (1u << 14);                                 // SBO = 128 B, version 1
        const uint64_t b_hi = ((uint64_t)desc_hi << 32) | ((uint64_t)((uint32_t)(N * 16) >> 4) << 16);   // LBO = k-group stride of a weight tile
        const uint32_t kstep_b = (uint32_t)(2 * N * 16) >> 4;
        TsWalk w(Rtiles, A.nchunk);
        int rt, c0, n;
        uint32_t ic = 0, seg = 0;
        while (w.next_seg(rt, c0, n)) {
            const uint32_t sp = seg & 1;
            mbar_wait(&tempty[sp], ((seg >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (sp * NIS + iss) * N;
            bool first = true;
            for (int i = 0; i < n; ++i, ++ic) {
                if ((int)(ic % NIS) != iss) continue;
                const uint32_t slot = ic % TS_NA, stage = ic % TS_NW;
                mbar_wait(&afull[slot], (ic / TS_NA) & 1);
                mbar_wait(&wfull[stage], (ic / TS_NW) & 1);
                tc_fence_after();
                const uint32_t bbase = smem_u32(wring + stage * tile_bytes) >> 4;
                const uint32_t abase = tmem_base + a_col0 + slot * 32;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    umma_i8_ts(d_tmem, abase + 8 * j, b_hi | (uint64_t)(bbase + j * kstep_b), idesc, (first && j == 0) ? 0u : 1u);
                first = false;
                umma_commit_elect(&aempty[slot]);
                umma_commit_elect(&wempty[stage]);
            }
            umma_commit_elect(&tfull[sp]);        // (arrives even if this issuer had no chunk in the segment)
            ++seg;
        }
    } else if (warp >= 17 && warp <= 20) {
        // ===================================== epilogue =====================================
        // lane = bit row r, column = plane row m: for a fixed m the warp's 32 lanes are 32 consecutive r of acc[m][.],
        // one coalesced 128-byte reduction
        const int quad = warp & 3;
        TsWalk w(Rtiles, A.nchunk);
        int rt, c0, n;
        uint32_t seg = 0, ic = 0;
        while (w.next_seg(rt, c0, n)) {
            const uint32_t sp = seg & 1;
            // which issuers contributed to this segment
            bool used[2] = {false, false};
            if (NIS == 1) used[0] = true;
            else {
                used[ic % 2] = true;
                if (n > 1) used[(ic + 1) % 2] = true;
            }
            ic += n;
            mbar_wait_relaxed(&tfull[sp], (seg >> 1) & 1);
            tc_fence_after();
            int* out = A.acc + (size_t)rt * 128 + quad * 32 + lane;
            for (int m0 = 0; m0 < N; m0 += 16) {
                uint32_t v[16], u[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) { v[i] = 0u; u[i] = 0u; }
                const uint32_t lanebits = (uint32_t)(quad * 32) << 16;
                if (used[0]) tmem_ld16(tmem_base + (sp * NIS + 0) * N + m0 + lanebits, v);
                if (NIS == 2 && used[1]) tmem_ld16(tmem_base + (sp * NIS + 1) * N + m0 + lanebits, u);
                tmem_ld_wait();
                if (m0 + 16 >= N) {                         // accumulators read: hand them back to the issuers
                    tc_fence_before();
                    mbar_arrive(&tempty[sp]);
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int val = (int)v[i] + (int)u[i];
                    if (m0 + i < A.mrows && val != 0) atomicAdd(out + (size_t)(m0 + i) * A.Rpad, val);
                }
            }
            ++seg;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 21) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

static size_t bitgemm_smem(int mrp, bool raw_tma) {
    return (size_t)BG_NW * 8 * mrp * 16 + (size_t)BG_NSTRIP * BG_STRIP_BYTES + (raw_tma ? BG_NRAW * BG_RAW_BYTES : 0) +
           4 * BG_TBUF_WORDS * 4 + 512;
}

static int launch_bitgemm(const uint32_t* bits, const uint32_t* tiled, long long R, int rw, const uint8_t* wblob, int* acc,
                          long long Rpad, int nchunk, const EmShape& s, const int* done, cudaStream_t st) {
    BgArgs A;
    memset(&A, 0, sizeof(A));
    A.bits = bits; A.tiled = tiled; A.wblob = wblob; A.acc = acc; A.R = R; A.Rpad = Rpad; A.rw = rw; A.nchunk = nchunk;
    A.Mtiles = s.Mtiles; A.mrp = s.mrp; A.mrows = s.mrows; A.Rtiles = (int)(Rpad / 256); A.done = done;
#ifdef TSR_DEBUG_TRACE      // experiments only: force the per-thread LDG fetch / the SS-form kernel
    const bool raw_tma = tiled != nullptr && getenv("TSR_DEBUG_BITGEMM_LDG") == nullptr;
    const bool allow_ts = getenv("TSR_DEBUG_BITGEMM_SS") == nullptr;
#else
    const bool raw_tma = tiled != nullptr;
    const bool allow_ts = true;
#endif
    const int nsm = sm_count();
    if (raw_tma && s.Mtiles == 1 && allow_ts) {
        // all components in one M tile: bit rows as the A operand in tensor memory (bitgemm_ts_kernel)
        size_t nw = (size_t)(144 * 1024) / s.tile_bytes;
        if (nw > TS_NW_MAX) nw = TS_NW_MAX;
        A.nw = (int)nw;
        const size_t tsmem = nw * s.tile_bytes + TS_NRAW * TS_RAW_BYTES + 1280;
        const long long ttotal = (Rpad / 128) * nchunk;
        const int tgrid = (int)(ttotal < nsm ? ttotal : nsm);
        if (tgrid < 1) return TSR_OK;
        static bool ts_attr = false;
        if (!ts_attr) {
            TSR_CUDA(cudaFuncSetAttribute(bitgemm_ts_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            TSR_CUDA(cudaFuncSetAttribute(bitgemm_ts_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            ts_attr = true;
        }
        if (s.mrp <= 64) bitgemm_ts_kernel<2><<<tgrid, TS_THREADS, tsmem, st>>>(A);
        else bitgemm_ts_kernel<1><<<tgrid, TS_THREADS, tsmem, st>>>(A);
        return check_launch("bitgemm_ts_kernel");
    }
    const size_t smem = bitgemm_smem(s.mrp, raw_tma);
    static bool attr_set = false;
    if (!attr_set) {
        TSR_CUDA(cudaFuncSetAttribute(bitgemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        TSR_CUDA(cudaFuncSetAttribute(bitgemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const long long total = (long long)A.Rtiles * A.Mtiles * nchunk;
    const int grid = (int)(total < nsm ? total : nsm);
    if (grid < 1) return TSR_OK;
    if (raw_tma) bitgemm_kernel<true><<<grid, BG_THREADS, smem, st>>>(A);
    else bitgemm_kernel<false><<<grid, BG_THREADS, smem, st>>>(A);
    return check_launch("bitgemm_kernel");
}

// bits [R][rw] -> tile order [128-row tile][chunk][128 rows][4 words] (rows >= R and words >= rw read as zero): what a CTA
// needs for one chunk of one row tile is then one contiguous 2 KB block (sequential DRAM reads instead of one 32-byte
// sector per row at a stride of the row pitch).  One thread per (tile, chunk, row): 16 bytes.
__global__ void tile_bits_kernel(const uint32_t* __restrict__ bits, long long R, int rw, int nchunk, long long ntile,
                                 uint4* __restrict__ tiled) {
    const long long total = ntile * nchunk * 128;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i & 127);
        const long long tc = i >> 7;
        const int c = (int)(tc % nchunk);
        const long long row = (tc / nchunk) * 128 + r;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (row < R) {
            const uint32_t* p = bits + row * rw + 4 * c;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (4 * c + j < rw) w[j] = __ldg(p + j);
        }
        tiled[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// Small per-call scratch in the workspace.  Per-CTA partial results are combined with integer atomics; 782 CTAs
// hitting the same 17 addresses serialise in the L2 atomic unit (~12 us), so every value has EM_SLOTS replicas
// (CTA b uses slot b % EM_SLOTS) that the consumer reduces.
//   [respQ int64[SLOTS][K] | llQ int64[SLOTS] | sk double[K] | colmax uint32[SLOTS][K]]
#define EM_SLOTS 32
struct EmSmall {
    long long* respQ;
    long long* llQ;
    double* sk;                 // per-component posterior scale (power of two), published by the planes kernel
    uint32_t* colmax;
    unsigned int* ticket;       // CTAs of em_aff_planes_kernel that have finished (the last one writes the statistics tail)
};
static EmSmall em_small(uint8_t* ws, const EmShape& s) {
    EmSmall m;
    m.respQ = reinterpret_cast<long long*>(ws + s.off_small);
    m.llQ = m.respQ + (size_t)EM_SLOTS * s.K;
    m.sk = reinterpret_cast<double*>(m.llQ + EM_SLOTS);
    m.colmax = reinterpret_cast<uint32_t*>(m.sk + s.K);
    m.ticket = reinterpret_cast<unsigned int*>(m.colmax + (size_t)EM_SLOTS * s.K);
    return m;
}
// power-of-two scale of component k's posteriors: the smallest 2^e >= column max (0 for an empty column)
__device__ __forceinline__ float em_col_scale(uint32_t colmax_bits) {
    const float mx = __uint_as_float(colmax_bits);
    if (!(mx > 0.f)) return 0.f;
    int e;
    const float fr = frexpf(mx, &e);                       // mx = fr * 2^e, fr in [0.5, 1)
    return ldexpf(1.f, fr == 0.5f ? e - 1 : e);
}

// q / EM_QMAX * s_k with the division split into an integer quotient and a remainder term: exact for multiples of EM_QMAX
// (hard assignments give the bit counts themselves) without the ~20-instruction double division per element.  No FMA
// contraction: two roundings, as the tests' NumPy restatement.
__device__ __forceinline__ double em_q_to_stat(long long q, double sk) {
    const long long qi = q / (long long)EM_QMAX, qr = q - qi * (long long)EM_QMAX;
    return __dadd_rn((double)qi, __dmul_rn((double)qr, 1.0 / (double)EM_QMAX)) * sk;
}

// ------------------------------------------------------------------------------------------------
// SIMT cross-checks (impl = TSR_IMPL_SIMT): sequential fp32 adds of the set bits' weights, the reference's
// arithmetic for the scores; exact integer sums of the quantised posteriors for the numerators.
__global__ void em_scores_simt_kernel(const uint32_t* __restrict__ X, long long N, int D, int rw, int K,
                                      const float* __restrict__ logT, const float* __restrict__ logInvT,
                                      const long long* __restrict__ c, float* __restrict__ scratch, long long Npad, const int* done) {
    if (done && *done) return;
    const long long total = N * K;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / N);
        const long long n = i - (long long)k * N;
        const uint32_t* row = X + n * rw;
        const float* lt = logT + (size_t)k * D;
        const float* li = logInvT + (size_t)k * D;
        float acc = 0.f;
        for (int wi = 0; wi < rw; ++wi) {
            uint32_t x = row[wi];
            while (x) {
                const int d = wi * 32 + __ffs(x) - 1;
                x &= x - 1;
                acc = __fadd_rn(acc, __fsub_rn(lt[d], li[d]));
            }
        }
        scratch[(size_t)k * Npad + n] = (float)((double)acc + (double)c[k] / EM_C_SCALE);
    }
}

// numerators of the SIMT path: thread per (k, d), rows in order, the same quantised posteriors as the tensor path
__global__ void em_mstep_simt_kernel(const uint32_t* __restrict__ X, long long N, int D, int rw, int K,
                                     const float* __restrict__ aff, const double* __restrict__ skd, double* __restrict__ stats,
                                     const int* done) {
    if (done && *done) return;
    const long long total = (long long)K * D;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / D), d = (int)(i - (long long)k * D);
        const float sk = (float)skd[k];
        const float isk = sk >= 1.17549435e-38f ? 1.f / sk : 0.f;               // exact: s_k is a power of two
        long long acc = 0;
        for (long long n = 0; n < N; ++n)
            if ((X[n * rw + (d >> 5)] >> (d & 31)) & 1u) acc += (long long)rintf(fmaxf(aff[n * K + k], 0.f) * isk * (float)EM_QMAX);
        stats[i] = em_q_to_stat(acc, (double)sk);
    }
}

// ------------------------------------------------------------------------------------------------
// Parameters -> E-step operand.  One CTA per component k.
//   FROM_STATS: the M-step proper (bernoulli_mixture.py:177-193): weights = resp / N, T = numer / N / weights,
//   clipped to [min_prob, 1 - min_prob], logs in float32.
//   Both: W = logT - logInvT, scale[k] = max|W| / EM_QMAX, c[k] = sum_d logInvT (double, fixed order), then the three
//   balanced base-256 digits of round(W / scale) in the bitgemm tile order.
struct ParamArgs {
    const double* stats;        // FROM_STATS
    long long N_total;
    float min_prob;
    float *templates, *logT, *logInvT, *weights, *logw;   // FROM_STATS: written; else logT / logInvT are read
    int D, K, KT, mrp, nchunk;
    uint8_t* wblob;             // may be null (parameters only)
    long long* c;               // 2^-36 fixed point
    float* scale;
    const int* done;
};

template <bool FROM_STATS>
__global__ void __launch_bounds__(256) em_params_kernel(ParamArgs P) {
    if (P.done && *P.done) return;
    __shared__ float s_mx[256];
    __shared__ double s_c[256];
    __shared__ float s_scale;
    const int k = blockIdx.x, tid = threadIdx.x;
    const size_t row = (size_t)k * P.D;
    float w = 0.f;
    if (FROM_STATS) w = (float)P.stats[(size_t)P.K * P.D + k] / (float)P.N_total;     // np.mean(affinities, 0)
    float mx = 0.f;
    double csum = 0.0;
    for (int d = tid; d < P.D; d += 256) {
        float lt, li;
        if (FROM_STATS) {
            float t = (float)P.stats[row + d];                                 // np.dot(affinities.T, data) as float32
            t = t / (float)P.N_total;                                          // /= num_data
            t = t / w;                                                         // /= weights
            t = fminf(fmaxf(t, P.min_prob), 1.f - P.min_prob);                 // threshold_templates
            lt = logf(t); li = logf(1.f - t);
            P.templates[row + d] = t; P.logT[row + d] = lt; P.logInvT[row + d] = li;
        } else {
            lt = P.logT[row + d]; li = P.logInvT[row + d];
        }
        mx = fmaxf(mx, fabsf(lt - li));
        csum += (double)li;
    }
    s_mx[tid] = mx; s_c[tid] = csum;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (tid < off) { s_mx[tid] = fmaxf(s_mx[tid], s_mx[tid + off]); s_c[tid] += s_c[tid + off]; }
        __syncthreads();
    }
    if (tid == 0) {
        const float sc = s_mx[0] > 0.f ? s_mx[0] / (float)EM_QMAX : 1.f;
        s_scale = sc;
        if (P.scale) P.scale[k] = sc;
        if (P.c) P.c[k] = llrint(s_c[0] * EM_C_SCALE);
        if (FROM_STATS) { P.weights[k] = w; P.logw[k] = logf(w); }
    }
    __syncthreads();
    if (!P.wblob) return;
    const float sc = s_scale;
    const int mt = k / P.KT, kl = k - mt * P.KT;
    // one thread per 32-feature word: 2 k-groups x 3 planes x 16 bytes
    for (int it = tid; it < P.nchunk * 4; it += 256) {
        const int c = it >> 2, wi = it & 3;
        const int d0 = c * 128 + wi * 32;
        uint32_t out[3][2][4];
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int q = 0; q < 4; ++q) out[p][h][q] = 0u;
#pragma unroll
        for (int s = 0; s < 8; ++s)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int d = d0 + 8 * jj + s;
                int q = 0;
                if (d < P.D) {
                    // the stores of this block above are visible after the barrier
                    const float wv = P.logT[row + d] - P.logInvT[row + d];
                    float qf = rintf(wv / sc);
                    qf = fminf(fmaxf(qf, -(float)EM_QMAX), (float)EM_QMAX);
                    q = (int)qf;
                }
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const int digit = ((q + 128) & 255) - 128;
                    q = (q - digit) >> 8;
                    out[p][s >> 2][s & 3] |= (uint32_t)(digit & 255) << (8 * jj);
                }
            }
        uint8_t* tile = P.wblob + ((size_t)mt * P.nchunk + c) * ((size_t)8 * P.mrp * 16);
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int h = 0; h < 2; ++h)
                *reinterpret_cast<uint4*>(tile + ((size_t)(2 * wi + h) * P.mrp + p * P.KT + kl) * 16) =
                    make_uint4(out[p][h][0], out[p][h][1], out[p][h][2], out[p][h][3]);
    }
}

// The M-step proper of the iteration loop, fully parallel: one WARP per 32-feature word of one component, lane =
// feature (coalesced reads of the statistics, coalesced writes of templates / logs).  Everything for a parameter --
// divide, clip, logs, the three weight digits -- happens in one pass: no reduction is needed for the quantisation step
// because the clip bounds |W| = |logT - logInvT| by log((1 - p_min) / p_min).  A lane's digit is byte jj = lane / 8 of
// word s = lane % 8 of the tile row pair (see bitgemm_kernel); lanes 0..7 collect the four bytes of their word with
// three shuffles.  c[k] is summed in 2^-36 fixed point with integer atomics (order independent).
// grid = (ceil(nchunk * 4 / 8), K), 8 warps per CTA.
__global__ void __launch_bounds__(256) em_finalize_fast_kernel(ParamArgs P, float wbound) {
    if (P.done && *P.done) return;
    __shared__ long long s_c[8];
    const int k = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int it = blockIdx.x * 8 + wp;                    // 32-feature word of this warp
    const size_t row = (size_t)k * P.D;
    const float w = (float)P.stats[(size_t)P.K * P.D + k] / (float)P.N_total;          // np.mean(affinities, 0)
    const float sc = wbound > 0.f ? wbound / (float)EM_QMAX : 1.f;
    long long csum = 0;
    if (it < P.nchunk * 4) {
        const int c = it >> 2, wi = it & 3;
        const int d = c * 128 + wi * 32 + lane;
        int q = 0;
        if (d < P.D) {
            float t = (float)P.stats[row + d];                             // np.dot(affinities.T, data) as float32
            t = t / (float)P.N_total;                                      // /= num_data
            t = t / w;                                                     // /= weights
            t = fminf(fmaxf(t, P.min_prob), 1.f - P.min_prob);             // threshold_templates
            const float lt = logf(t), li = logf(1.f - t);
            P.templates[row + d] = t; P.logT[row + d] = lt; P.logInvT[row + d] = li;
            csum = llrint((double)li * EM_C_SCALE);
            float qf = rintf((lt - li) / sc);
            qf = fminf(fmaxf(qf, -(float)EM_QMAX), (float)EM_QMAX);
            q = (int)qf;
        }
        const int mt = k / P.KT, kl = k - mt * P.KT;
        uint8_t* tile = P.wblob + ((size_t)mt * P.nchunk + c) * ((size_t)8 * P.mrp * 16);
        const int s8 = lane & 7;                                           // word of the 32-byte row pair this lane belongs to
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const int digit = ((q + 128) & 255) - 128;
            q = (q - digit) >> 8;
            uint32_t v = (uint32_t)(digit & 255);
            v |= __shfl_down_sync(0xffffffffu, v, 8) << 8;                 // lanes l, l+8, l+16, l+24 = bytes jj = 0..3
            const uint32_t hi2 = __shfl_down_sync(0xffffffffu, v, 16);
            v |= hi2 << 16;
            if (lane < 8)
                *reinterpret_cast<uint32_t*>(tile + ((size_t)(2 * wi + (s8 >> 2)) * P.mrp + p * P.KT + kl) * 16 + 4 * (s8 & 3)) = v;
        }
    }
    for (int off = 16; off > 0; off >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, off);
    if (lane == 0) s_c[wp] = csum;
    __syncthreads();
    if (tid == 0) {
        long long tot = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += s_c[i];
        atomicAdd(reinterpret_cast<unsigned long long*>(P.c + k), (unsigned long long)tot);
        if (blockIdx.x == 0) { P.scale[k] = sc; P.weights[k] = w; P.logw[k] = logf(w); }
    }
}

__global__ void em_zero_c_kernel(long long* c, int K, const int* done) {
    if (done && *done) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K) c[i] = 0;
}

// ------------------------------------------------------------------------------------------------
// Row kernel of the E-step.  Thread = data row.
//   INT_SCORES: scores from the int32 digit-plane accumulators (which are cleared for the next E-step);
//   otherwise scratch already holds float scores [K][Npad] (SIMT cross-check).
struct PostArgs {
    int* accE;
    float* scratch;             // [K][Npad] float scores of the SIMT cross-check kernel
    const long long* c;         // 2^-36 fixed point
    const float* scale;
    const float* logw;
    long long N, Npad;
    int K, KT, Mtiles, mrp;
    float* aff;                 // [N][K] posteriors (update)
    float* scores_out;          // [N][K] or null
    uint32_t* colmax;           // [K] max_n aff[n][k] as float bits (update)
    long long* llQ;             // += sum_n round(ll * 2^20) (update)
    int hard, update;
    const int* done;
};

// KREG > 0 (K <= KREG): the scores of a row are read ONCE into registers (3*K independent loads in flight, the
// accumulators cleared behind them) and the three passes run out of registers.  KREG == 0: any K, the passes re-read
// the accumulators (L2 resident) instead.
template <bool INT_SCORES, int KREG>
__global__ void __launch_bounds__(128) em_posterior_kernel(PostArgs P) {
    if (P.done && *P.done) return;
    extern __shared__ __align__(16) uint8_t psm[];
    uint32_t* s_max = reinterpret_cast<uint32_t*>(psm);               // [K]
    float* s_aff = reinterpret_cast<float*>(psm + (((size_t)P.K * 4 + 15) & ~(size_t)15));   // KREG > 0: [128][K + 1] staging of the posteriors
    __shared__ long long s_ll;
    const int tid = threadIdx.x;
    const long long n = blockIdx.x * 128LL + tid;
    const bool live = n < P.N;
    for (int k = tid; k < P.K; k += 128) s_max[k] = 0u;
    if (tid == 0) s_ll = 0;
    __syncthreads();
    const size_t pstride = (size_t)P.KT * P.Npad;
    // template score of component k in double: exact integer digit-plane sums, one scaling (or the SIMT kernel's
    // float score); `clear` zeroes the accumulators behind the last read so that the next E-step starts from zero
    auto score = [&](int k, bool clear) -> double {
        if (INT_SCORES) {
            const int mt = k / P.KT, kl = k - mt * P.KT;
            int* a = P.accE + ((size_t)mt * P.mrp + kl) * P.Npad + n;
            const int a0 = a[0], a1 = a[pstride], a2 = a[2 * pstride];
            if (clear) { a[0] = 0; a[pstride] = 0; a[2 * pstride] = 0; }
            const double q = (double)a0 + 256.0 * (double)a1 + 65536.0 * (double)a2;
            return q * (double)P.scale[k] + (double)P.c[k] / EM_C_SCALE;
        }
        return (double)P.scratch[(size_t)k * P.Npad + n];
    };
    constexpr int NR = KREG > 0 ? KREG : 1;
    double sv[NR];
    if (KREG > 0) {
        if (INT_SCORES) {      // all loads first (K <= 42: one M tile, rows plane * KT + k)
            int a0[NR], a1[NR], a2[NR];
            int* a = P.accE + n;
#pragma unroll
            for (int k = 0; k < NR; ++k)
                if (k < P.K) { a0[k] = a[(size_t)k * P.Npad]; a1[k] = a[(size_t)k * P.Npad + pstride]; a2[k] = a[(size_t)k * P.Npad + 2 * pstride]; }
#pragma unroll
            for (int k = 0; k < NR; ++k)
                if (k < P.K) {
                    a[(size_t)k * P.Npad] = 0; a[(size_t)k * P.Npad + pstride] = 0; a[(size_t)k * P.Npad + 2 * pstride] = 0;
                    const double q = (double)a0[k] + 256.0 * (double)a1[k] + 65536.0 * (double)a2[k];
                    sv[k] = q * (double)P.scale[k] + (double)P.c[k] / EM_C_SCALE;
                }
        } else {
#pragma unroll
            for (int k = 0; k < NR; ++k)
                if (k < P.K) sv[k] = (double)P.scratch[(size_t)k * P.Npad + n];
        }
    }
    // pass 1: scores out, running maxima (soft-max in double up to the exponent: the posteriors then carry the
    // weight quantisation only, ~1e-6, instead of the 3e-5 a float32 score of magnitude 500 would add)
    double mx = -INFINITY, smx = -INFINITY;
    int arg = 0;
    auto pass1 = [&](int k, double s) {
        if (live && P.scores_out) P.scores_out[n * P.K + k] = (float)s;
        if (s > smx) { smx = s; arg = k; }                  // hard assignment: first arg-max of the template scores (:273-275)
        mx = fmax(mx, s + (double)P.logw[k]);
    };
    if (KREG > 0) {
#pragma unroll
        for (int k = 0; k < NR; ++k)
            if (k < P.K) pass1(k, sv[k]);
    } else {
        for (int k = 0; k < P.K; ++k) pass1(k, score(k, !P.update));
    }
    if (!P.update) return;
    // pass 2: log-sum-exp (:269-270)
    double sum = 0.0;
    if (KREG > 0) {
#pragma unroll
        for (int k = 0; k < NR; ++k)
            if (k < P.K) sum += (double)expf((float)(sv[k] + (double)P.logw[k] - mx));
    } else {
        for (int k = 0; k < P.K; ++k) sum += (double)expf((float)(score(k, false) + (double)P.logw[k] - mx));
    }
    const float inv = (float)(1.0 / sum);
    long long my_ll = live ? llrint((log(sum) + mx) * EM_LL_SCALE) : 0;
    // pass 3: posteriors and their column maxima
    auto pass3 = [&](int k, double s) {
        const double v = s + (double)P.logw[k];
        float p = 0.f;
        if (live) p = P.hard ? (k == arg ? 1.f : 0.f) : fminf(expf((float)(v - mx)) * inv, 1.f);
        if (KREG > 0) s_aff[tid * (P.K + 1) + k] = p;          // written out as one contiguous 128 x K block below
        else if (live) P.aff[n * P.K + k] = p;
        const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(p));     // p >= 0: bit patterns order like values
        if ((tid & 31) == 0 && wmax) atomicMax(&s_max[k], wmax);
    };
    if (KREG > 0) {
#pragma unroll
        for (int k = 0; k < NR; ++k)
            if (k < P.K) pass3(k, sv[k]);
    } else {
        for (int k = 0; k < P.K; ++k) pass3(k, score(k, true));
    }
    for (int off = 16; off > 0; off >>= 1) my_ll += __shfl_xor_sync(0xffffffffu, my_ll, off);
    if ((tid & 31) == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&s_ll), (unsigned long long)my_ll);
    __syncthreads();
    if (KREG > 0) {      // rows n0 .. n0+127 of aff are one contiguous block of 128 * K floats: coalesced stores
        const long long base = blockIdx.x * 128LL * P.K;
        const long long lim = P.N * P.K;
        for (int e = tid; e < 128 * P.K; e += 128) {
            const int r = e / P.K, k = e - r * P.K;
            if (base + e < lim) P.aff[base + e] = s_aff[r * (P.K + 1) + k];
        }
    }
    const int slot = blockIdx.x % EM_SLOTS;
    for (int k = tid; k < P.K; k += 128)
        if (s_max[k]) atomicMax(P.colmax + (size_t)slot * P.K + k, s_max[k]);
    if (tid == 0) atomicAdd(reinterpret_cast<unsigned long long*>(P.llQ + slot), (unsigned long long)s_ll);
}

// K <= 32: four threads per row.  CTA = 32 rows x 4 warps; warp q owns the components k = q, q + 4, ... (<= 8 of them) of
// the 32 rows (lane = row: every accumulator load is one coalesced 128-byte line), so a thread reads its 3 * 8 integers once,
// keeps 8 scores in registers (full occupancy instead of the 128 registers of a thread-per-row kernel) and the row-wise
// max / sum go through shared memory (two barriers).
template <bool INT_SCORES>
__global__ void __launch_bounds__(128) em_posterior4_kernel(PostArgs P) {
    if (P.done && *P.done) return;
    extern __shared__ __align__(16) uint8_t psm[];
    float* s_aff = reinterpret_cast<float*>(psm);                      // [32][K + 1]
    __shared__ double s_mx[4][32], s_sc[4][32], s_sum[4][32];
    __shared__ int s_arg[4][32];
    const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
    const long long n = blockIdx.x * 32LL + lane;
    const bool live = n < P.N;
    const size_t pstride = (size_t)P.KT * P.Npad;
    constexpr int KQ = 8;
    double sv[KQ];
    int a0[KQ], a1[KQ], a2[KQ];
    if (INT_SCORES) {
        int* a = P.accE + n;
#pragma unroll
        for (int j = 0; j < KQ; ++j) {
            const int k = q + 4 * j;
            if (k < P.K) { a0[j] = a[(size_t)k * P.Npad]; a1[j] = a[(size_t)k * P.Npad + pstride]; a2[j] = a[(size_t)k * P.Npad + 2 * pstride]; }
        }
#pragma unroll
        for (int j = 0; j < KQ; ++j) {
            const int k = q + 4 * j;
            if (k < P.K) {
                a[(size_t)k * P.Npad] = 0; a[(size_t)k * P.Npad + pstride] = 0; a[(size_t)k * P.Npad + 2 * pstride] = 0;
                const double qq = (double)a0[j] + 256.0 * (double)a1[j] + 65536.0 * (double)a2[j];
                sv[j] = qq * (double)P.scale[k] + (double)P.c[k] * (1.0 / EM_C_SCALE);
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < KQ; ++j) {
            const int k = q + 4 * j;
            if (k < P.K) sv[j] = (double)P.scratch[(size_t)k * P.Npad + n];
        }
    }
    // partial maxima of this warp's components
    double mx = -INFINITY, smx = -INFINITY;
    int arg = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < KQ; ++j) {
        const int k = q + 4 * j;
        if (k < P.K) {
            if (live && P.scores_out) P.scores_out[n * P.K + k] = (float)sv[j];
            if (sv[j] > smx) { smx = sv[j]; arg = k; }
            mx = fmax(mx, sv[j] + (double)P.logw[k]);
        }
    }
    if (!P.update) return;
    s_mx[q][lane] = mx; s_sc[q][lane] = smx; s_arg[q][lane] = arg;
    __syncthreads();
    mx = fmax(fmax(s_mx[0][lane], s_mx[1][lane]), fmax(s_mx[2][lane], s_mx[3][lane]));
    // hard assignment: first arg-max of the template scores (:273-275): largest score, lowest index among equals
    smx = s_sc[0][lane]; arg = s_arg[0][lane];
#pragma unroll
    for (int w = 1; w < 4; ++w) {
        const double o = s_sc[w][lane];
        const int oa = s_arg[w][lane];
        if (o > smx || (o == smx && oa < arg)) { smx = o; arg = oa; }
    }
    float ev[KQ];
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < KQ; ++j) {
        const int k = q + 4 * j;
        ev[j] = 0.f;
        if (k < P.K) { ev[j] = expf((float)(sv[j] + (double)P.logw[k] - mx)); sum += (double)ev[j]; }
    }
    s_sum[q][lane] = sum;
    __syncthreads();
    sum = (s_sum[0][lane] + s_sum[1][lane]) + (s_sum[2][lane] + s_sum[3][lane]);
    const float inv = (float)(1.0 / sum);
#pragma unroll
    for (int j = 0; j < KQ; ++j) {
        const int k = q + 4 * j;
        if (k < P.K) {
            float p = 0.f;
            if (live) p = P.hard ? (k == arg ? 1.f : 0.f) : fminf(ev[j] * inv, 1.f);
            s_aff[lane * (P.K + 1) + k] = p;
            const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(p));     // p >= 0: bit patterns order like values
            if (lane == 0 && wmax) atomicMax(P.colmax + (size_t)(blockIdx.x % EM_SLOTS) * P.K + k, wmax);
        }
    }
    if (q == 0) {
        long long my_ll = live ? llrint((log(sum) + mx) * EM_LL_SCALE) : 0;
        for (int off = 16; off > 0; off >>= 1) my_ll += __shfl_xor_sync(0xffffffffu, my_ll, off);
        if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(P.llQ + blockIdx.x % EM_SLOTS), (unsigned long long)my_ll);
    }
    __syncthreads();
    // rows n0 .. n0+31 of aff are one contiguous block of 32 * K floats: coalesced stores
    const long long base = blockIdx.x * 32LL * P.K, lim = P.N * P.K;
    for (int e = threadIdx.x; e < 32 * P.K; e += 128) {
        const int r = e / P.K, k = e - r * P.K;
        if (base + e < lim) P.aff[base + e] = s_aff[r * (P.K + 1) + k];
    }
}

// column maxima of caller-supplied posteriors (initialisation; M-step from given posteriors)
__global__ void em_colmax_kernel(const float* __restrict__ aff, long long N, int K, uint32_t* __restrict__ colmax, const int* done) {
    if (done && *done) return;
    const long long total = N * K;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const float p = fmaxf(aff[i], 0.f);
        if (p > 0.f) atomicMax(colmax + (size_t)(blockIdx.x % EM_SLOTS) * K + (i % K), __float_as_uint(p));
    }
}

// Position of row n in its 128-row chunk of the M-step contraction: word i = (n % 128) / 32, r = n % 32 = 8 jj + s
// ->  k-group 2i + s/4, byte 4 (s%4) + jj   (the order the bitgemm expanders lay a word out).
__device__ __forceinline__ int em_tile_pos(int t /* 0..127 */) {
    const int i = t >> 5, r = t & 31, s = r & 7, jj = r >> 3;
    return ((2 * i + (s >> 2)) << 4) | (4 * (s & 3) + jj);        // (k-group << 4) | byte
}

// Posteriors -> the M-step's int8 operand.  CTA = 128 consecutive rows = one chunk of the contraction.
// q(n,k) = round(aff[n][k] / s_k * EM_QMAX), three balanced base-256 digits, one tile per M tile; respQ[slot][k] += sum_n q.
// Every CTA reduces the column-maximum replicas to the scales s_k itself (EM_SLOTS * K loads); CTA 0 publishes them.
__global__ void __launch_bounds__(128) em_aff_planes_kernel(const float* __restrict__ aff, long long N, long long Npad, int K, int KT,
                                                            int Mtiles, int mrp, uint8_t* __restrict__ ablob, EmSmall S,
                                                            double* __restrict__ stats_tail, int write_ll, const int* done) {
    if (done && *done) return;
    extern __shared__ __align__(16) uint8_t psm[];
    __shared__ int s_last;
    uint8_t* tile = psm;
    int* s_resp = reinterpret_cast<int*>(psm + (size_t)8 * mrp * 16);
    float* s_isk = reinterpret_cast<float*>(s_resp + K);                                  // 1 / s_k (0 for an empty column)
    float* s_aff = reinterpret_cast<float*>(psm + (((size_t)8 * mrp * 16 + (size_t)K * 8 + 15) & ~(size_t)15));
    const int tid = threadIdx.x;
    const long long n = blockIdx.x * 128LL + tid;
    const bool live = n < N;
    for (int k = tid; k < K; k += 128) {
        s_resp[k] = 0;
        uint32_t mx = 0;
        for (int sl = 0; sl < EM_SLOTS; ++sl) mx = max(mx, S.colmax[(size_t)sl * K + k]);
        const float sk = em_col_scale(mx);
        s_isk[k] = sk >= 1.17549435e-38f ? 1.f / sk : 0.f;                               // power-of-two s_k: aff / s_k is exact and <= 1
        if (blockIdx.x == 0) S.sk[k] = (double)sk;                                       // for the statistics tail / em_pack_kernel
    }
    const bool staged = K <= 32;
    if (staged) {        // the block's 128 x K posteriors are contiguous in global memory: coalesced loads, pitch K + 1 in smem
        const long long base = blockIdx.x * 128LL * K, lim = N * K;
        for (int e = tid; e < 128 * K; e += 128) {
            const int r = e / K, k = e - r * K;
            s_aff[r * (K + 1) + k] = base + e < lim ? aff[base + e] : 0.f;
        }
    }
    __syncthreads();
    const int pos = em_tile_pos(tid);
    const size_t tile_bytes = (size_t)8 * mrp * 16;
    for (int mt = 0; mt < Mtiles; ++mt) {
        for (int i = tid; i < (int)(tile_bytes / 16); i += 128) reinterpret_cast<uint4*>(tile)[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        const int k1 = (mt + 1) * KT < K ? (mt + 1) * KT : K;
        for (int k = mt * KT; k < k1; ++k) {
            const float p = live ? fmaxf(staged ? s_aff[tid * (K + 1) + k] : aff[n * K + k], 0.f) : 0.f;
            int q = (int)rintf(p * s_isk[k] * (float)EM_QMAX);
            const int qsum = __reduce_add_sync(0xffffffffu, q);
            if ((tid & 31) == 0 && qsum) atomicAdd(&s_resp[k], qsum);
            const int kl = k - mt * KT;
#pragma unroll
            for (int pl = 0; pl < 3; ++pl) {
                const int digit = ((q + 128) & 255) - 128;
                q = (q - digit) >> 8;
                tile[((size_t)(pos >> 4) * mrp + pl * KT + kl) * 16 + (pos & 15)] = (uint8_t)(digit & 255);
            }
        }
        __syncthreads();
        uint4* dst = reinterpret_cast<uint4*>(ablob + ((size_t)mt * (Npad / 128) + blockIdx.x) * tile_bytes);
        for (int i = tid; i < (int)(tile_bytes / 16); i += 128) dst[i] = reinterpret_cast<const uint4*>(tile)[i];
        __syncthreads();
    }
    const int slot = blockIdx.x % EM_SLOTS;
    for (int k = tid; k < K; k += 128)
        if (s_resp[k]) atomicAdd(reinterpret_cast<unsigned long long*>(S.respQ + (size_t)slot * K + k), (unsigned long long)(long long)s_resp[k]);
    // The last CTA to get here turns the integer sums into the float64 tail of the statistics vector (responsibility sums,
    // log-likelihood): warp per component, lanes = the EM_SLOTS replicas.  (Was a launch of its own: ~5 us per iteration.)
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(S.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int lane = tid & 31, wp = tid >> 5;
    for (int k = wp; k < K; k += 4) {
        long long rq = __ldcg(S.respQ + (size_t)lane * K + k);
        for (int off = 16; off > 0; off >>= 1) rq += __shfl_xor_sync(0xffffffffu, rq, off);
        if (lane == 0) stats_tail[k] = (double)rq / (double)EM_QMAX * __ldcg(S.sk + k);
    }
    if (wp == 0 && write_ll) {
        long long ll = __ldcg(S.llQ + lane);
        for (int off = 16; off > 0; off >>= 1) ll += __shfl_xor_sync(0xffffffffu, ll, off);
        if (lane == 0) stats_tail[K] = (double)ll / EM_LL_SCALE;
    }
    if (tid == 0) *S.ticket = 0u;
}

// int32 digit-plane accumulators of the M-step -> float64 numerators; clears the accumulators
__global__ void __launch_bounds__(256) em_pack_kernel(int* __restrict__ accM, int D, int Dpad, int K, int KT, int mrp,
                                                      const double* __restrict__ sk, double* __restrict__ stats, const int* done) {
    if (done && *done) return;
    const int k = blockIdx.y, d = blockIdx.x * 256 + threadIdx.x;       // grid = (ceil(D / 256), K): no index divisions
    if (d >= D) return;
    const int mt = k / KT, kl = k - mt * KT;
    int* a = accM + ((size_t)mt * mrp + kl) * Dpad + d;
    const size_t ps = (size_t)KT * Dpad;
    const long long q = (long long)a[0] + 256LL * (long long)a[ps] + 65536LL * (long long)a[2 * ps];
    a[0] = 0; a[ps] = 0; a[2 * ps] = 0;
    stats[(size_t)k * D + d] = em_q_to_stat(q, __ldg(sk + k));
}

__global__ void em_empty_tail_kernel(double* stats_tail, int n, const int* done) {
    if (done && *done) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) stats_tail[i] = 0.0;
}

// Stopping rule, evaluated on the (all-reduced) log-likelihood.  Mirrors the host loops
//   while (abs((old - new) / old) > tol and it < max_iter) or it < min_iter      (bernoulli_mixture.py:139-156)
//   while new - old > tol                                                          (bernoulli_em.py:24-33)
__global__ void em_control_kernel(EmCtl* ctl, const double* stats_ll) {
    if (ctl->done) return;
    double* trace = reinterpret_cast<double*>(ctl + 1);
    const double ll = *stats_ll;
    if (ctl->ntrace == 0) ctl->ll_old = ctl->mode == 0 ? ll / (1.0 - ctl->tol) - 10.0 : -INFINITY;
    if (ctl->ntrace < ctl->cap) trace[ctl->ntrace] = ll;
    ctl->ntrace++;
    bool go;
    if (ctl->mode == 0)
        go = (fabs((ctl->ll_old - ll) / ctl->ll_old) > ctl->tol && ctl->iterations < ctl->max_iter) || ctl->iterations < ctl->min_iter;
    else
        go = (ll - ctl->ll_old > ctl->tol) && ctl->iterations < ctl->max_iter;
    if (!go) { ctl->done = 1; return; }
    ctl->ll_old = ll;
    ctl->iterations++;
}

// ------------------------------------------------------------------------------------------------
// Sum of the ranks' statistics vectors over NVLink peer memory (one launch, no NCCL call): every rank keeps its
// vector in a symmetric allocation that all peers have mapped; rank r's CTA b
//   1. publishes "rank r has finished writing its vector" by storing the launch's epoch into slot [r][b] of every
//      peer's flag array (system-scope release; the vector was written by earlier kernels of the same stream),
//   2. waits until slot [t][b] of its own flag array carries the epoch for every peer t (system-scope acquire),
//   3. reads its share of all W vectors straight from the peers' memory and adds them in rank order (every rank
//      forms the same sums in the same order: bit-identical results on all ranks) into a LOCAL output vector.
// The caller alternates between two symmetric source vectors (epoch parity): a rank can only overwrite a source two
// iterations later, after it has passed the next launch's step 2, which every peer reaches only after its own reads
// of this launch are complete (stream order) -- so no closing barrier is needed.  Per-CTA flags make step 2 a
// per-CTA wait instead of a grid-wide one.  Once the control block has stopped, every rank skips the launch alike.
#define EM_AR_CTAS 64
struct PeerArgs {
    const uint8_t* const* bufs;     // device array [world]: base of every rank's symmetric allocation
    long long src_off, flag_off;    // byte offsets of the source vector and of the flag array [world][EM_AR_CTAS] uint32
    long long n;                    // doubles
    double* out;                    // one-phase form: local result vector
    long long out_off;              // two-phase form (>= 0): byte offset of the result vector inside every rank's allocation
    unsigned int* epoch;            // [EM_AR_CTAS] launch counters of this rank (zero-initialised once)
    int rank, world;
    const int* done;
};
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// no "memory" clobber: the loads of one element from all peers may be in flight together (an NVLink round trip is ~2 us);
// they are issued after the __syncthreads that follows the flag acquire, which the compiler does not move them across
__device__ __forceinline__ double2 ld_relaxed_sys_f64x2(const double* p) {
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__global__ void __launch_bounds__(1024) em_allreduce_peer_kernel(PeerArgs P) {
    if (P.done && *P.done) return;
    __shared__ unsigned int s_epoch;
    const int b = blockIdx.x;
    if (threadIdx.x == 0) s_epoch = ++P.epoch[b];
    __syncthreads();
    const unsigned int e = s_epoch;
    if ((int)threadIdx.x < P.world) {
        const int t = threadIdx.x;
        // (the vector was written by earlier kernels of this stream: the release below is cumulative over it)
        unsigned int* theirs = reinterpret_cast<unsigned int*>(const_cast<uint8_t*>(P.bufs[t]) + P.flag_off);
        st_release_sys(theirs + (size_t)P.rank * EM_AR_CTAS + b, e);
        const unsigned int* mine = reinterpret_cast<const unsigned int*>(P.bufs[P.rank] + P.flag_off);
        unsigned int spins = 0;
        while (ld_acquire_sys(mine + (size_t)t * EM_AR_CTAS + b) < e) {       // epochs only grow: "<" tolerates a peer one launch ahead
            __nanosleep(20);
            if (++spins > (1u << 26)) {
                printf("tsr_b200: peer allreduce timed out waiting for rank %d (rank %d block %d epoch %u)\n", t, P.rank, b, e);
                __trap();
            }
        }
    }
    __syncthreads();
    const long long npair = P.n >> 1;
    for (long long i = (long long)b * blockDim.x + threadIdx.x; i < npair; i += (long long)gridDim.x * blockDim.x) {
        double2 acc = make_double2(0.0, 0.0);
        for (int t0 = 0; t0 < P.world; t0 += 8) {          // up to 8 peers' loads in flight, added in rank order
            double2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (t0 + u < P.world) {
                    const double* q = reinterpret_cast<const double*>(P.bufs[t0 + u] + P.src_off) + 2 * i;
                    v[u] = (t0 + u == P.rank) ? *reinterpret_cast<const double2*>(q) : ld_relaxed_sys_f64x2(q);
                }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (t0 + u < P.world) { acc.x += v[u].x; acc.y += v[u].y; }
        }
        reinterpret_cast<double2*>(P.out)[i] = acc;
    }
    if ((P.n & 1) && b == 0 && threadIdx.x == 0) {
        double acc = 0.0;
        for (int t = 0; t < P.world; ++t) {
            double v;
            asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(reinterpret_cast<const double*>(P.bufs[t] + P.src_off) + P.n - 1) : "memory");
            acc += v;
        }
        P.out[P.n - 1] = acc;
    }
}

// Two-phase form for more than two ranks (reduce-scatter + all-gather by remote stores): rank r sums only ITS 1/W share
// of the vector (reads (W-1)/W of a vector over NVLink instead of W-1 vectors) and stores the result into the result
// slot of every rank's allocation; a second per-CTA flag round ("my share has landed everywhere") closes the launch.
// The sums are still formed in rank order by exactly one rank each, so every rank ends with the same bits.
// Flags: [2][world][EM_AR_CTAS] (start, landed).  A peer's stores into this rank's result slot of launch i+1 happen
// after that peer passed the start flags of launch i+1, i.e. after this rank's own launch i+1 began: the kernels that
// read the result of launch i (stopping rule, finalize) are earlier in this rank's stream.
__global__ void __launch_bounds__(1024) em_allreduce_peer2_kernel(PeerArgs P) {
    if (P.done && *P.done) return;
    __shared__ unsigned int s_epoch;
    const int b = blockIdx.x, W = P.world, r = P.rank;
    if (threadIdx.x == 0) s_epoch = ++P.epoch[b];
    __syncthreads();
    const unsigned int e = s_epoch;
    const unsigned int* mine = reinterpret_cast<const unsigned int*>(P.bufs[r] + P.flag_off);
    auto flag_round = [&](int which) {
        if ((int)threadIdx.x < W) {
            const int t = threadIdx.x;
            unsigned int* theirs = reinterpret_cast<unsigned int*>(const_cast<uint8_t*>(P.bufs[t]) + P.flag_off) + (size_t)which * W * EM_AR_CTAS;
            st_release_sys(theirs + (size_t)r * EM_AR_CTAS + b, e);
            const unsigned int* m = mine + (size_t)which * W * EM_AR_CTAS + (size_t)t * EM_AR_CTAS + b;
            unsigned int spins = 0;
            while (ld_acquire_sys(m) < e) {
                __nanosleep(20);
                if (++spins > (1u << 26)) {
                    printf("tsr_b200: peer allreduce (phase %d) timed out waiting for rank %d (rank %d block %d epoch %u)\n", which, t, r, b, e);
                    __trap();
                }
            }
        }
        __syncthreads();
    };
    flag_round(0);
    const long long npair = P.n >> 1;
    const long long cp = (npair + W - 1) / W;
    const long long lo = (long long)r * cp, hi = lo + cp < npair ? lo + cp : npair;
    for (long long i = lo + (long long)b * blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x) {
        double2 acc = make_double2(0.0, 0.0);
        for (int t0 = 0; t0 < W; t0 += 8) {
            double2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (t0 + u < W) {
                    const double* q = reinterpret_cast<const double*>(P.bufs[t0 + u] + P.src_off) + 2 * i;
                    v[u] = (t0 + u == r) ? *reinterpret_cast<const double2*>(q) : ld_relaxed_sys_f64x2(q);
                }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (t0 + u < W) { acc.x += v[u].x; acc.y += v[u].y; }
        }
        for (int t = 0; t < W; ++t) {
            double* q = reinterpret_cast<double*>(const_cast<uint8_t*>(P.bufs[t]) + P.out_off) + 2 * i;
            asm volatile("st.relaxed.sys.global.v2.f64 [%0], {%1,%2};" ::"l"(q), "d"(acc.x), "d"(acc.y) : "memory");
        }
    }
    if ((P.n & 1) && r == 0 && b == 0 && threadIdx.x == 0) {
        double acc = 0.0;
        for (int t = 0; t < W; ++t) {
            double v;
            asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(reinterpret_cast<const double*>(P.bufs[t] + P.src_off) + P.n - 1) : "memory");
            acc += v;
        }
        for (int t = 0; t < W; ++t)
            asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(reinterpret_cast<double*>(const_cast<uint8_t*>(P.bufs[t]) + P.out_off) + P.n - 1), "d"(acc) : "memory");
    }
    __threadfence_system();          // this thread's remote stores are visible system-wide before the CTA's "landed" flags go out
    __syncthreads();
    flag_round(1);
}

// XT[d][n] = X[n][d]: each warp transposes a 32x32 bit block with 32 ballots.
__global__ void transpose_bits_kernel(const uint32_t* __restrict__ X, long long N, int D, int rowwords,
                                      uint32_t* __restrict__ XT, long long nt_words) {
    const int lane = threadIdx.x & 31;
    const long long nblk_n = nt_words;                       // 32-row blocks (incl. zero padding)
    const int nblk_d = (D + 31) / 32;
    const long long total = nblk_n * nblk_d;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; wid < total; wid += nwarps) {
        const long long bn = wid / nblk_d;
        const int bd = (int)(wid - bn * nblk_d);
        const long long n = bn * 32 + lane;
        const uint32_t word = n < N ? X[n * rowwords + bd] : 0u;
        uint32_t mine = 0;
#pragma unroll
        for (int bit = 0; bit < 32; ++bit) {
            const uint32_t col = __ballot_sync(0xffffffffu, (word >> bit) & 1u);   // bit `bit` of 32 rows
            if (lane == bit) mine = col;
        }
        const int d = bd * 32 + lane;
        if (d < D) XT[(size_t)d * nt_words + bn] = mine;
    }
}

}  // namespace tsr

using namespace tsr;

extern "C" {

size_t tsr_em_workspace_bytes(long long N, int D, int K) {
    if (N < 0 || N >= (1ll << 24) || D < 1 || K < 1 || K > EM_MAX_K) return 0;
    return em_shape(N, D, K).total;
}

long long tsr_em_npad(long long N) { return N < 0 ? 0 : (N + 255) / 256 * 256; }

size_t tsr_em_control_bytes(int max_trace) { return max_trace < 0 ? 0 : sizeof(EmCtl) + (size_t)max_trace * 8; }

int tsr_em_workspace_init(void* workspace, long long N, int D, int K, void* stream) {
    TSR_REQUIRE(workspace, "tsr_em_workspace_init: null pointer");
    TSR_REQUIRE(N >= 0 && N < (1ll << 24) && D >= 1 && K >= 1 && K <= EM_MAX_K,
                "tsr_em_workspace_init: need 1 <= num_mix <= %d and fewer than 2^24 rows per rank", EM_MAX_K);
    int rc = require_sm100("tsr_em_workspace_init");
    if (rc) return rc;
    const EmShape s = em_shape(N, D, K);
    // weight / posterior tiles (rows no component maps to must read as zero), the int32 accumulators, the small scratch
    TSR_CUDA(cudaMemsetAsync(workspace, 0, s.off_scratch, (cudaStream_t)stream));
    return TSR_OK;
}

int tsr_em_control_init(void* ctl, int max_trace, double tol, int mode, int min_iter, int max_iter, void* stream) {
    TSR_REQUIRE(ctl && max_trace >= 1 && (mode == 0 || mode == 1), "tsr_em_control_init: bad arguments");
    int rc = require_sm100("tsr_em_control_init");
    if (rc) return rc;
    EmCtl h;
    memset(&h, 0, sizeof(h));
    h.mode = mode; h.min_iter = min_iter; h.max_iter = max_iter; h.cap = max_trace; h.tol = tol;
    TSR_CUDA(cudaMemcpyAsync(ctl, &h, sizeof(h), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    TSR_CUDA(cudaStreamSynchronize((cudaStream_t)stream));      // `h` is a stack object
    return TSR_OK;
}

int tsr_em_control_step(void* ctl, const double* stats, int D, int K, void* stream) {
    TSR_REQUIRE(ctl && stats && D >= 1 && K >= 1, "tsr_em_control_step: bad arguments");
    int rc = require_sm100("tsr_em_control_step");
    if (rc) return rc;
    em_control_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((EmCtl*)ctl, stats + (size_t)K * D + K);
    return check_launch("tsr_em_control_step");
}

static int em_prepare(const EmShape& s, uint8_t* ws, const float* logT, const float* logInvT, const int* done, cudaStream_t st) {
    ParamArgs P;
    memset(&P, 0, sizeof(P));
    P.logT = const_cast<float*>(logT); P.logInvT = const_cast<float*>(logInvT);
    P.D = s.D; P.K = s.K; P.KT = s.KT; P.mrp = s.mrp; P.nchunk = s.nchunkE;
    P.wblob = ws + s.off_wblob; P.c = (long long*)(ws + s.off_c); P.scale = (float*)(ws + s.off_scale); P.done = done;
    em_params_kernel<false><<<s.K, 256, 0, st>>>(P);
    return check_launch("tsr_em_estep(prepare)");
}

// posteriors (device, [N][K]) -> digit planes + responsibility sums (+ log-likelihood) in the statistics tail
static int em_planes(const EmShape& s, uint8_t* ws, const float* aff, double* stats, int write_ll, const int* done, cudaStream_t st) {
    const size_t psm = ((s.tile_bytes + (size_t)s.K * 8 + 15) & ~(size_t)15) + (s.K <= 32 ? (size_t)128 * (s.K + 1) * 4 : 0);
    static size_t psm_set = 0;
    if (psm > psm_set) {
        TSR_CUDA(cudaFuncSetAttribute(em_aff_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
        psm_set = psm;
    }
    const EmSmall sm = em_small(ws, s);
    em_aff_planes_kernel<<<(unsigned)(s.Npad / 128), 128, psm, st>>>(aff, s.N, s.Npad, s.K, s.KT, s.Mtiles, s.mrp, ws + s.off_ablob, sm,
                                                                     stats + (size_t)s.K * s.D, write_ll, done);
    return check_launch("em_aff_planes_kernel");
}

size_t tsr_em_tiled_words(long long R, int rw) {
    if (R < 0 || rw < 1) return 0;
    return (size_t)((R + 255) / 256 * 256) * (size_t)((rw + 3) / 4) * 4;
}

int tsr_em_tile_bits(const uint32_t* bits, long long R, int rw, uint32_t* tiled, void* stream) {
    TSR_REQUIRE(bits && tiled, "tsr_em_tile_bits: null pointer");
    TSR_REQUIRE(R >= 1 && rw >= 1, "tsr_em_tile_bits: bad sizes");
    int rc = require_sm100("tsr_em_tile_bits");
    if (rc) return rc;
    const long long ntile = (R + 255) / 256 * 2;
    tile_bits_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(bits, R, rw, (rw + 3) / 4, ntile, reinterpret_cast<uint4*>(tiled));
    return check_launch("tsr_em_tile_bits");
}

int tsr_em_estep(const uint32_t* Xbits, const uint32_t* Xtiled_or_null, long long N, int D, int rowwords, int K, const float* logT_or_null,
                 const float* logInvT_or_null, const float* logw, float* aff_or_null, float* scores_or_null,
                 double* stats_or_null, int hard, int update, void* workspace, const void* ctl_or_null, int impl,
                 void* stream) {
    TSR_REQUIRE(Xbits && logw && workspace, "tsr_em_estep: null pointer");
    TSR_REQUIRE((logT_or_null == nullptr) == (logInvT_or_null == nullptr), "tsr_em_estep: pass both logT and logInvT, or neither");
    TSR_REQUIRE(K >= 1 && K <= EM_MAX_K, "tsr_em_estep: need 1 <= num_mix <= %d (got %d)", EM_MAX_K, K);
    TSR_REQUIRE(N >= 0 && N < (1ll << 24) && D >= 1 && rowwords == tsr_rowwords(D), "tsr_em_estep: bad sizes (fewer than 2^24 rows per rank)");
    TSR_REQUIRE(impl == TSR_IMPL_TENSOR || impl == TSR_IMPL_SIMT, "tsr_em_estep: unknown impl %d", impl);
    TSR_REQUIRE(!update || stats_or_null, "tsr_em_estep: update needs the statistics vector");
    TSR_REQUIRE(impl == TSR_IMPL_TENSOR || logT_or_null, "tsr_em_estep(simt): needs logT / logInvT");
    int rc = require_sm100("tsr_em_estep");
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const EmShape s = em_shape(N, D, K);
    uint8_t* ws = (uint8_t*)workspace;
    const int* done = ctl_or_null ? &((const EmCtl*)ctl_or_null)->done : nullptr;
    if (N == 0) {          // a rank without rows contributes zeros to the allreduce
        if (update) em_empty_tail_kernel<<<(K + 1 + 255) / 256, 256, 0, st>>>(stats_or_null + (size_t)K * D, K + 1, done);
        return check_launch("tsr_em_estep(empty)");
    }
    if (logT_or_null) {
        rc = em_prepare(s, ws, logT_or_null, logInvT_or_null, done, st);
        if (rc) return rc;
    }
    const EmSmall sm = em_small(ws, s);
    PostArgs P;
    memset(&P, 0, sizeof(P));
    P.accE = (int*)(ws + s.off_accE); P.scratch = (float*)(ws + s.off_scratch); P.c = (const long long*)(ws + s.off_c);
    P.scale = (const float*)(ws + s.off_scale); P.logw = logw; P.N = N; P.Npad = s.Npad; P.K = K; P.KT = s.KT; P.Mtiles = s.Mtiles;
    P.mrp = s.mrp; P.aff = aff_or_null ? aff_or_null : (float*)(ws + s.off_aff); P.scores_out = scores_or_null; P.hard = hard;
    P.update = update; P.done = done; P.colmax = sm.colmax; P.llQ = sm.llQ;
    if (update)     // (a stopped loop skips the kernels that would consume this, so clearing unconditionally is harmless)
        TSR_CUDA(cudaMemsetAsync(ws + s.off_small, 0, s.off_scratch - s.off_small, st));
    const size_t psm = (((size_t)K * 4 + 15) & ~(size_t)15) + (K <= 32 ? (size_t)128 * (K + 1) * 4 : 0);
    const unsigned pgrid = (unsigned)(s.Npad / 128);
    if (impl == TSR_IMPL_SIMT) {
        long long blocks = (N * K + 127) / 128;
        const long long cap = (long long)sm_count() * 32;
        if (blocks > cap) blocks = cap;
        em_scores_simt_kernel<<<(unsigned)blocks, 128, 0, st>>>(Xbits, N, D, rowwords, K, logT_or_null, logInvT_or_null, P.c, P.scratch,
                                                               s.Npad, done);
        rc = check_launch("tsr_em_estep(simt scores)");
        if (rc) return rc;
        if (K <= 32) em_posterior4_kernel<false><<<(unsigned)(s.Npad / 32), 128, (size_t)32 * (K + 1) * 4, st>>>(P);
        else em_posterior_kernel<false, 0><<<pgrid, 128, psm, st>>>(P);
    } else {
        rc = launch_bitgemm(Xbits, Xtiled_or_null, N, rowwords, ws + s.off_wblob, P.accE, s.Npad, s.nchunkE, s, done, st);
        if (rc) return rc;
        if (K <= 32) em_posterior4_kernel<true><<<(unsigned)(s.Npad / 32), 128, (size_t)32 * (K + 1) * 4, st>>>(P);
        else em_posterior_kernel<true, 0><<<pgrid, 128, psm, st>>>(P);
    }
    rc = check_launch("tsr_em_estep(posterior)");
    if (rc || !update) return rc;
    return em_planes(s, ws, P.aff, stats_or_null, 1, done, st);
}

int tsr_transpose_bits(const uint32_t* Xbits, long long N, int D, int rowwords, uint32_t* XTbits, long long N_pad,
                       void* stream) {
    TSR_REQUIRE(Xbits && XTbits, "tsr_transpose_bits: null pointer");
    TSR_REQUIRE(N >= 1 && D >= 1 && rowwords == tsr_rowwords(D) && N_pad >= N && N_pad % 64 == 0, "tsr_transpose_bits: bad sizes");
    int rc = require_sm100("tsr_transpose_bits");
    if (rc) return rc;
    transpose_bits_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(Xbits, N, D, rowwords, XTbits, N_pad / 32);
    return check_launch("tsr_transpose_bits");
}

int tsr_em_mstep(const uint32_t* Xbits, const uint32_t* XTbits_or_null, const uint32_t* XTtiled_or_null, long long N, int D, int rowwords, int K,
                 const float* aff_or_null, double* stats, void* workspace, const void* ctl_or_null, int impl, void* stream) {
    TSR_REQUIRE(Xbits && stats && workspace, "tsr_em_mstep: null pointer");
    TSR_REQUIRE(K >= 1 && K <= EM_MAX_K, "tsr_em_mstep: need 1 <= num_mix <= %d (got %d)", EM_MAX_K, K);
    TSR_REQUIRE(N >= 0 && N < (1ll << 24), "tsr_em_mstep: at most 2^24 rows per rank (int32 digit-plane accumulators)");
    TSR_REQUIRE(D >= 1 && rowwords == tsr_rowwords(D), "tsr_em_mstep: bad sizes");
    TSR_REQUIRE(impl == TSR_IMPL_TENSOR || impl == TSR_IMPL_SIMT, "tsr_em_mstep: unknown impl %d", impl);
    int rc = require_sm100("tsr_em_mstep");
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) {          // a rank without rows contributes zeros to the allreduce
        TSR_CUDA(cudaMemsetAsync(stats, 0, ((size_t)K * D + (aff_or_null ? K : 0)) * 8, st));
        return TSR_OK;
    }
    const EmShape s = em_shape(N, D, K);
    uint8_t* ws = (uint8_t*)workspace;
    const int* done = ctl_or_null ? &((const EmCtl*)ctl_or_null)->done : nullptr;
    const EmSmall sm = em_small(ws, s);
    if (aff_or_null) {     // digit planes from the caller's posteriors (otherwise the last E-step left them in the workspace)
        TSR_CUDA(cudaMemsetAsync(ws + s.off_small, 0, s.off_scratch - s.off_small, st));
        long long blocks = (N * K + 255) / 256;
        const long long cap = (long long)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        em_colmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(aff_or_null, N, K, sm.colmax, done);
        rc = em_planes(s, ws, aff_or_null, stats, 0, done, st);
        if (rc) return rc;
    }
    if (impl == TSR_IMPL_SIMT) {
        const float* aff = aff_or_null ? aff_or_null : (const float*)(ws + s.off_aff);
        long long blocks = ((long long)K * D + 127) / 128;
        const long long cap = (long long)sm_count() * 32;
        if (blocks > cap) blocks = cap;
        em_mstep_simt_kernel<<<(unsigned)blocks, 128, 0, st>>>(Xbits, N, D, rowwords, K, aff, sm.sk, stats, done);
        return check_launch("tsr_em_mstep(simt)");
    }
    TSR_REQUIRE(XTbits_or_null || XTtiled_or_null, "tsr_em_mstep: the tensor path needs the transposed bits (tsr_transpose_bits), plain or tiled");
    rc = launch_bitgemm(XTbits_or_null, XTtiled_or_null, D, (int)(s.Npad / 32), ws + s.off_ablob, (int*)(ws + s.off_accM), s.Dpad,
                        (int)s.nchunkM, s, done, st);
    if (rc) return rc;
    em_pack_kernel<<<dim3((unsigned)((D + 255) / 256), (unsigned)K), 256, 0, st>>>((int*)(ws + s.off_accM), D, s.Dpad, K, s.KT, s.mrp, sm.sk,
                                                                                   stats, done);
    return check_launch("tsr_em_mstep(pack)");
}

size_t tsr_em_peer_flag_bytes(int world) { return world < 1 ? 0 : 2 * (size_t)world * EM_AR_CTAS * sizeof(unsigned int); }
size_t tsr_em_peer_epoch_bytes(void) { return EM_AR_CTAS * sizeof(unsigned int); }

int tsr_em_allreduce_peer(const void* const* peer_bufs_dev, int rank, int world, long long src_offset_bytes, long long flag_offset_bytes,
                          long long n_doubles, double* out, long long out_offset_bytes, void* epoch_counters, const void* ctl_or_null,
                          void* stream) {
    TSR_REQUIRE(peer_bufs_dev && epoch_counters && (out || out_offset_bytes >= 0), "tsr_em_allreduce_peer: null pointer");
    TSR_REQUIRE(out_offset_bytes < 0 || (out_offset_bytes & 15) == 0, "tsr_em_allreduce_peer: result offset must be 16-byte aligned");
    TSR_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, "tsr_em_allreduce_peer: bad rank %d / world %d", rank, world);
    TSR_REQUIRE(n_doubles >= 0 && (src_offset_bytes & 15) == 0 && (flag_offset_bytes & 3) == 0 && (out_offset_bytes >= 0 || (reinterpret_cast<uintptr_t>(out) & 15) == 0),
                "tsr_em_allreduce_peer: offsets / output must be 16-byte aligned");
    int rc = require_sm100("tsr_em_allreduce_peer");
    if (rc) return rc;
    PeerArgs P;
    P.bufs = reinterpret_cast<const uint8_t* const*>(peer_bufs_dev);
    P.src_off = src_offset_bytes; P.flag_off = flag_offset_bytes; P.n = n_doubles; P.out = out; P.out_off = out_offset_bytes;
    P.epoch = (unsigned int*)epoch_counters; P.rank = rank; P.world = world;
    P.done = ctl_or_null ? &((const EmCtl*)ctl_or_null)->done : nullptr;
    if (out_offset_bytes >= 0) em_allreduce_peer2_kernel<<<EM_AR_CTAS, 1024, 0, (cudaStream_t)stream>>>(P);
    else em_allreduce_peer_kernel<<<EM_AR_CTAS, 1024, 0, (cudaStream_t)stream>>>(P);
    return check_launch("tsr_em_allreduce_peer");
}

int tsr_em_finalize(const double* stats, long long N_total, int D, int K, float min_prob, float* templates, float* logT,
                    float* logInvT, float* weights, float* logw, void* workspace_or_null, long long N_local,
                    const void* ctl_or_null, void* stream) {
    TSR_REQUIRE(stats && templates && logT && logInvT && weights && logw, "tsr_em_finalize: null pointer");
    TSR_REQUIRE(K >= 1 && K <= EM_MAX_K && D >= 1 && N_total >= 1, "tsr_em_finalize: bad sizes");
    int rc = require_sm100("tsr_em_finalize");
    if (rc) return rc;
    ParamArgs P;
    memset(&P, 0, sizeof(P));
    P.stats = stats; P.N_total = N_total; P.min_prob = min_prob; P.templates = templates; P.logT = logT; P.logInvT = logInvT;
    P.weights = weights; P.logw = logw; P.D = D; P.K = K;
    P.done = ctl_or_null ? &((const EmCtl*)ctl_or_null)->done : nullptr;
    if (workspace_or_null) {       // also lay the next E-step's weight planes out (tsr_em_estep with logT = NULL uses them)
        TSR_REQUIRE(N_local >= 0 && N_local < (1ll << 24), "tsr_em_finalize: bad N_local");
        const EmShape s = em_shape(N_local, D, K);
        uint8_t* ws = (uint8_t*)workspace_or_null;
        P.KT = s.KT; P.mrp = s.mrp; P.nchunk = s.nchunkE;
        P.wblob = ws + s.off_wblob; P.c = (long long*)(ws + s.off_c); P.scale = (float*)(ws + s.off_scale);
        if (min_prob > 0.f && min_prob < 0.5f) {       // the clip bounds the weights: one fully parallel pass
            cudaStream_t st = (cudaStream_t)stream;
            em_zero_c_kernel<<<(K + 255) / 256, 256, 0, st>>>(P.c, K, P.done);
            const float wbound = logf(1.f - min_prob) - logf(min_prob);
            dim3 grid((unsigned)((s.nchunkE * 4 + 7) / 8), (unsigned)K);
            em_finalize_fast_kernel<<<grid, 256, 0, st>>>(P, wbound * 1.0001f);
            return check_launch("tsr_em_finalize");
        }
    } else {
        P.KT = K < EM_KT_MAX ? K : EM_KT_MAX;
    }
    em_params_kernel<true><<<K, 256, 0, (cudaStream_t)stream>>>(P);
    return check_launch("tsr_em_finalize");
}

}  // extern "C"
