// detect.cu -- K3: sliding Bernoulli-template log-likelihood-ratio detection
// (compute_likelihood_linear_filter.pyx:135-176 `detect`, one call per mixture component in
// get_train_data.py:964-992; fused multi-template form as detect_float_max :221-278).
//
//   scores[g][m] = sum_{tau<L} sum_{d<D} bit(g+tau, d) * LF[m][tau][d]          g = b*T + t
//
// Tensor path: the sliding window is folded into the N dimension.  One GEMM
//   C[g][(m,tau)] = sum_d bit(g,d) * LF[m][tau][d]      (rows = every frame of every utterance)
// runs on tcgen05 with the frame bits expanded to bf16 {0,1} in shared memory (A, exact) and the
// filters as bf16 hi+lo planes accumulated into the same TMEM columns (B, bulk-copied tiles in
// UMMA core-matrix order); the epilogue sums the diagonals  scores[g][m] = sum_tau C[g+tau][(m,tau)]
// in fp32 through a small shared-memory staging tile, carrying the L-1 unfinished rows from one
// 384-row step to the next.  Every frame row is therefore expanded and multiplied exactly once
// per template group: the "sliding reuse" is total.
//
// CTA (persistent, 1 per SM, 576 threads):
//   warps 0-3   epilogue: TMEM -> smem staging -> diagonal sums -> global
//   warp  4     MMA issuer + TMEM allocation
//   warp  5     filter-tile loader (cp.async.bulk, one 2*NT*128-byte copy per 64-feature chunk)
//   warps 6-17  frame-bit expanders (one row of the 384-row step each)
#include "common.cuh"

namespace tsr {

#define DT_MSUB_MAX 3
#define DT_NSTAGE 2
#define DT_THREADS 576
#define DT_PRODUCERS 384
#define DT_TAU_CHUNK 40    // diagonal sums are staged through smem in chunks of <= 40 filter frames
#define DT_HDR_BYTES 128   // blob header: [0] plane format, [1] max |w| bits

struct FilterLayout {
    int M, L, D, rowwords, nchunks;
    int tpt;      // templates per N tile
    int NT;       // N tile = roundup16(tpt*L)
    int ntiles;   // ceil(M / tpt)
    int msub;     // 128-row accumulators per step (msub*NT <= 512)
    size_t tile_bytes;   // one (ntile, chunk): 2 planes * NT rows * 128 B
    size_t total;
};

__host__ __device__ inline FilterLayout filter_layout(int M, int L, int D) {
    FilterLayout f;
    f.M = M; f.L = L; f.D = D;
    f.rowwords = ((D + 63) / 64) * 2;
    f.nchunks = f.rowwords / 2;
    int tpt = 160 / L;
    if (tpt < 1) tpt = 1;
    if (tpt > M) tpt = M;
    f.tpt = tpt;
    f.NT = (tpt * L + 15) & ~15;
    f.ntiles = (M + tpt - 1) / tpt;
    int msub = 512 / f.NT;
    f.msub = msub > DT_MSUB_MAX ? DT_MSUB_MAX : msub;
    f.tile_bytes = (size_t)2 * f.NT * 128;
    f.total = DT_HDR_BYTES + (size_t)f.ntiles * f.nchunks * f.tile_bytes;
    return f;
}

__global__ void filters_maxabs_kernel(const float* __restrict__ LF, size_t n, uint8_t* __restrict__ blob) {
    uint32_t* hdr = reinterpret_cast<uint32_t*>(blob);
    uint32_t m = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        m = max(m, __float_as_uint(fabsf(LF[i])));
    for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(&hdr[1], m);
}
__global__ void filters_publish_format_kernel(uint8_t* __restrict__ blob, int plane_format) {
    uint32_t* hdr = reinterpret_cast<uint32_t*>(blob);
    hdr[0] = pick_format(hdr[1], plane_format);
}

// One thread per 16-bit element of a plane.  Element (ntile, chunk, plane, kgroup, row n, e).
__global__ void prepare_filters_kernel(const float* __restrict__ LF, uint8_t* __restrict__ blob, FilterLayout f,
                                       int plane_format) {
    const uint32_t fmt = pick_format(reinterpret_cast<const uint32_t*>(blob)[1], plane_format);
    blob += DT_HDR_BYTES;
    const size_t nelem = (size_t)f.ntiles * f.nchunks * f.NT * 64;   // per plane
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nelem; i += (size_t)gridDim.x * blockDim.x) {
        int e = (int)(i & 7);
        size_t r = i >> 3;
        int n = (int)(r % f.NT); r /= f.NT;
        int kg = (int)(r & 7); r >>= 3;
        int c = (int)(r % f.nchunks);
        int nt = (int)(r / f.nchunks);
        int d = c * 64 + kg * 8 + e;
        int ml = n / f.L, tau = n - ml * f.L;
        int m = nt * f.tpt + ml;
        float w = 0.f;
        if (ml < f.tpt && m < f.M && d < f.D) w = LF[((size_t)m * f.L + tau) * f.D + d];
        uint16_t hi, lo;
        split_hi_lo((double)w, fmt, hi, lo);
        size_t base = ((size_t)nt * f.nchunks + c) * f.tile_bytes;
        size_t o = ((size_t)kg * f.NT + n) * 16 + e * 2;
        *reinterpret_cast<uint16_t*>(blob + base + o) = hi;
        *reinterpret_cast<uint16_t*>(blob + base + (size_t)f.NT * 128 + o) = lo;
    }
}

struct DetectArgs {
    const uint32_t* bits;   // [R][rowwords]
    const uint8_t* blob;
    float* scores;          // [R][M]
    long long R;            // B*T
    int T;
    int rows_per_unit;      // owned output rows per unit
    int units_per_ntile;
    int steps_per_unit;
    FilterLayout f;
};

struct DetectSmem {
    uint32_t a, b, stage, outacc, bars, stage_bytes, a_bytes, total;
};
__host__ __device__ inline DetectSmem detect_smem_layout(const FilterLayout& f) {
    DetectSmem S;
    S.a_bytes = f.msub * 16384;
    S.stage_bytes = S.a_bytes + (uint32_t)f.tile_bytes;
    S.a = 0;
    S.b = S.a_bytes;                                   // within a stage
    S.stage = DT_NSTAGE * S.stage_bytes;               // epilogue staging [128][L+1] floats
    S.outacc = S.stage + ((128 * ((f.L < DT_TAU_CHUNK ? f.L : DT_TAU_CHUNK) + 1) * 4 + 127) & ~127u);
    S.bars = S.outacc + (((f.msub * 128 + f.L) * f.tpt * 4 + 127) & ~127u);
    S.total = S.bars + 128;
    return S;
}

__global__ void __launch_bounds__(DT_THREADS, 1) detect_tc_kernel(DetectArgs A) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const FilterLayout& f = A.f;
    const DetectSmem S = detect_smem_layout(f);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S.bars);   // [DT_NSTAGE]
    uint64_t* empty_bar = full_bar + DT_NSTAGE;                        // [DT_NSTAGE]
    uint64_t* tfull_bar = empty_bar + DT_NSTAGE;                       // [1]
    uint64_t* tempty_bar = tfull_bar + 1;                              // [DT_MSUB_MAX]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + DT_MSUB_MAX);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t fmt = reinterpret_cast<const uint32_t*>(A.blob)[0];
    const int L = f.L, NT = f.NT, msub = f.msub, nchunks = f.nchunks, tpt = f.tpt;
    const int step_rows = msub * 128;
    const long long nunits = (long long)f.ntiles * A.units_per_ntile;

    if (threadIdx.x == 0) {
        for (int s = 0; s < DT_NSTAGE; ++s) { mbar_init(&full_bar[s], DT_PRODUCERS + 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tfull_bar, 1);
        for (int s = 0; s < DT_MSUB_MAX; ++s) mbar_init(&tempty_bar[s], 128);
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp >= 6) {
        // ============================ frame-bit expanders ============================
        const int pr = threadIdx.x - 192;             // row within the step
        const int sub = pr >> 7, r = pr & 127;
        const uint32_t one = fmt_one(fmt);
        uint32_t it = 0;                               // global chunk counter (pipeline position)
        for (long long u = blockIdx.x; u < nunits; u += gridDim.x) {
            const long long ur = u / f.ntiles;
            const long long o0 = ur * A.rows_per_unit;
            long long o1 = o0 + A.rows_per_unit; if (o1 > A.R) o1 = A.R;
            long long in_end = o1 + L - 1; if (in_end > A.R) in_end = A.R;
            for (int step = 0; step < A.steps_per_unit; ++step) {
                const long long g = o0 + (long long)step * step_rows + pr;
                const bool live = pr < step_rows && g < in_end;
                const uint2* src = reinterpret_cast<const uint2*>(A.bits + (live ? g : 0) * f.rowwords);
                for (int c = 0; c < nchunks; ++c, ++it) {
                    const int stage = it % DT_NSTAGE;
                    const uint32_t ph = (it / DT_NSTAGE) & 1;
                    uint2 w = live ? __ldg(src + c) : make_uint2(0u, 0u);
                    mbar_wait(&empty_bar[stage], ph ^ 1);
                    if (pr < step_rows) {
                        uint8_t* dst = smem + stage * S.stage_bytes + sub * 16384 + r * 16;
#pragma unroll
                        for (int kg = 0; kg < 4; ++kg)
                            *reinterpret_cast<uint4*>(dst + kg * 2048) = expand_byte((w.x >> (8 * kg)) & 0xFF, one);
#pragma unroll
                        for (int kg = 0; kg < 4; ++kg)
                            *reinterpret_cast<uint4*>(dst + (kg + 4) * 2048) = expand_byte((w.y >> (8 * kg)) & 0xFF, one);
                    }
                    fence_proxy_async_smem();
                    mbar_arrive(&full_bar[stage]);
                }
            }
        }
    } else if (warp == 5) {
        // ============================ filter-tile loader ============================
        if (lane == 0) {
            uint32_t it = 0;
            for (long long u = blockIdx.x; u < nunits; u += gridDim.x) {
                const int nt = (int)(u % f.ntiles);
                const uint8_t* src = A.blob + DT_HDR_BYTES + (size_t)nt * nchunks * f.tile_bytes;
                for (int step = 0; step < A.steps_per_unit; ++step)
                    for (int c = 0; c < nchunks; ++c, ++it) {
                        const int stage = it % DT_NSTAGE;
                        const uint32_t ph = (it / DT_NSTAGE) & 1;
                        mbar_wait(&empty_bar[stage], ph ^ 1);
                        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)f.tile_bytes);
                        bulk_g2s(smem + stage * S.stage_bytes + S.b, src + (size_t)c * f.tile_bytes, (uint32_t)f.tile_bytes,
                                 &full_bar[stage]);
                    }
            }
        }
    } else if (warp == 4) {
        // ================================ MMA issuer ================================
        const uint32_t idesc = umma_idesc_f32acc(fmt, 128, NT);
        uint32_t it = 0, stepno = 0;
        for (long long u = blockIdx.x; u < nunits; u += gridDim.x)
            for (int step = 0; step < A.steps_per_unit; ++step, ++stepno) {
                for (int c = 0; c < nchunks; ++c, ++it) {
                    const int stage = it % DT_NSTAGE;
                    const uint32_t ph = (it / DT_NSTAGE) & 1;
                    mbar_wait(&full_bar[stage], ph);
                    tc_fence_after();
                    const uint32_t abase = smem_u32(smem + stage * S.stage_bytes);
                    const uint32_t bbase = abase + S.b;
                    for (int sub = 0; sub < msub; ++sub) {
                        if (c == 0) {
                            mbar_wait(&tempty_bar[sub], (stepno & 1) ^ 1);
                            tc_fence_after();
                        }
                        if (lane == 0) {
#pragma unroll
                            for (int plane = 0; plane < 2; ++plane)
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) {
                                    uint64_t adesc = umma_desc_kmajor_noswz(abase + sub * 16384 + ks * 4096, 2048, 128);
                                    uint64_t bdesc = umma_desc_kmajor_noswz(bbase + plane * NT * 128 + ks * 2 * NT * 16, NT * 16, 128);
                                    umma_f16(tmem_base + sub * NT, adesc, bdesc, idesc, !(c == 0 && plane == 0 && ks == 0));
                                }
                        }
                        __syncwarp();
                    }
                    if (lane == 0) umma_commit(&empty_bar[stage]);
                    __syncwarp();
                }
                if (lane == 0) umma_commit(tfull_bar);
                __syncwarp();
            }
    } else {
        // ================================= epilogue =================================
        float* stg = reinterpret_cast<float*>(smem + S.stage);       // [128][TC+1]
        float* outacc = reinterpret_cast<float*>(smem + S.outacc);   // [step_rows + L - 1][tpt]
        const int e = threadIdx.x;                                    // 0..127 == TMEM lane
        const int TC = L < DT_TAU_CHUNK ? L : DT_TAU_CHUNK;
        const int pitch = TC + 1;
        const int nacc = (step_rows + L - 1) * tpt;
        uint32_t stepno = 0;
        for (long long u = blockIdx.x; u < nunits; u += gridDim.x) {
            const int nt = (int)(u % f.ntiles);
            const long long ur = u / f.ntiles;
            const long long o0 = ur * A.rows_per_unit;
            long long o1 = o0 + A.rows_per_unit; if (o1 > A.R) o1 = A.R;
            for (int i = e; i < nacc; i += 128) outacc[i] = 0.f;
            named_bar_sync(1, 128);
            for (int step = 0; step < A.steps_per_unit; ++step, ++stepno) {
                mbar_wait(tfull_bar, stepno & 1);
                tc_fence_after();
                for (int sub = 0; sub < msub; ++sub) {
                    const uint32_t taddr = tmem_base + sub * NT + ((uint32_t)(warp * 32) << 16);
                    for (int ml = 0; ml < tpt; ++ml) {
                        for (int c0 = 0; c0 < L; c0 += TC) {          // chunk of filter frames [c0, c0+cl)
                            const int cl = (L - c0) < TC ? (L - c0) : TC;
                            for (int t0 = 0; t0 < cl; t0 += 8) {
                                uint32_t v[8];
                                tmem_ld8(taddr + ml * L + c0 + t0, v);
                                tmem_ld_wait();
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    if (t0 + i < cl) stg[e * pitch + t0 + i] = __uint_as_float(v[i]);
                            }
                            if (ml == tpt - 1 && c0 + cl >= L) {   // accumulator drained: hand it back to the MMA warp
                                tc_fence_before();
                                mbar_arrive(&tempty_bar[sub]);
                            }
                            named_bar_sync(1, 128);
                            // out index q <-> row r0 = q - (L-1) at tau = 0; this chunk adds tau in [c0, c0+cl)
                            for (int q = e; q < 128 + L - 1; q += 128) {
                                const int r0 = q - (L - 1);
                                int tlo = -r0 > c0 ? -r0 : c0;
                                int thi = (128 - r0) < (c0 + cl) ? (128 - r0) : (c0 + cl);
                                if (tlo < thi) {
                                    float part = 0.f;
                                    for (int tau = tlo; tau < thi; ++tau) part += stg[(r0 + tau) * pitch + (tau - c0)];
                                    outacc[(sub * 128 + q) * tpt + ml] += part;
                                }
                            }
                            named_bar_sync(1, 128);
                        }
                    }
                }
                // rows [0, step_rows) of outacc are complete: out row g = step_row0 - (L-1) + oi
                const long long gbase = o0 + (long long)step * step_rows - (L - 1);
                for (int i = e; i < step_rows * tpt; i += 128) {
                    const int oi = i / tpt, ml = i - oi * tpt;
                    const long long g = gbase + oi;
                    const int m = nt * tpt + ml;
                    if (g >= o0 && g < o1 && m < f.M) {
                        const int tl = (int)(g % A.T);
                        A.scores[g * f.M + m] = (tl <= A.T - L) ? outacc[i] : 0.f;
                    }
                }
                named_bar_sync(1, 128);
                // carry the L-1 unfinished rows to the front, clear the rest (read phase, barrier,
                // write phase; (L-1)*tpt < 256 so a thread holds at most two values)
                {
                    const int keep = (L - 1) * tpt;
                    float tmp[2] = {0.f, 0.f};
                    int cnt = 0;
                    for (int i = e; i < keep; i += 128) tmp[cnt++] = outacc[i + step_rows * tpt];
                    named_bar_sync(1, 128);
                    cnt = 0;
                    for (int i = e; i < keep; i += 128) outacc[i] = tmp[cnt++];
                    for (int i = keep + e; i < nacc; i += 128) outacc[i] = 0.f;
                    named_bar_sync(1, 128);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------ SIMT
// One thread per (g, m): walks the set bits of the L rows in (tau, d) order and adds fp32 filter
// values sequentially -- the reference's exact arithmetic (detect, :162-175).  Cross-check kernel.
__global__ void detect_simt_kernel(const uint32_t* __restrict__ bits, const float* __restrict__ LF,
                                   float* __restrict__ scores, long long R, int T, int D, int rowwords, int M, int L) {
    const long long total = R * M;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long g = i / M;
        const int m = (int)(i - g * M);
        const int tl = (int)(g % T);
        float acc = 0.f;
        if (tl <= T - L) {
            for (int tau = 0; tau < L; ++tau) {
                const uint32_t* row = bits + (g + tau) * rowwords;
                const float* w = LF + ((size_t)m * L + tau) * D;
                for (int wi = 0; wi < rowwords; ++wi) {
                    uint32_t x = row[wi];
                    while (x) {
                        int bpos = __ffs(x) - 1;
                        x &= x - 1;
                        acc = __fadd_rn(acc, w[wi * 32 + bpos]);
                    }
                }
            }
        }
        scores[i] = acc;
    }
}

// Real-valued features: separate multiply and add in (tau, d) order (detect_float :206-216).
__global__ void detect_float_kernel(const float* __restrict__ Fm, const float* __restrict__ LF,
                                    const float* __restrict__ init, float* __restrict__ scores, long long R, int T,
                                    int D, int M, int L) {
    const long long total = R * M;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long g = i / M;
        const int m = (int)(i - g * M);
        const int tl = (int)(g % T);
        float acc = 0.f;
        if (tl <= T - L) {
            if (init) acc = init[m];          // detect_float_max starts at cs[m] (:258)
            for (int tau = 0; tau < L; ++tau) {
                const float* x = Fm + (g + tau) * D;
                const float* w = LF + ((size_t)m * L + tau) * D;
                for (int d = 0; d < D; ++d) acc = __fadd_rn(acc, __fmul_rn(w[d], x[d]));
            }
        }
        scores[i] = acc;
    }
}

// best/arg over templates, first maximum wins (strict >, detect_float_max :268-277).
__global__ void detect_reduce_kernel(const float* __restrict__ scores, const float* __restrict__ cs, long long R, int T,
                                     int M, int L, int tail_mode, float* __restrict__ best, uint16_t* __restrict__ arg) {
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < R; g += (long long)gridDim.x * blockDim.x) {
        const int tl = (int)(g % T);
        const bool valid = tl <= T - L;
        float bv = 0.f;
        int bm = 0;
        if (valid || tail_mode == 1) {
            for (int m = 0; m < M; ++m) {
                float v = (valid ? scores[g * M + m] : 0.f) + (cs ? cs[m] : 0.f);
                if (m == 0 || v > bv) { bv = v; bm = m; }
            }
        }
        if (best) best[g] = bv;
        if (arg) arg[g] = (uint16_t)bm;
    }
}

}  // namespace tsr

using namespace tsr;

extern "C" {

size_t tsr_filter_blob_bytes(int M, int L, int D) {
    if (M < 1 || L < 1 || L > 256 || D < 1) return 0;
    return filter_layout(M, L, D).total;
}

int tsr_prepare_filters(const float* LF, int M, int L, int D, int plane_format, void* blob, void* stream) {
    TSR_REQUIRE(LF && blob, "tsr_prepare_filters: null pointer");
    TSR_REQUIRE(M >= 1 && M <= 65535, "tsr_prepare_filters: need 1 <= M <= 65535 templates (uint16 argmax), got %d", M);
    TSR_REQUIRE(L >= 1 && L <= 256, "tsr_prepare_filters: template length %d outside [1,256]", L);
    TSR_REQUIRE(D >= 1, "tsr_prepare_filters: D < 1");
    int rc = require_sm100("tsr_prepare_filters");
    if (rc) return rc;
    TSR_REQUIRE(plane_format >= 0 && plane_format <= 2, "tsr_prepare_filters: plane_format must be 0 (auto), 1 (bf16) or 2 (fp16)");
    FilterLayout f = filter_layout(M, L, D);
    cudaStream_t st = (cudaStream_t)stream;
    TSR_CUDA(cudaMemsetAsync(blob, 0, DT_HDR_BYTES, st));
    filters_maxabs_kernel<<<148 * 2, 256, 0, st>>>(LF, (size_t)M * L * D, (uint8_t*)blob);
    prepare_filters_kernel<<<148 * 8, 256, 0, st>>>(LF, (uint8_t*)blob, f, plane_format);
    filters_publish_format_kernel<<<1, 1, 0, st>>>((uint8_t*)blob, plane_format);
    return check_launch("tsr_prepare_filters");
}

int tsr_detect(const uint32_t* featbits, int B, int T, int D, int rowwords, const void* filter_blob,
               const float* LF_or_null, int M, int L, float* scores, int impl, void* stream) {
    if (B == 0) return TSR_OK;
    TSR_REQUIRE(featbits && scores, "tsr_detect: null pointer");
    TSR_REQUIRE(B >= 0 && T >= 1 && D >= 1 && M >= 1 && L >= 1, "tsr_detect: bad sizes");
    TSR_REQUIRE(rowwords == tsr_rowwords(D), "tsr_detect: rowwords must be tsr_rowwords(D)");
    TSR_REQUIRE(impl == TSR_IMPL_TENSOR || impl == TSR_IMPL_SIMT, "tsr_detect: unknown impl %d", impl);
    if (B == 0) return TSR_OK;
    int rc = require_sm100("tsr_detect");
    if (rc) return rc;
    const long long R = (long long)B * T;
    if (L > T) {   // no valid offset: all-zero output, like the reference's zero-initialised array
        TSR_CUDA(cudaMemsetAsync(scores, 0, sizeof(float) * R * M, (cudaStream_t)stream));
        return TSR_OK;
    }
    if (impl == TSR_IMPL_SIMT) {
        TSR_REQUIRE(LF_or_null, "tsr_detect(simt): needs the fp32 filters");
        long long blocks = (R * M + 127) / 128;
        if (blocks > 148LL * 32) blocks = 148LL * 32;
        detect_simt_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(featbits, LF_or_null, scores, R, T, D, rowwords, M, L);
        return check_launch("tsr_detect(simt)");
    }
    TSR_REQUIRE(filter_blob, "tsr_detect: needs the prepared filter blob");
    TSR_REQUIRE(L <= 256, "tsr_detect: template length %d > 256 unsupported on the tensor path", L);
    DetectArgs A;
    A.bits = featbits; A.blob = (const uint8_t*)filter_blob; A.scores = scores; A.R = R; A.T = T;
    A.f = filter_layout(M, L, D);
    DetectSmem S = detect_smem_layout(A.f);
    if (S.total > 227 * 1024) {
        set_error("tsr_detect: L=%d needs %u B of shared memory (> 227 KB)", L, S.total);
        return TSR_ERR_UNSUPPORTED;
    }
    // Partition the rows into units so that (ntiles * units) CTAs-worth of work balance over the SMs.
    const int nsm = sm_count();
    const int step_rows = A.f.msub * 128;
    long long best_cost = -1;
    int best_U = 1;
    const long long maxU = (R + step_rows - 1) / step_rows;
    for (long long U = 1; U <= maxU && U <= 8192; ++U) {
        long long rows = (R + U - 1) / U;
        long long steps = (rows + L - 1 + step_rows - 1) / step_rows;
        long long waves = ((long long)A.f.ntiles * U + nsm - 1) / nsm;
        long long cost = waves * steps;
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_U = (int)U; }
    }
    A.units_per_ntile = best_U;
    A.rows_per_unit = (int)((R + best_U - 1) / best_U);
    A.steps_per_unit = (A.rows_per_unit + L - 1 + step_rows - 1) / step_rows;
    const long long nunits = (long long)A.f.ntiles * best_U;
    TSR_CUDA(cudaFuncSetAttribute(detect_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.total));
    int blocks = (int)(nunits < nsm ? nunits : nsm);
    detect_tc_kernel<<<blocks, DT_THREADS, S.total, (cudaStream_t)stream>>>(A);
    return check_launch("tsr_detect");
}

int tsr_detect_reduce(const float* scores, const float* cs, int B, int T, int M, int L, int tail_mode, float* best,
                      uint16_t* arg, void* stream) {
    if (B == 0) return TSR_OK;
    TSR_REQUIRE(scores && (best || arg), "tsr_detect_reduce: null pointer");
    TSR_REQUIRE(B >= 0 && T >= 1 && M >= 1 && M <= 65535 && L >= 1, "tsr_detect_reduce: bad sizes");
    TSR_REQUIRE(tail_mode == 0 || tail_mode == 1, "tsr_detect_reduce: tail_mode must be 0 or 1");
    if (B == 0) return TSR_OK;
    int rc = require_sm100("tsr_detect_reduce");
    if (rc) return rc;
    const long long R = (long long)B * T;
    long long blocks = (R + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    detect_reduce_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(scores, cs, R, T, M, L, tail_mode, best, arg);
    return check_launch("tsr_detect_reduce");
}

int tsr_detect_float(const float* Fm, int B, int T, int D, const float* LF, int M, int L, const float* init_or_null,
                     float* scores, void* stream) {
    if (B == 0) return TSR_OK;
    TSR_REQUIRE(Fm && LF && scores, "tsr_detect_float: null pointer");
    TSR_REQUIRE(B >= 0 && T >= 1 && D >= 1 && M >= 1 && L >= 1, "tsr_detect_float: bad sizes");
    if (B == 0) return TSR_OK;
    int rc = require_sm100("tsr_detect_float");
    if (rc) return rc;
    const long long R = (long long)B * T;
    long long blocks = (R * M + 127) / 128;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    detect_float_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(Fm, LF, init_or_null, scores, R, T, D, M, L);
    return check_launch("tsr_detect_float");
}

}  // extern "C"
