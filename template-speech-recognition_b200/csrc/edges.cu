// edges.cu -- the edge front end that feeds K1a:
//   N4  8-orientation edge magnitudes from the (smoothed, log) spectrogram:
//       edge_signal_proc._edge_map_no_threshold (edge_signal_proc.py:1697-1836);
//   N3  per-(orientation, 40-frame block) quantile thresholding with an absolute floor:
//       threshold_edge_block / threshold_edgemap / _edge_map_threshold_segments
//       (edge_signal_proc.py:1184-1307, 1537-1675), writing the packed edge bytes K1a consumes.
// All arithmetic is float64 with the reference's operations (one subtraction per difference,
// comparisons, an exact order statistic), so the results are bit-exact.  HBM/latency-bound byte work:
// the fused entry (tsr_edges_from_spectrogram) reads 8 B per spectrogram cell and writes 1 B per cell;
// the 8 magnitude planes (64 B per cell) are recomputed from a shared-memory tile, never stored.
#include "common.cuh"

namespace tsr {

// ------------------------------------------------------------------------------ N4
// Spectrogram s[r*ld + c] (frequency rows x time), c = column of absolute frame t in the addressed tile.
// Orientation pair p = e/2 has the difference field D and two neighbours along its direction;
// E = +-D where D is a local extremum with the right sign, on the sub-rectangle the reference writes
// (everything else stays 0).  e odd = the negative member of the pair.
__device__ __forceinline__ double edge_mag(const double* __restrict__ s, int ld, int F, int T, int r, int t, int c, int e) {
    const int pair = e >> 1;
    const bool tin = t >= 1 && t <= T - 3, rin = r >= 1 && r <= F - 3;
    double d, l, g;
    if (pair == 0) {            // time differences, E[:, 1:-2]  (edge_signal_proc.py:1744-1761)
        if (!tin) return 0.0;
        const double* p = s + (long long)r * ld + c;
        d = p[1] - p[0]; l = p[0] - p[-1]; g = p[2] - p[1];
    } else if (pair == 1) {     // frequency differences, rows 1 .. F-3  (:1764-1781)
        if (!rin) return 0.0;
        const double* p = s + (long long)r * ld + c;
        d = p[ld] - p[0]; l = p[0] - p[-ld]; g = p[2 * ld] - p[ld];
    } else if (pair == 2) {     // SW-NE: D[r][c] = S[r+1][c+1] - S[r][c], neighbours (r-1,c-1), (r+1,c+1)  (:1784-1801)
        if (!(tin && rin)) return 0.0;
        const double* p = s + (long long)r * ld + c;
        d = p[ld + 1] - p[0]; l = p[0] - p[-ld - 1]; g = p[2 * ld + 2] - p[ld + 1];
    } else {                    // SE-NW: D[r][c] = S[r][c+1] - S[r+1][c], neighbours (r+1,c-1), (r-1,c+1)  (:1804-1821)
        if (!(tin && rin)) return 0.0;
        const double* p = s + (long long)r * ld + c;
        d = p[1] - p[ld]; l = p[ld] - p[2 * ld - 1]; g = p[-ld + 2] - p[1];
    }
    return (e & 1) ? (-d) * (double)((d <= l) && (d <= g) && (d <= 0.0)) : d * (double)((d >= l) && (d >= g) && (d >= 0.0));
}

// S double [B][F][T]; mags double [B][8][F][T] (the reference's (8H, time) array per utterance).
__global__ void __launch_bounds__(256) edge_map_kernel(const double* __restrict__ S, double* __restrict__ E, long long n,
                                                       int F, int T) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(idx % T);
        long long q = idx / T;
        const int r = (int)(q % F);
        q /= F;
        const int e = (int)(q & 7);
        const long long b = q >> 3;
        E[idx] = edge_mag(S + b * (long long)F * T, T, F, T, r, t, t, e);
    }
}

// ------------------------------------------------------------------------------ N3
// One CTA (8 warps) per (utterance, time block).
//   1. all threads: one evaluation per (cell, orientation pair) gives the positive and the negative member's
//      magnitude; positive magnitudes are appended (as IEEE bit patterns, which order like the positive
//      doubles) to the orientation's candidate list in shared memory (capacity `cap`, count kept exactly);
//   2. warp e owns orientation e: exact radix select (8 passes of 8 bits, private 256-bin histogram, warp-level
//      synchronisation only) of the candidate of rank int(quantile * n) -- over the list, or, when the list
//      overflowed, over re-evaluated magnitudes; threshold = max(tau, abst[e]);
//   3. all threads: packed edge bytes, bit e = magnitude positive and >= threshold.
// The ragged last block is thresholded on the window of the final `block_len` frames
// (edge_signal_proc.py:1595-1633) and written both to its own columns of `bits` and, as a whole window, to
// `tail_bits` (the reference spreads that window before keeping its last columns).
// x / d for 0 <= x < 2^31 through a multiply-high (magic = floor((2^32 - 1) / d) + 1; 0 stands for d == 1): the per-cell
// index arithmetic of the kernel below was a fifth of its instructions as runtime integer divisions
__host__ inline uint32_t div_magic(int d) { return d > 1 ? (uint32_t)(0xFFFFFFFFull / (unsigned)d) + 1u : 0u; }
__device__ __forceinline__ int div_fast(int x, int d, uint32_t magic, int& rem) {
    if (magic == 0u) { rem = 0; return x; }
    int q = (int)__umulhi((uint32_t)x, magic);
    rem = x - q * d;
    if (rem < 0) { rem += d; --q; }
    return q;
}

struct ThresholdArgs {
    const double* src;    // FUSED: spectrogram [B][F][T]; else magnitudes [B][8][F][T]
    uint8_t* bits;        // [B][T][F]
    uint8_t* tail_bits;   // [B][tail_window][F] or null
    double* taus;         // [B][nblocks][8] quantile values (report_level), or null
    int F, T, block_len, nfull, tail_len, tail_window, cap;
    uint32_t m_F, m_len[2], m_ld[2];   // division magics: F; window length / tile pitch of a whole block [0] and of the tail window [1]
    double quantile;
    double abst[8];
};

// magnitudes of orientations 2*pair (pos) and 2*pair+1 (neg) at one cell: the two share the difference field
__device__ __forceinline__ void edge_mag_pair(const double* __restrict__ s, int ld, int F, int T, int r, int t, int c, int pair,
                                              double& pos, double& neg) {
    pos = 0.0; neg = 0.0;
    const bool tin = t >= 1 && t <= T - 3, rin = r >= 1 && r <= F - 3;
    const double* p = s + (long long)r * ld + c;
    double d, l, g;
    if (pair == 0) {
        if (!tin) return;
        d = p[1] - p[0]; l = p[0] - p[-1]; g = p[2] - p[1];
    } else if (pair == 1) {
        if (!rin) return;
        d = p[ld] - p[0]; l = p[0] - p[-ld]; g = p[2 * ld] - p[ld];
    } else if (pair == 2) {
        if (!(tin && rin)) return;
        d = p[ld + 1] - p[0]; l = p[0] - p[-ld - 1]; g = p[2 * ld + 2] - p[ld + 1];
    } else {
        if (!(tin && rin)) return;
        d = p[1] - p[ld]; l = p[ld] - p[2 * ld - 1]; g = p[-ld + 2] - p[1];
    }
    pos = d * (double)((d >= l) && (d >= g) && (d >= 0.0));
    neg = (-d) * (double)((d <= l) && (d <= g) && (d <= 0.0));
}

template <bool FUSED>
__global__ void __launch_bounds__(256) threshold_blocks_kernel(ThresholdArgs A) {
    extern __shared__ unsigned long long s_dyn[];      // [8][cap] candidate keys, then (FUSED) the F x (len+3) tile
    __shared__ unsigned int hist[8][256];
    __shared__ unsigned int s_cnt[8];
    __shared__ double s_thr[8];
    const int b = blockIdx.y, blk = blockIdx.x;
    const bool tail = blk >= A.nfull;
    const int t0 = tail ? A.T - A.tail_window : blk * A.block_len;
    const int len = tail ? A.tail_window : A.block_len;
    const int nel = A.F * len, F = A.F, T = A.T, cap = A.cap;
    const int ld = len + 3;
    const uint32_t m_len = A.m_len[tail ? 1 : 0], m_ld = A.m_ld[tail ? 1 : 0];
    unsigned long long* keys = s_dyn;
    double* s_tile = reinterpret_cast<double*>(s_dyn + 8 * (size_t)cap);   // column 0 = frame t0 - 1
    const double* base = A.src + (long long)b * (FUSED ? 1 : 8) * F * T;
    if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
    if (FUSED) {
        for (int i = threadIdx.x; i < F * ld; i += blockDim.x) {
            int c;
            const int r = div_fast(i, ld, m_ld, c), t = t0 - 1 + c;
            s_tile[i] = (t >= 0 && t < T) ? base[(long long)r * T + t] : 0.0;
        }
    }
    __syncthreads();
    auto mag_pair = [&](int pair, int r, int tl, double& pos, double& neg) {    // tl = frame inside the window
        if (FUSED) {
            edge_mag_pair(s_tile, ld, F, T, r, t0 + tl, tl + 1, pair, pos, neg);
        } else {
            const double* m = base + ((long long)(2 * pair) * F + r) * T + t0 + tl;
            pos = m[0]; neg = m[(long long)F * T];
        }
    };
    // ---- 1. candidate lists (cell outside, the four pairs unrolled inside: one index split per cell, and every copy of the
    // evaluation specialised to its pair; the order of a list does not matter to the order statistic)
    for (int i = threadIdx.x; i < nel; i += blockDim.x) {
        int tl1;
        const int r = div_fast(i, len, m_len, tl1);
#pragma unroll
        for (int pair = 0; pair < 4; ++pair) {
            double pos, neg;
            mag_pair(pair, r, tl1, pos, neg);
            if (pos > 0.0) {
                const unsigned int slot = atomicAdd(&s_cnt[2 * pair], 1u);
                if (slot < (unsigned)cap) keys[(size_t)(2 * pair) * cap + slot] = (unsigned long long)__double_as_longlong(pos);
            }
            if (neg > 0.0) {
                const unsigned int slot = atomicAdd(&s_cnt[2 * pair + 1], 1u);
                if (slot < (unsigned)cap) keys[(size_t)(2 * pair + 1) * cap + slot] = (unsigned long long)__double_as_longlong(neg);
            }
        }
    }
    __syncthreads();
    // ---- 2. order statistic per orientation
    {
        const int e = threadIdx.x >> 5, lane = threadIdx.x & 31;
        unsigned int* h = hist[e];
        const unsigned int npos = s_cnt[e];
        const bool listed = npos <= (unsigned)cap;
        const unsigned long long* mykeys = keys + (size_t)e * cap;
        double tau = 0.0, thr = __longlong_as_double(0x7ff0000000000000ll);   // +inf: nothing passes an empty block
        if (npos) {
            // rank as the reference computes it: int(quantile_level * n) in float64 (:1293)
            unsigned int rank = (unsigned int)(long long)(A.quantile * (double)npos);
            if (rank >= npos) rank = npos - 1;          // quantile < 1 is checked on the host; guard anyway
            unsigned long long prefix = 0ull;
            for (int pass = 0; pass < 8; ++pass) {
                const int shift = 56 - 8 * pass;
                const unsigned long long himask = pass == 0 ? 0ull : ~0ull << (shift + 8);
#pragma unroll
                for (int k = 0; k < 8; ++k) h[lane * 8 + k] = 0;
                __syncwarp();
                if (listed) {
                    for (unsigned int i = lane; i < npos; i += 32) {
                        const unsigned long long u = mykeys[i];
                        if ((u & himask) == prefix) atomicAdd(&h[(u >> shift) & 255], 1u);
                    }
                } else {                                // list overflowed: walk the window again
                    for (int i = lane; i < nel; i += 32) {
                        const int r = i / len;
                        double pos, neg;
                        mag_pair(e >> 1, r, i - r * len, pos, neg);
                        const double v = (e & 1) ? neg : pos;
                        if (v > 0.0) {
                            const unsigned long long u = (unsigned long long)__double_as_longlong(v);
                            if ((u & himask) == prefix) atomicAdd(&h[(u >> shift) & 255], 1u);
                        }
                    }
                }
                __syncwarp();
                // lane owns bins 8*lane .. 8*lane+7: inclusive scan of the lane sums, then walk the winning lane's bins
                unsigned int mine[8], sum = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) { mine[k] = h[lane * 8 + k]; sum += mine[k]; }
                unsigned int incl = sum;
                for (int off = 1; off < 32; off <<= 1) {
                    const unsigned int o = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += o;
                }
                const unsigned int win = __ffs(__ballot_sync(0xffffffffu, rank < incl)) - 1;
                unsigned int rk = rank - (incl - sum), digit = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (digit == (unsigned)k && rk >= mine[k]) { rk -= mine[k]; digit = k + 1; }
                }
                digit = __shfl_sync(0xffffffffu, lane * 8 + digit, win);
                rank = __shfl_sync(0xffffffffu, rk, win);
                prefix |= (unsigned long long)digit << shift;
                __syncwarp();
            }
            tau = __longlong_as_double((long long)prefix);
            thr = tau > A.abst[e] ? tau : A.abst[e];
        }
        if (lane == 0) {
            s_thr[e] = thr;
            if (A.taus) A.taus[((long long)b * gridDim.x + blk) * 8 + e] = tau;
        }
    }
    __syncthreads();
    // ---- 3. packed bytes: bit e = magnitude positive and >= threshold (:1286, :1297)
    for (int i = threadIdx.x; i < nel; i += blockDim.x) {
        int r;
        const int tl = div_fast(i, F, A.m_F, r);        // consecutive threads = consecutive frequency bins of a frame
        const int t = t0 + tl;
        uint32_t byte = 0;
#pragma unroll
        for (int pair = 0; pair < 4; ++pair) {
            double pos, neg;
            mag_pair(pair, r, tl, pos, neg);
            byte |= (uint32_t)(pos > 0.0 && pos >= s_thr[2 * pair]) << (2 * pair);
            byte |= (uint32_t)(neg > 0.0 && neg >= s_thr[2 * pair + 1]) << (2 * pair + 1);
        }
        if (!tail) {
            A.bits[((long long)b * T + t) * F + r] = (uint8_t)byte;
        } else {
            if (t >= T - A.tail_len) A.bits[((long long)b * T + t) * F + r] = (uint8_t)byte;
            if (A.tail_bits) A.tail_bits[((long long)b * A.tail_window + tl) * F + r] = (uint8_t)byte;
        }
    }
}

static int launch_threshold(bool fused, const char* what, const double* src, uint8_t* bits, uint8_t* tail_bits, double* taus,
                            int B, int F, int T, int block_len, double quantile, const double* host_abst, void* stream) {
    if (B == 0) return TSR_OK;
    TSR_REQUIRE(src && bits && host_abst, "%s: null pointer", what);
    TSR_REQUIRE(B > 0 && F >= 1 && T >= 1 && block_len >= 1, "%s: bad sizes", what);
    TSR_REQUIRE(!fused || (F >= 4 && T >= 4), "%s: need at least 4 frequency rows and 4 frames (got %d x %d)", what, F, T);
    TSR_REQUIRE(quantile >= 0.0 && quantile < 1.0, "%s: quantile must lie in [0, 1) (the reference indexes sorted[int(q*n)])",
                what);
    TSR_REQUIRE(B <= 65535, "%s: B > 65535; split the batch", what);
    int rc = require_sm100(what);
    if (rc) return rc;
    ThresholdArgs A;
    A.src = src; A.bits = bits; A.tail_bits = tail_bits; A.taus = taus;
    A.F = F; A.T = T; A.block_len = block_len;
    A.nfull = T / block_len;
    A.tail_len = T % block_len;
    A.tail_window = T < block_len ? T : block_len;
    A.quantile = quantile;
    for (int e = 0; e < 8; ++e) A.abst[e] = host_abst[e];
    A.m_F = div_magic(F);
    A.m_len[0] = div_magic(block_len); A.m_ld[0] = div_magic(block_len + 3);
    A.m_len[1] = div_magic(A.tail_window); A.m_ld[1] = div_magic(A.tail_window + 3);
    dim3 grid(A.nfull + (A.tail_len ? 1 : 0), B);
    const int nel = F * A.tail_window;
    A.cap = nel < 640 ? nel : 640;                       // candidate list capacity per orientation (overflow re-evaluates)
    size_t smem = 8 * (size_t)A.cap * sizeof(unsigned long long);
    if (fused) smem += (size_t)F * ((size_t)A.tail_window + 3) * sizeof(double);
    TSR_REQUIRE(smem <= 200 * 1024, "%s: F x (block_len + 3) spectrogram tile (%zu B) exceeds shared memory", what, smem);
    if (fused) {
        cudaFuncSetAttribute(threshold_blocks_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        threshold_blocks_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(A);
    } else {
        cudaFuncSetAttribute(threshold_blocks_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        threshold_blocks_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(A);
    }
    return check_launch(what);
}

}  // namespace tsr

using namespace tsr;

extern "C" {

int tsr_edge_map(const double* S, double* mags, int B, int F, int T, void* stream) {
    if (B == 0) return TSR_OK;
    TSR_REQUIRE(S && mags, "tsr_edge_map: null pointer");
    TSR_REQUIRE(B > 0 && F >= 4 && T >= 4, "tsr_edge_map: need at least 4 frequency rows and 4 frames (got %d x %d)", F, T);
    int rc = require_sm100("tsr_edge_map");
    if (rc) return rc;
    const long long n = (long long)B * 8 * F * T;
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)sm_count() * 32) blocks = (long long)sm_count() * 32;
    edge_map_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(S, mags, n, F, T);
    return check_launch("tsr_edge_map");
}

int tsr_threshold_edges(const double* mags, uint8_t* bits, uint8_t* tail_bits_or_null, double* taus_or_null, int B, int F,
                        int T, int block_len, double quantile, const double* host_abst_threshold, void* stream) {
    return launch_threshold(false, "tsr_threshold_edges", mags, bits, tail_bits_or_null, taus_or_null, B, F, T, block_len,
                            quantile, host_abst_threshold, stream);
}

int tsr_edges_from_spectrogram(const double* S, uint8_t* bits, uint8_t* tail_bits_or_null, double* taus_or_null, int B, int F,
                               int T, int block_len, double quantile, const double* host_abst_threshold, void* stream) {
    return launch_threshold(true, "tsr_edges_from_spectrogram", S, bits, tail_bits_or_null, taus_or_null, B, F, T, block_len,
                            quantile, host_abst_threshold, stream);
}

}  // extern "C"
