// spread.cu -- bit-packed max-filters: K1a edge spreading, K1b part-code spreading, part-graph
// spreading.  Integer / bit work, HBM-bound; every result is bit-exact by construction.
#include "common.cuh"

namespace tsr {

// ================================================================================= K1a
// out(t,f) = OR_i ( in(t + dt_i, f + df_i) & mask_i ), zero outside the utterance and outside
// the block of t.  The (dt, df, mask) list is built on the host from the per-orientation
// directions (edge_signal_proc.py:1406-1417) or the box ('all', :1623-1632).
#define TSR_MAX_SPREAD_TAPS 160
struct SpreadTaps {
    int n;
    int max_dt, max_df;
    int8_t dt[TSR_MAX_SPREAD_TAPS];
    int8_t df[TSR_MAX_SPREAD_TAPS];
    uint8_t mask[TSR_MAX_SPREAD_TAPS];
};

// CTA = (utterance b, tile of TT time rows).  The tile plus halo is staged in shared memory as
// rows of `pitch` bytes (4-aligned, zero padded left/right by padf >= max_df rounded to 4), so a
// thread producing one 32-bit output word (4 frequency bins) fetches each tap as an unaligned
// 4-byte window = two aligned LDS.32 + one funnel shift.
__global__ void __launch_bounds__(256) spread_edges_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                           int T, int F, int TT, int block_len, SpreadTaps taps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * TT;
    const int padf = (taps.max_df + 3) & ~3;
    const int Fw = (F + 3) & ~3;
    const int pitch = Fw + 2 * padf + 4;            // +4: the "hi" word of the last window
    const int rows = TT + 2 * taps.max_dt;
    const uint8_t* src = in + (size_t)b * T * F;

    // zero everything, then fill the in-range part
    for (int i = threadIdx.x * 4; i < rows * pitch; i += blockDim.x * 4) *reinterpret_cast<uint32_t*>(smem + i) = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < rows * F; i += blockDim.x) {
        int r = i / F, f = i - r * F;
        int t = t0 - taps.max_dt + r;
        if (t >= 0 && t < T) smem[r * pitch + padf + f] = src[(size_t)t * F + f];
    }
    __syncthreads();

    const int words_per_row = Fw >> 2;
    const bool aligned_out = (F & 3) == 0;
    for (int i = threadIdx.x; i < TT * words_per_row; i += blockDim.x) {
        int r = i / words_per_row, w = i - r * words_per_row;
        int t = t0 + r;
        if (t >= T) break;
        int blk_lo = 0, blk_hi = T;
        if (block_len > 0) {
            blk_lo = (t / block_len) * block_len;
            blk_hi = min(T, blk_lo + block_len);
        }
        uint32_t acc = 0;
        for (int k = 0; k < taps.n; ++k) {
            int ts = t + taps.dt[k];
            if (ts < blk_lo || ts >= blk_hi) continue;
            int byte_off = (r + taps.max_dt + taps.dt[k]) * pitch + padf + 4 * w + taps.df[k];
            int a = byte_off & ~3;
            uint32_t lo = *reinterpret_cast<const uint32_t*>(smem + a);
            uint32_t hi = *reinterpret_cast<const uint32_t*>(smem + a + 4);
            uint32_t v = __funnelshift_r(lo, hi, (byte_off & 3) * 8);
            acc |= v & (0x01010101u * taps.mask[k]);
        }
        uint8_t* dst = out + ((size_t)b * T + t) * F + 4 * w;
        if (aligned_out) {
            *reinterpret_cast<uint32_t*>(dst) = acc;
        } else {
            for (int q = 0; q < 4 && 4 * w + q < F; ++q) dst[q] = (uint8_t)(acc >> (8 * q));
        }
    }
}

// ================================================================================= K1b
// codes uint8 [B][To][Fo] -> featbits uint32 [B][To][rowwords], feature j*K + (p-1).
// CTA = (b, tile of RT output rows); three tile-wide smem phases, no atomics:
//   colset[r][j][:] = OR_{|x0-i|<=r0} onehot(code[x0][j])      (time direction, Kw words per cell)
//   wset[r][j][:]   = OR_{|x1-j|<=r1} colset[r][x1][:]          (frequency direction)
//   word w of packed row r gathers the K-bit sets overlapping features [32w, 32w+32).
#define TSR_SC_MAX_KW 8
__global__ void __launch_bounds__(256) spread_codes_kernel(const uint8_t* __restrict__ codes,
                                                           uint32_t* __restrict__ featbits, int To, int Fo, int K,
                                                           int r0, int r1, int rowwords, int RT) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int Kw = (K + 31) >> 5;
    const int b = blockIdx.y;
    const int i0 = blockIdx.x * RT;
    const int rows_in = RT + 2 * r0;
    const int nrow = min(RT, To - i0);
    uint8_t* scode = smem;                                                                 // [rows_in][Fo]
    uint32_t* colset = reinterpret_cast<uint32_t*>(smem + ((rows_in * Fo + 15) & ~15));    // [RT][Fo][Kw]
    uint32_t* wset = colset + RT * Fo * Kw;                                                // [RT][Fo][Kw]
    int2* wtab = reinterpret_cast<int2*>(wset + RT * Fo * Kw);                             // [rowwords] (j, o) of bit 32w

    for (int i = threadIdx.x; i < rows_in * Fo; i += blockDim.x) {
        int r = i / Fo, j = i - r * Fo;
        int x0 = i0 - r0 + r;
        scode[i] = (x0 >= 0 && x0 < To) ? codes[((size_t)b * To + x0) * Fo + j] : 0;
    }
    for (int w = threadIdx.x; w < rowwords; w += blockDim.x) {
        int d = 32 * w;
        wtab[w] = make_int2(d / K, d % K);
    }
    __syncthreads();

    // phase 1: one thread per (row, j) builds the Kw-word set of the 2*r0+1 codes above/below
    for (int q = threadIdx.x; q < nrow * Fo; q += blockDim.x) {
        int r = q / Fo, j = q - r * Fo;
        uint32_t s[TSR_SC_MAX_KW];
#pragma unroll
        for (int kw = 0; kw < TSR_SC_MAX_KW; ++kw) s[kw] = 0;
        for (int a = 0; a <= 2 * r0; ++a) {
            int p = scode[(r + a) * Fo + j];
            if (p > 0) {
                uint32_t bit = 1u << ((p - 1) & 31);
                int kwp = (p - 1) >> 5;
#pragma unroll
                for (int kw = 0; kw < TSR_SC_MAX_KW; ++kw)
                    if (kw == kwp) s[kw] |= bit;
            }
        }
#pragma unroll
        for (int kw = 0; kw < TSR_SC_MAX_KW; ++kw)
            if (kw < Kw) colset[q * Kw + kw] = s[kw];
    }
    __syncthreads();
    // phase 2: OR along frequency
    for (int q = threadIdx.x; q < nrow * Fo; q += blockDim.x) {
        int r = q / Fo, j = q - r * Fo;
        int lo = max(0, j - r1), hi = min(Fo - 1, j + r1);
        uint32_t s[TSR_SC_MAX_KW];
#pragma unroll
        for (int kw = 0; kw < TSR_SC_MAX_KW; ++kw) s[kw] = 0;
        for (int x1 = lo; x1 <= hi; ++x1) {
            const uint32_t* c = colset + (r * Fo + x1) * Kw;
#pragma unroll
            for (int kw = 0; kw < TSR_SC_MAX_KW; ++kw)
                if (kw < Kw) s[kw] |= c[kw];
        }
#pragma unroll
        for (int kw = 0; kw < TSR_SC_MAX_KW; ++kw)
            if (kw < Kw) wset[q * Kw + kw] = s[kw];
    }
    __syncthreads();
    // phase 3: gather into tightly packed rows, coalesced 32-bit stores
    const int D = Fo * K;
    for (int q = threadIdx.x; q < nrow * rowwords; q += blockDim.x) {
        int r = q / rowwords, w = q - r * rowwords;
        uint32_t word = 0;
        int d = 32 * w;
        const int dend = min(32 * w + 32, D);
        int2 jo = wtab[w];
        int j = jo.x, o = jo.y;
        while (d < dend) {
            int n = min(K - o, dend - d);
            const uint32_t* sp = wset + (r * Fo + j) * Kw;
            uint32_t lo32 = sp[o >> 5];
            uint32_t hi32 = ((o >> 5) + 1 < Kw) ? sp[(o >> 5) + 1] : 0u;
            uint32_t chunk = __funnelshift_r(lo32, hi32, o & 31);
            if (n < 32) chunk &= (1u << n) - 1u;
            word |= chunk << (d - 32 * w);
            d += n;
            o += n;
            if (o >= K) { o = 0; ++j; }
        }
        featbits[((size_t)b * To + i0 + r) * rowwords + w] = word;
    }
}

// ============================================================================ part graph
// In-place, per cell: for p1 ascending: if x[p1] then x |= graph[p1]  (code_parts.pyx:496-503;
// sequential, so bits switched on by a lower p1 propagate through higher p1 in the same pass).
// CTA loops over rows; the row lives in smem; each thread owns one cell's K-bit set.
#define TSR_GRAPH_MAX_KW 8
__global__ void __launch_bounds__(128) spread_graph_kernel(uint32_t* __restrict__ featbits,
                                                           const uint8_t* __restrict__ graph, long long R, int Fo,
                                                           int K, int rowwords) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int Kw = (K + 31) >> 5;
    uint32_t* grow = reinterpret_cast<uint32_t*>(smem);   // [K][Kw] adjacency bitsets
    uint32_t* srow = grow + K * Kw;                        // [rowwords + 1]
    uint32_t* orow = srow + rowwords + 1;                  // [rowwords + 1]
    for (int q = threadIdx.x; q < K * Kw; q += blockDim.x) {
        int p1 = q / Kw, kw = q - p1 * Kw;
        uint32_t s = 0;
        for (int bit = 0; bit < 32 && kw * 32 + bit < K; ++bit)
            if (graph[p1 * K + kw * 32 + bit]) s |= 1u << bit;
        grow[q] = s;
    }
    __syncthreads();
    for (long long r = blockIdx.x; r < R; r += gridDim.x) {
        uint32_t* row = featbits + r * rowwords;
        for (int w = threadIdx.x; w <= rowwords; w += blockDim.x) {
            srow[w] = w < rowwords ? row[w] : 0u;
            orow[w] = 0u;
        }
        __syncthreads();
        for (int j = threadIdx.x; j < Fo; j += blockDim.x) {
            uint32_t x[TSR_GRAPH_MAX_KW];
            const int d0 = j * K;
#pragma unroll
            for (int kw = 0; kw < TSR_GRAPH_MAX_KW; ++kw) {
                x[kw] = 0;
                if (kw < Kw) {
                    int d = d0 + 32 * kw;
                    uint32_t lo32 = srow[d >> 5];
                    uint32_t hi32 = ((d >> 5) + 1 <= rowwords) ? srow[(d >> 5) + 1] : 0u;
                    uint32_t v = __funnelshift_r(lo32, hi32, d & 31);
                    int n = K - 32 * kw;
                    if (n < 32) v &= (1u << n) - 1u;
                    x[kw] = v;
                }
            }
            for (int p1 = 0; p1 < K; ++p1) {
                uint32_t wsel = 0;
#pragma unroll
                for (int kw = 0; kw < TSR_GRAPH_MAX_KW; ++kw)
                    if (kw == (p1 >> 5)) wsel = x[kw];
                if ((wsel >> (p1 & 31)) & 1u) {
#pragma unroll
                    for (int kw = 0; kw < TSR_GRAPH_MAX_KW; ++kw)
                        if (kw < Kw) x[kw] |= grow[p1 * Kw + kw];
                }
            }
#pragma unroll
            for (int kw = 0; kw < TSR_GRAPH_MAX_KW; ++kw) {
                if (kw < Kw && x[kw]) {
                    int d = d0 + 32 * kw;
                    int sh = d & 31;
                    atomicOr(&orow[d >> 5], x[kw] << sh);
                    if (sh) atomicOr(&orow[(d >> 5) + 1], x[kw] >> (32 - sh));
                }
            }
        }
        __syncthreads();
        for (int w = threadIdx.x; w < rowwords; w += blockDim.x) row[w] = srow[w] | orow[w];
        __syncthreads();
    }
}

}  // namespace tsr

using namespace tsr;

extern "C" {

int tsr_spread_edges(const uint8_t* bits_in, uint8_t* bits_out, int B, int T, int F, int radius, int mode,
                     const int8_t* host_dirs, int block_len, void* stream) {
    if (B == 0) return TSR_OK;
    TSR_REQUIRE(bits_in && bits_out, "tsr_spread_edges: null pointer");
    TSR_REQUIRE(bits_in != bits_out, "tsr_spread_edges: in-place operation is not supported");
    TSR_REQUIRE(B >= 0 && T >= 0 && F >= 0 && radius >= 0 && block_len >= 0, "tsr_spread_edges: negative size");
    TSR_REQUIRE(mode == 0 || mode == 1, "tsr_spread_edges: mode must be 0 (line) or 1 (box)");
    TSR_REQUIRE(mode == 1 || host_dirs, "tsr_spread_edges: line mode needs host_dirs[8][2]");
    if ((long long)B * T * F == 0) return TSR_OK;
    SpreadTaps taps;
    taps.n = 0;
    taps.max_dt = taps.max_df = 0;
    auto add = [&](int dt, int df, int mask) -> bool {
        for (int i = 0; i < taps.n; ++i)
            if (taps.dt[i] == dt && taps.df[i] == df) {
                taps.mask[i] |= (uint8_t)mask;
                return true;
            }
        if (taps.n >= TSR_MAX_SPREAD_TAPS) return false;
        taps.dt[taps.n] = (int8_t)dt;
        taps.df[taps.n] = (int8_t)df;
        taps.mask[taps.n] = (uint8_t)mask;
        taps.n++;
        if (abs(dt) > taps.max_dt) taps.max_dt = abs(dt);
        if (abs(df) > taps.max_df) taps.max_df = abs(df);
        return true;
    };
    bool ok = true;
    if (mode == 0) {
        for (int e = 0; e < 8; ++e)
            for (int k = -radius; k <= radius; ++k) {
                int df = k * host_dirs[2 * e], dt = k * host_dirs[2 * e + 1];
                TSR_REQUIRE(abs(df) <= 120 && abs(dt) <= 120, "tsr_spread_edges: radius*direction too large");
                ok = ok && add(dt, df, 1 << e);
            }
    } else {
        TSR_REQUIRE(radius <= 5, "tsr_spread_edges: box radius > 5 unsupported (%d taps)", (2 * radius + 1) * (2 * radius + 1));
        for (int a = -radius; a <= radius; ++a)
            for (int c = -radius; c <= radius; ++c) ok = ok && add(a, c, 0xFF);
    }
    if (!ok) {
        set_error("tsr_spread_edges: more than %d distinct taps", TSR_MAX_SPREAD_TAPS);
        return TSR_ERR_UNSUPPORTED;
    }
    int rc = require_sm100("tsr_spread_edges");
    if (rc) return rc;
    int TT = (block_len > 0 && block_len <= 64) ? block_len : 32;
    if (TT > T) TT = T;
    const int padf = (taps.max_df + 3) & ~3;
    const int pitch = ((F + 3) & ~3) + 2 * padf + 4;
    size_t smem = (size_t)(TT + 2 * taps.max_dt) * pitch;
    smem = (smem + 15) & ~(size_t)15;
    TSR_REQUIRE(smem <= 200 * 1024, "tsr_spread_edges: tile does not fit shared memory (F=%d radius=%d)", F, radius);
    if (smem > 48 * 1024)
        TSR_CUDA(cudaFuncSetAttribute(spread_edges_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TSR_REQUIRE(B <= 65535, "tsr_spread_edges: B > 65535; split the batch");
    dim3 grid((T + TT - 1) / TT, B);
    spread_edges_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(bits_in, bits_out, T, F, TT, block_len, taps);
    return check_launch("tsr_spread_edges");
}

int tsr_spread_codes(const uint8_t* codes, uint32_t* featbits, int B, int To, int Fo, int K, int r0, int r1,
                     int rowwords, void* stream) {
    if (B == 0) return TSR_OK;
    TSR_REQUIRE(codes && featbits, "tsr_spread_codes: null pointer");
    TSR_REQUIRE(B >= 0 && To >= 0 && Fo >= 1 && r0 >= 0 && r1 >= 0, "tsr_spread_codes: negative size");
    TSR_REQUIRE(K >= 1 && K <= 255, "tsr_spread_codes: need 1 <= num_parts <= 255 (uint8 codes), got %d", K);
    TSR_REQUIRE(rowwords == tsr_rowwords(Fo * K), "tsr_spread_codes: rowwords must be tsr_rowwords(Fo*K)");
    if ((long long)B * To == 0) return TSR_OK;
    int rc = require_sm100("tsr_spread_codes");
    if (rc) return rc;
    const int Kw = (K + 31) >> 5;
    int RT = 37;
    while (RT > 1 && (size_t)2 * RT * Fo * Kw * 4 > 44 * 1024) RT = (RT + 1) / 2;
    if (RT > To) RT = To;
    size_t smem = (((size_t)(RT + 2 * r0) * Fo + 15) & ~(size_t)15) + (size_t)2 * RT * Fo * Kw * 4 + (size_t)rowwords * 8;
    TSR_REQUIRE(smem <= 200 * 1024, "tsr_spread_codes: tile does not fit shared memory");
    if (smem > 48 * 1024)
        TSR_CUDA(cudaFuncSetAttribute(spread_codes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TSR_REQUIRE(B <= 65535, "tsr_spread_codes: B > 65535; split the batch");
    dim3 grid((To + RT - 1) / RT, B);
    spread_codes_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(codes, featbits, To, Fo, K, r0, r1, rowwords, RT);
    return check_launch("tsr_spread_codes");
}

int tsr_spread_graph(uint32_t* featbits, const uint8_t* graph, long long R, int Fo, int K, int rowwords, void* stream) {
    TSR_REQUIRE(featbits && graph, "tsr_spread_graph: null pointer");
    TSR_REQUIRE(K >= 1 && K <= 32 * TSR_GRAPH_MAX_KW, "tsr_spread_graph: need 1 <= K <= %d", 32 * TSR_GRAPH_MAX_KW);
    TSR_REQUIRE(rowwords == tsr_rowwords(Fo * K), "tsr_spread_graph: rowwords must be tsr_rowwords(Fo*K)");
    if (R == 0) return TSR_OK;
    int rc = require_sm100("tsr_spread_graph");
    if (rc) return rc;
    const int Kw = (K + 31) >> 5;
    size_t smem = ((size_t)K * Kw + 2 * (rowwords + 1)) * 4;
    TSR_REQUIRE(smem <= 200 * 1024, "tsr_spread_graph: does not fit shared memory");
    if (smem > 48 * 1024)
        TSR_CUDA(cudaFuncSetAttribute(spread_graph_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks = R < 148 * 8 ? R : 148 * 8;
    spread_graph_kernel<<<(unsigned)blocks, 128, smem, (cudaStream_t)stream>>>(featbits, graph, R, Fo, K, rowwords);
    return check_launch("tsr_spread_graph");
}

}  // extern "C"
