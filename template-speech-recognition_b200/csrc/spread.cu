// spread.cu -- bit-packed max-filters: K1a edge spreading, K1b part-code spreading, part-graph
// spreading.  Integer / bit work, HBM-bound; every result is bit-exact by construction.
#include <algorithm>
#include <utility>

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
// rows of `pitch` bytes (4-aligned, zero padded left/right by padf >= max_df rounded to 4).  A
// thread produces one 32-bit output word (4 frequency bins).  Taps arrive sorted by dt; with
// |df| <= 4 (WORDS path) the three words (w-1, w, w+1) of a source row are loaded once per distinct
// dt and every tap of that row is a funnel shift of them; otherwise each tap fetches an unaligned
// 4-byte window = two aligned LDS.32 + one funnel shift.
__global__ void __launch_bounds__(256) spread_edges_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                           int T, int F, int TT, int block_len, int tail_lo, SpreadTaps taps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * TT;
    const int padf = (taps.max_df + 3) & ~3;
    const int Fw = (F + 3) & ~3;
    const int pitch = Fw + 2 * padf + 4;            // +4: the "hi" word of the last window
    const int rows = TT + 2 * taps.max_dt;
    const uint8_t* src = in + (size_t)b * T * F;
    const bool words = ((F & 3) == 0) && ((reinterpret_cast<uintptr_t>(in) & 3) == 0);

    // zero everything, then fill the in-range part
    for (int i = threadIdx.x * 16; i < rows * pitch; i += blockDim.x * 16)
        *reinterpret_cast<uint4*>(smem + i) = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (words) {
        const int wpr = F >> 2;
        const uint32_t* srcw = reinterpret_cast<const uint32_t*>(src);
        for (int i = threadIdx.x; i < rows * wpr; i += blockDim.x) {
            int r = i / wpr, w = i - r * wpr;
            int t = t0 - taps.max_dt + r;
            if (t >= 0 && t < T) *reinterpret_cast<uint32_t*>(smem + r * pitch + padf + 4 * w) = __ldg(srcw + (size_t)t * wpr + w);
        }
    } else {
        for (int i = threadIdx.x; i < rows * F; i += blockDim.x) {
            int r = i / F, f = i - r * F;
            int t = t0 - taps.max_dt + r;
            if (t >= 0 && t < T) smem[r * pitch + padf + f] = src[(size_t)t * F + f];
        }
    }
    __syncthreads();

    const int words_per_row = Fw >> 2;
    const bool near = taps.max_df <= 4;
    for (int i = threadIdx.x; i < TT * words_per_row; i += blockDim.x) {
        int r = i / words_per_row, w = i - r * words_per_row;
        int t = t0 + r;
        if (t >= T) break;
        int blk_lo = 0, blk_hi = T;
        if (block_len > 0) {
            blk_lo = (t / block_len) * block_len;
            blk_hi = min(T, blk_lo + block_len);
            if (tail_lo >= 0 && blk_lo + block_len > T) blk_lo = tail_lo;   // ragged last block: window of the final block_len frames
        }
        uint32_t acc = 0;
        if (near) {
            int cur_dt = -1000;
            bool rowok = false;
            uint32_t l = 0, m = 0, h = 0;
            for (int k = 0; k < taps.n; ++k) {
                const int dt = taps.dt[k];
                if (dt != cur_dt) {                     // uniform: taps are sorted by dt
                    cur_dt = dt;
                    const int ts = t + dt;
                    rowok = ts >= blk_lo && ts < blk_hi;
                    if (rowok) {
                        const uint32_t* rp = reinterpret_cast<const uint32_t*>(smem + (r + taps.max_dt + dt) * pitch + padf) + w;
                        m = rp[0];
                        if (padf) { l = rp[-1]; h = rp[1]; }
                    }
                }
                if (!rowok) continue;
                const int df = taps.df[k];
                const uint32_t v = df == 0 ? m : (df > 0 ? __funnelshift_rc(m, h, 8 * df) : __funnelshift_lc(l, m, -8 * df));
                acc |= v & (0x01010101u * taps.mask[k]);
            }
        } else {
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
        }
        uint8_t* dst = out + ((size_t)b * T + t) * F + 4 * w;
        if (words) {
            *reinterpret_cast<uint32_t*>(dst) = acc;
        } else {
            for (int q = 0; q < 4 && 4 * w + q < F; ++q) dst[q] = (uint8_t)(acc >> (8 * q));
        }
    }
}

// Line mode with unit directions (every component of every orientation in {-1,0,1}: the only table
// the reference produces, edge_signal_proc.py:1751-1821) and radius R <= 3: the taps fall into four
// classes -- H (df=+-k, dt=0), V (df=0, dt=+-k), D (df=k, dt=k), A (df=-k, dt=k) -- each with one
// orientation mask.  A thread owns one 32-bit word column (4 frequency bins) of a run of RL rows and
// slides down it with the 2R+1 source rows (left, middle, right words) in registers: 3 LDS + ~10R
// shift/logic instructions per output word instead of a table walk.  Runs never cross a block.
struct SpreadLineMasks { uint32_t h, v, d, a; };   // orientation masks x 0x01010101

template <int R>
__global__ void __launch_bounds__(128) spread_edges_line_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                                int T, int wpr, int TT, int RL, int block_len, int tail_lo,
                                                                SpreadLineMasks mk) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint32_t* sw = reinterpret_cast<uint32_t*>(smem);
    constexpr int W = 2 * R + 1;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * TT;
    const int pitch = wpr + 2;                       // one zero word left and right of every row
    const int rows = TT + 2 * R;
    const uint32_t* src = in + (size_t)b * T * wpr;
    for (int i = threadIdx.x; i < rows * pitch; i += blockDim.x) {
        const int r = i / pitch, c = i - r * pitch - 1;
        const int t = t0 - R + r;
        sw[i] = (c >= 0 && c < wpr && t >= 0 && t < T) ? __ldg(src + (size_t)t * wpr + c) : 0u;
    }
    __syncthreads();
    const int nrun = (TT + RL - 1) / RL;
    for (int task = threadIdx.x; task < nrun * wpr; task += blockDim.x) {
        const int run = task / wpr, w = task - run * wpr;
        const int ra = run * RL;                      // tile-relative first row of the run
        const int ta = t0 + ra;
        if (ta >= T) continue;
        const int rb = min(min(ra + RL, TT), T - t0);
        int lo = 0, hi = T;                           // block of the run in absolute frames
        if (block_len > 0) {
            lo = (ta / block_len) * block_len; hi = min(T, lo + block_len);
            if (tail_lo >= 0 && lo + block_len > T) lo = tail_lo;            // ragged last block: window of the final block_len frames
        }
        const uint32_t* col = sw + R * pitch + 1 + w; // tile row 0, word w
        uint32_t wl[W], wm[W], wh[W];
        auto load = [&](int slot, int r) {            // tile-relative source row r (may be a halo row)
            const int t = t0 + r;
            const bool ok = t >= lo && t < hi;
            const uint32_t* q = col + r * pitch;
            wl[slot] = ok ? q[-1] : 0u;
            wm[slot] = ok ? q[0] : 0u;
            wh[slot] = ok ? q[1] : 0u;
        };
#pragma unroll
        for (int k = 0; k < W - 1; ++k) load(k, ra - R + k);
        uint32_t* dst = out + ((size_t)b * T + t0) * wpr + w;
        for (int r = ra; r < rb; r += W) {
#pragma unroll
            for (int u = 0; u < W; ++u) {
                if (r + u < rb) {
                    load((u + W - 1) % W, r + u + R);
                    const int c0 = (u + R) % W;
                    uint32_t acc = wm[c0];
#pragma unroll
                    for (int j = 1; j <= R; ++j)
                        acc |= (__funnelshift_r(wm[c0], wh[c0], 8 * j) | __funnelshift_l(wl[c0], wm[c0], 8 * j)) & mk.h;
#pragma unroll
                    for (int k = 1; k <= R; ++k) {
                        const int sp = (u + R + k) % W, sn = (u + R - k + W) % W;     // rows t+k, t-k
                        const uint32_t pr = __funnelshift_r(wm[sp], wh[sp], 8 * k);   // row t+k at f+k
                        const uint32_t pl = __funnelshift_l(wl[sp], wm[sp], 8 * k);   // row t+k at f-k
                        const uint32_t nr = __funnelshift_r(wm[sn], wh[sn], 8 * k);   // row t-k at f+k
                        const uint32_t nl = __funnelshift_l(wl[sn], wm[sn], 8 * k);   // row t-k at f-k
                        acc |= ((wm[sp] | wm[sn]) & mk.v) | ((pr | nl) & mk.d) | ((pl | nr) & mk.a);
                    }
                    dst[(size_t)(r + u) * wpr] = acc;
                }
            }
        }
    }
}

// ================================================================================= K1b
// codes uint8 [B][To][Fo] -> featbits uint32 [B][To][rowwords], feature j*K + (p-1).
// The (2*r0+1) x (2*r1+1) box dilation is separable.  CTA = (b, tile of RT output rows):
//   A. H[x0][:] = frequency-direction dilation of input row x0 as a PACKED bit row: every nonzero
//      code p at (x0, x1) ORs bit (x1+k)*K + p-1, |k| <= r1, into row x0 of the smem tile
//      (2*r1+1 shared-memory atomics per nonzero code; lanes run along x0 and the row pitch is odd,
//      so the atomics of one instruction fall into distinct banks);
//   B. out[i][w] = OR_{a <= 2*r0} H[i+a][w]: plain word-wise ORs on the packed rows.  A thread owns
//      one word column and slides down the rows with the window in registers: one LDS, two LOP3
//      and one coalesced STG per output word.
// Rows of H outside the utterance stay zero (clipping to the array, code_parts.pyx:445-452).
template <int R0>
__device__ __forceinline__ void sc_vertical_or(const uint32_t* __restrict__ H, uint32_t* __restrict__ out, int pitch,
                                               int rowwords, int nrow, int r0, int D) {
    if constexpr (R0 >= 0) {
        // 64-bit words (rowwords, the pitch and H's word offset are even, the output rows 8-byte aligned): half the
        // LDS / STG instructions of the 32-bit walk -- the kernel is issue bound.  tasks = (row segment, word pair).
        const int npair = rowwords >> 1;
        const int nseg = max(1, (int)blockDim.x / npair);
        const int per = (nrow + nseg - 1) / nseg;
        constexpr int W = 2 * R0 + 1;
        for (int task = threadIdx.x; task < nseg * npair; task += blockDim.x) {
            const int seg = task / npair, w = 2 * (task - seg * npair);
            const int ra = seg * per, rb = min(nrow, ra + per);
            if (ra >= rb) continue;
            // bits >= D of the row are padding and stay zero (out-of-range cells spill into them in phase A)
            const int live = D - 32 * w;
            uint2 keep;
            keep.x = live >= 32 ? 0xFFFFFFFFu : (live <= 0 ? 0u : ((1u << live) - 1u));
            keep.y = live >= 64 ? 0xFFFFFFFFu : (live <= 32 ? 0u : ((1u << (live - 32)) - 1u));
            const uint32_t* hp = H + ra * pitch + w;
            uint32_t* op = out + (size_t)ra * rowwords + w;
            uint2 win[W];
#pragma unroll
            for (int a = 0; a < W - 1; ++a) win[a] = *reinterpret_cast<const uint2*>(hp + a * pitch);
            hp += (W - 1) * pitch;
            for (int r = ra; r < rb; r += W) {
#pragma unroll
                for (int u = 0; u < W; ++u) {
                    if (r + u < rb) {
                        win[(u + W - 1) % W] = *reinterpret_cast<const uint2*>(hp);
                        hp += pitch;
                        uint2 acc = make_uint2(0u, 0u);
#pragma unroll
                        for (int a = 0; a < W; ++a) { acc.x |= win[a].x; acc.y |= win[a].y; }
                        *reinterpret_cast<uint2*>(op) = make_uint2(acc.x & keep.x, acc.y & keep.y);
                        op += rowwords;
                    }
                }
            }
        }
    } else {
        // generic radius: tasks = (row segment, word column); a thread takes tasks blockDim.x apart
        const int nseg = max(1, (int)blockDim.x / rowwords);
        const int per = (nrow + nseg - 1) / nseg;
        for (int task = threadIdx.x; task < nseg * rowwords; task += blockDim.x) {
            const int seg = task / rowwords, w = task - seg * rowwords;
            const int ra = seg * per, rb = min(nrow, ra + per);
            if (ra >= rb) continue;
            const int live = D - 32 * w;
            const uint32_t keep = live >= 32 ? 0xFFFFFFFFu : (live <= 0 ? 0u : ((1u << live) - 1u));
            const uint32_t* hp = H + ra * pitch + w;
            uint32_t* op = out + (size_t)ra * rowwords + w;
            for (int r = ra; r < rb; ++r) {
                uint32_t acc = 0;
                for (int a = 0; a <= 2 * r0; ++a) acc |= hp[a * pitch];
                *op = acc & keep;
                hp += pitch;
                op += rowwords;
            }
        }
    }
}
template <int R1>
__device__ __forceinline__ void sc_scatter(uint32_t* hrow, int d, int K, int r1) {
    if constexpr (R1 >= 0) {
#pragma unroll
        for (int k = 0; k < 2 * R1 + 1; ++k, d += K) atomicOr(hrow + (d >> 5), 1u << (d & 31));
    } else {
        for (int k = 0; k <= 2 * r1; ++k, d += K) atomicOr(hrow + (d >> 5), 1u << (d & 31));
    }
}

// H rows: [offw slack words | rowwords | right slack]; pitch odd.
template <int R1>
__global__ void __launch_bounds__(256) spread_codes_kernel(const uint8_t* __restrict__ codes,
                                                           uint32_t* __restrict__ featbits, int To, int Fo, int K,
                                                           int r0, int r1, int rowwords, int RT, int pitch, int offw) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int b = blockIdx.y;
    const int i0 = blockIdx.x * RT;
    const int rows_in = RT + 2 * r0;
    const int nrow = min(RT, To - i0);
    uint32_t* H = reinterpret_cast<uint32_t*>(smem);                    // [rows_in][pitch]

    const int hwords = rows_in * pitch;
    for (int i = threadIdx.x * 4; i < ((hwords + 3) & ~3); i += blockDim.x * 4)
        *reinterpret_cast<uint4*>(H + i) = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    // A: horizontal dilation into packed rows.  Rows outside the utterance contribute nothing.
    const int xlo = max(0, r0 - i0), xhi = min(rows_in, To - i0 + r0);   // tile rows inside the utterance
    const uint8_t* gsrc = codes + ((long long)b * To + i0 - r0) * Fo;   // tile row 0
    const int dbase = offw * 32 - r1 * K - 1;                            // bit of (column 0 - r1, part 0)
    if (((Fo & 3) == 0) && ((reinterpret_cast<uintptr_t>(codes) & 3) == 0)) {
        // a thread takes 4 consecutive cells (one 32-bit word of codes); silent rows skip as whole warps
        const int Fw = Fo >> 2;
        const int n = (xhi - xlo) * Fw;
        const uint32_t* gw = reinterpret_cast<const uint32_t*>(gsrc) + xlo * Fw;
        const float inv = 1.0f / (float)Fw;
        for (int q = threadIdx.x; q < n; q += blockDim.x) {
            const uint32_t c4 = __ldg(gw + q);
            if (c4 == 0u) continue;
            int x0 = (int)(((float)q + 0.5f) * inv);
            const int wq = q - x0 * Fw;
            x0 += xlo;
            uint32_t* hrow = H + x0 * pitch;
            const int d0 = dbase + 4 * wq * K;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int pcode = (c4 >> (8 * i)) & 255;
                if (pcode) sc_scatter<R1>(hrow, d0 + i * K + pcode, K, r1);
            }
        }
    } else {
        const int n = (xhi - xlo) * Fo;
        const float inv = 1.0f / (float)Fo;
        for (int q = threadIdx.x; q < n; q += blockDim.x) {
            const int pcode = __ldg(gsrc + xlo * Fo + q);
            if (pcode == 0) continue;
            int x0 = (int)(((float)q + 0.5f) * inv);
            const int x1 = q - x0 * Fo;
            x0 += xlo;
            sc_scatter<R1>(H + x0 * pitch, dbase + x1 * K + pcode, K, r1);
        }
    }
    __syncthreads();
    uint32_t* out = featbits + ((size_t)b * To + i0) * rowwords;
    const uint32_t* H0 = H + offw;
    const int D = Fo * K;
    switch (r0) {
        case 0: sc_vertical_or<0>(H0, out, pitch, rowwords, nrow, r0, D); break;
        case 1: sc_vertical_or<1>(H0, out, pitch, rowwords, nrow, r0, D); break;
        case 2: sc_vertical_or<2>(H0, out, pitch, rowwords, nrow, r0, D); break;
        case 3: sc_vertical_or<3>(H0, out, pitch, rowwords, nrow, r0, D); break;
        case 5: sc_vertical_or<5>(H0, out, pitch, rowwords, nrow, r0, D); break;
        default: sc_vertical_or<-1>(H0, out, pitch, rowwords, nrow, r0, D); break;
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
                     const int8_t* host_dirs, int block_len, int tail_window, void* stream) {
    if (B == 0) return TSR_OK;
    TSR_REQUIRE(bits_in && bits_out, "tsr_spread_edges: null pointer");
    TSR_REQUIRE(bits_in != bits_out, "tsr_spread_edges: in-place operation is not supported");
    TSR_REQUIRE(B >= 0 && T >= 0 && F >= 0 && radius >= 0 && block_len >= 0, "tsr_spread_edges: negative size");
    TSR_REQUIRE(mode == 0 || mode == 1, "tsr_spread_edges: mode must be 0 (line) or 1 (box)");
    TSR_REQUIRE(mode == 1 || host_dirs, "tsr_spread_edges: line mode needs host_dirs[8][2]");
    if ((long long)B * T * F == 0) return TSR_OK;
    // tail_window: the frames past the last whole block see neighbours down to frame T - block_len (the reference
    // spreads the window of the final block_len frames and keeps its last T % block_len columns)
    const int tail_lo = (tail_window && block_len > 0 && T % block_len) ? (T > block_len ? T - block_len : 0) : -1;
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
    TSR_REQUIRE(B <= 65535, "tsr_spread_edges: B > 65535; split the batch");
    // fast path: line mode, unit directions, radius 1..3, word-aligned rows
    if (mode == 0 && radius >= 1 && radius <= 3 && (F & 3) == 0 && ((uintptr_t)bits_in & 3) == 0 && ((uintptr_t)bits_out & 3) == 0) {
        bool unit = true;
        SpreadLineMasks mk = {0u, 0u, 0u, 0u};
        for (int e = 0; e < 8; ++e) {
            const int df = host_dirs[2 * e], dt = host_dirs[2 * e + 1];
            if (abs(df) > 1 || abs(dt) > 1) unit = false;
            const uint32_t bit = 0x01010101u << e;
            if (df != 0 && dt == 0) mk.h |= bit;
            else if (df == 0 && dt != 0) mk.v |= bit;
            else if (df * dt > 0) mk.d |= bit;
            else if (df * dt < 0) mk.a |= bit;
        }
        int RL = 0;     // rows per run: a divisor of block_len in [8,20], else the whole block, else 16
        if (block_len == 0) RL = 16;
        else {
            for (int c = 20; c >= 8 && !RL; --c) if (block_len % c == 0) RL = c;
            if (!RL && block_len <= 64) RL = block_len;
        }
        if (unit && RL) {
            int TT = block_len == 0 ? 128 : (block_len >= 128 ? block_len : block_len * (128 / block_len));
            if (TT > T) TT = block_len == 0 ? T : ((T + block_len - 1) / block_len) * block_len;
            const int wpr = F >> 2;
            const size_t smem = (size_t)(TT + 2 * radius) * (wpr + 2) * 4;
            if (smem <= 48 * 1024) {
                dim3 grid((T + TT - 1) / TT, B);
                const uint32_t* in32 = reinterpret_cast<const uint32_t*>(bits_in);
                uint32_t* out32 = reinterpret_cast<uint32_t*>(bits_out);
                cudaStream_t st = (cudaStream_t)stream;
                if (radius == 1) spread_edges_line_kernel<1><<<grid, 128, smem, st>>>(in32, out32, T, wpr, TT, RL, block_len, tail_lo, mk);
                else if (radius == 2) spread_edges_line_kernel<2><<<grid, 128, smem, st>>>(in32, out32, T, wpr, TT, RL, block_len, tail_lo, mk);
                else spread_edges_line_kernel<3><<<grid, 128, smem, st>>>(in32, out32, T, wpr, TT, RL, block_len, tail_lo, mk);
                return check_launch("tsr_spread_edges");
            }
        }
    }
    // sort the taps by dt (insertion sort; <= 160 entries): the kernel reloads a source row per distinct dt
    for (int i = 1; i < taps.n; ++i)
        for (int j = i; j > 0 && taps.dt[j - 1] > taps.dt[j]; --j) {
            std::swap(taps.dt[j - 1], taps.dt[j]);
            std::swap(taps.df[j - 1], taps.df[j]);
            std::swap(taps.mask[j - 1], taps.mask[j]);
        }
    // tile = about 128 frames, a whole number of blocks when spreading is per block
    int TT = 128;
    if (block_len > 0) TT = block_len >= 128 ? block_len : block_len * (128 / block_len);
    if (TT > T) TT = T;
    const int padf = (taps.max_df + 3) & ~3;
    const int pitch = ((F + 3) & ~3) + 2 * padf + 4;
    while (TT > 8 && (size_t)(TT + 2 * taps.max_dt) * pitch > 64 * 1024) TT = (TT + 1) / 2;
    size_t smem = (size_t)(TT + 2 * taps.max_dt) * pitch;
    smem = ((smem + 15) & ~(size_t)15) + 16;
    TSR_REQUIRE(smem <= 200 * 1024, "tsr_spread_edges: tile does not fit shared memory (F=%d radius=%d)", F, radius);
    if (smem > 48 * 1024)
        TSR_CUDA(cudaFuncSetAttribute(spread_edges_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((T + TT - 1) / TT, B);
    spread_edges_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(bits_in, bits_out, T, F, TT, block_len, tail_lo, taps);
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
    // packed smem rows: left slack of ceil(r1*K/32) words, the row, right slack for columns Fo..Fo+r1-1.  4 tiles per 296-row utterance at the BASELINE shape.
    // even word offset and even pitch (64-bit accesses in phase B), pitch = 2 mod 4 so that rows do not alias banks
    const int offw = ((r1 * K + 31) / 32 + 1) & ~1;
    const int needw = offw + ((Fo + r1) * K + 31) / 32 + 1;
    int pitch = ((needw > offw + rowwords ? needw : offw + rowwords) + 1) & ~1;
    if ((pitch & 3) == 0) pitch += 2;
    int RT = 74;
    while (RT > 4 && (size_t)(RT + 2 * r0) * pitch * 4 > 96 * 1024) RT = (RT + 1) / 2;
    if (RT > To) RT = To;
    size_t smem = (((size_t)(RT + 2 * r0) * pitch * 4 + 15) & ~(size_t)15) + 16;
    TSR_REQUIRE(smem <= 200 * 1024, "tsr_spread_codes: tile does not fit shared memory");
    TSR_REQUIRE(B <= 65535, "tsr_spread_codes: B > 65535; split the batch");
    dim3 grid((To + RT - 1) / RT, B);
    cudaStream_t st = (cudaStream_t)stream;
#define TSR_SC_LAUNCH(R1)                                                                                          \
    do {                                                                                                           \
        if (smem > 48 * 1024)                                                                                      \
            TSR_CUDA(cudaFuncSetAttribute(spread_codes_kernel<R1>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                          (int)smem));                                                             \
        spread_codes_kernel<R1><<<grid, 256, smem, st>>>(codes, featbits, To, Fo, K, r0, r1, rowwords, RT, pitch, \
                                                         offw);                                                    \
    } while (0)
    switch (r1) {
        case 0: TSR_SC_LAUNCH(0); break;
        case 1: TSR_SC_LAUNCH(1); break;
        case 2: TSR_SC_LAUNCH(2); break;
        case 3: TSR_SC_LAUNCH(3); break;
        case 5: TSR_SC_LAUNCH(5); break;
        default: TSR_SC_LAUNCH(-1); break;
    }
#undef TSR_SC_LAUNCH
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
    long long blocks = R < (long long)sm_count() * 8 ? R : (long long)sm_count() * 8;
    spread_graph_kernel<<<(unsigned)blocks, 128, smem, (cudaStream_t)stream>>>(featbits, graph, R, Fo, K, rowwords);
    return check_launch("tsr_spread_graph");
}

}  // extern "C"
