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
// Warp ids are ordered by how critical the role is (the SMSP arbiter favours higher ids):
//   warp  0     filter-tile loader (cp.async.bulk, one copy per pipeline stage)
//   warps 1-8   frame-bit expanders, two groups of 4 warps on alternate stages (a thread expands one row
//               of each 128-row sub-tile)
//   warps 9-12  epilogue: TMEM -> smem staging -> diagonal sums -> global (TMEM lane quadrant = warp % 4)
//   warp  13    MMA issuer + TMEM allocation
#include <type_traits>

#include <cstdlib>

#include "common.cuh"

namespace tsr {

#define DT_MSUB_MAX 3
#define DT_MAX_STAGES 4
#define DT_MAX_BSTAGES 8   // I8: ring of 64-feature filter tiles
#define DT_THREADS 480     // 15 warps: <= 136 registers per thread
#define DT_PRODUCERS 256   // two groups of 128 expander threads
#define DT_TAU_CHUNK 40    // diagonal sums are staged through smem in chunks of <= 40 filter frames
#define DT_TN_MAX_STAGES 8  // TN kernel: ring of 16 KB filter-tile pairs
#define DT_TN_STEP_MAX 256 // TN kernel: frames per step (UMMA N)
#define DT_TN_MAX_CARRY 96 // TN kernel: unfinished frames carried between steps (fold * (Lf - 1))
#define DT_NT_MAX 240      // widest N tile: tcgen05.mma issue costs ~100 cycles, so each MMA should carry >= that much work

// Operand kinds.  F16: hi+lo 16-bit float planes accumulate into the same fp32 TMEM columns.
// I8: the filters are fixed-point integers q = round(w / scale[m]) split into P balanced base-256
// digits (int8 planes) that occupy separate column groups; bits are expanded to uint8 {0,1};
// kind::i8 MMAs (K = 32, twice the f16 rate) accumulate exactly in int32, the diagonal sums stay
// integer, and the planes are recombined and scaled once per output value.
struct FilterLayout {
    int M, L, D, rowwords, nchunks;
    int kind;     // 0 = F16 planes, 1 = I8 planes (detect_tc_kernel), 2 = I8 planes in the TN layout (detect_tn_kernel)
    int P;        // planes (2 for F16; 2 or 3 for I8)
    int vpt;      // column groups per template: P for I8, 1 for F16
    int tptv;     // column groups ("virtual templates") per N tile, a multiple of vpt
    int NT;       // N tile = roundup16(tptv*L)
    int ntiles;   // ceil(M*vpt / tptv)
    int msub;     // 128-row accumulators per step (msub*NT <= 512)
    int nstage;   // smem pipeline depth
    int fps;      // features per filter tile: 64 (F16: 8 k-groups of 8 halves; I8: 4 k-groups of 16 bytes)
    int ncf;      // I8: 128-feature chunks per frame; the expanded frame strip of a chunk serves all `fold` phases
    int strip_rows, kgs;   // I8: rows of an A strip (msub*128 + fold - 1, rounded up to 8) and its k-group stride in bytes
    int fold;     // frames folded into the contraction ("super-row" = fold consecutive frames); 1 for F16
    int Lf;       // columns per group = ceil(L / fold): column tau' covers frames tau'*fold .. +fold-1
    int Lc;       // fold * (Lf - 1): row shift of the last column = unfinished rows carried between steps
    int Kf;       // contraction length per super-row = fold * rowwords * 32
    size_t tile_bytes;   // one (ntile, chunk) filter tile
    size_t hdr_bytes;    // [0] plane format, [1] max|w| bits, then float scale[M] (I8)
    size_t total;
};

__host__ __device__ inline FilterLayout filter_layout(int M, int L, int D, int planes) {
    FilterLayout f;
    f.M = M; f.L = L; f.D = D;
    f.rowwords = ((D + 63) / 64) * 2;
    f.kind = planes == TSR_PLANES_INT8X2 ? 2 : (planes == TSR_PLANES_INT8X3 ? 1 : 0);
    f.P = planes == TSR_PLANES_INT8X3 ? 3 : 2;
    f.vpt = f.kind ? f.P : 1;
    f.fps = 64;
    f.ncf = (f.rowwords * 32 + 127) / 128;
    f.strip_rows = f.kgs = 0;
    f.hdr_bytes = 128 + (((size_t)M * 4 + 127) & ~(size_t)127);
    if (f.kind == 2) {
        // TN layout (detect_tn_kernel): the filter columns are the UMMA *M* operand -- tiles of 128 columns = 128 / Lf
        // groups -- and the frames the N operand (<= 256 per step).  Lf is a power of two <= 8 so that the diagonal sum
        // over a group's Lf columns is a butterfly over adjacent TMEM lanes.  MMA work ~ ntiles * fold; near-ties go to
        // the smaller Lf (fewer tiles re-reading the frame bits, longer-lived A strips, shorter epilogue).
        double best = -1.0;
        int bestLf = 1;
        for (int Lf = 1; Lf <= 8; Lf *= 2) {
            const int fo = (L + Lf - 1) / Lf;
            if (fo * (Lf - 1) > DT_TN_MAX_CARRY) continue;
            const int tp = 128 / Lf;
            const int nt = (M * 2 + tp - 1) / tp;
            const double cost = (double)nt * fo * (1.0 + 0.01 * Lf);
            if (best < 0 || cost < best * 0.97) { best = cost; bestLf = Lf; }
        }
        f.Lf = bestLf;
        f.fold = (L + f.Lf - 1) / f.Lf;
        f.Lc = f.fold * (f.Lf - 1);
        f.Kf = f.fold * f.rowwords * 32;
        f.nchunks = f.ncf * f.fold * 2;
        f.tptv = 128 / f.Lf;
        if (f.tptv > M * 2) f.tptv = M * 2;
        f.NT = 128;
        f.ntiles = (M * 2 + f.tptv - 1) / f.tptv;
        f.msub = 1;
        f.tile_bytes = 128 * 64;
        f.strip_rows = (DT_TN_STEP_MAX + f.fold - 1 + 7) & ~7;
        f.kgs = f.strip_rows * 16;
        const long long rest = 227 * 1024 - 1024 - 2LL * 8 * f.kgs - (long long)f.Lc * 128 * 4;
        int ns = (int)(rest / (2 * (long long)f.tile_bytes));      // a ring stage = the two 64-feature tiles of one phase
        f.nstage = ns > DT_TN_MAX_STAGES ? DT_TN_MAX_STAGES : (ns < 2 ? 2 : ns);
        f.total = f.hdr_bytes + (size_t)f.ntiles * f.nchunks * f.tile_bytes;
        return f;
    }
    // Frame folding (I8 only: integer accumulation is exact however long the chain).  Folding `fold`
    // consecutive frames into K divides the number of (row, column) products the epilogue has to
    // diagonal-sum by `fold` at identical MMA work.  Pick the fold that fills the N tiles best;
    // ties go to the larger fold.
    int best_fold = 1;
    if (f.kind) {
        double best = -1.0;
        for (int fo = 1; fo <= L; ++fo) {
            const int Lf = (L + fo - 1) / fo;
            if (f.vpt * Lf > DT_NT_MAX) continue;
            int tp = (DT_NT_MAX / (f.vpt * Lf)) * f.vpt;
            if (tp > M * f.vpt) tp = M * f.vpt;
            const int NT = (tp * Lf + 15) & ~15;
            const int nt = (M * f.vpt + tp - 1) / tp;
            double eff = (double)(M * f.vpt) * L / fo / ((double)nt * NT);    // useful columns / issued columns
            eff *= NT >= 208 ? 1.0 : (double)NT / 208.0;                        // small-N MMAs are issue bound
            eff /= 1.0 + 0.65 / fo;                                             // epilogue ~0.65x the MMA time unfolded
            if (eff >= best - 1e-9) { best = eff; best_fold = fo; }
        }
    }
#if !defined(__CUDA_ARCH__) && defined(TSR_DEBUG_TRACE)
    if (f.kind)   // experiments only: force the fold (must be the same when the filters are prepared and when they are used)
        if (const char* ov = getenv("TSR_DEBUG_DETECT_FOLD")) {
            const int fo = atoi(ov);
            if (fo >= 1 && fo <= L && f.vpt * ((L + fo - 1) / fo) <= DT_NT_MAX) best_fold = fo;
        }
#endif
    f.fold = best_fold;
    f.Lf = (L + f.fold - 1) / f.fold;
    f.Lc = f.fold * (f.Lf - 1);
    f.Kf = f.fold * f.rowwords * 32;
    // I8 contraction order: (frame chunk cf, phase phi, 64-feature half) -- one expanded A strip per cf
    f.nchunks = f.kind ? f.ncf * f.fold * 2 : (f.Kf + f.fps - 1) / f.fps;
    int tptv = (DT_NT_MAX / (f.vpt * f.Lf)) * f.vpt;
    if (tptv < f.vpt) tptv = f.vpt;
    if (tptv > M * f.vpt) tptv = M * f.vpt;
    int msub_cap = DT_MSUB_MAX;
    for (;;) {   // shrink rows per step, then templates per tile, until two pipeline stages fit in 227 KB
        f.tptv = tptv;
        f.NT = (tptv * f.Lf + 15) & ~15;
        f.ntiles = (M * f.vpt + tptv - 1) / tptv;
        int msub = 512 / f.NT;
        f.msub = msub > msub_cap ? msub_cap : msub;
        f.tile_bytes = f.kind ? (size_t)f.NT * 64 : (size_t)2 * f.NT * 128;
        const size_t fixed = (((128 + DT_TAU_CHUNK) * (DT_TAU_CHUNK + 1) * 4 + 127) & ~(size_t)127) +
                             ((((size_t)f.msub * 128 + f.Lc + 1) * (f.tptv | 1) * 4 + 127) & ~(size_t)127) + 512;
        const long long avail = 227 * 1024 - (long long)fixed;
        int ns;
        if (f.kind) {   // two A strips + a ring of filter tiles
            f.strip_rows = (f.msub * 128 + f.fold - 1 + 7) & ~7;
            f.kgs = f.strip_rows * 16;
            const long long rest = avail - 2LL * 8 * f.kgs;
            ns = rest > 0 ? (int)(rest / (long long)f.tile_bytes) : 0;
        } else {
            const size_t stage = (size_t)f.msub * 16384 + f.tile_bytes;
            ns = avail > 0 ? (int)(avail / (long long)stage) : 0;
        }
        if (ns >= 2 || (f.msub == 1 && tptv == f.vpt)) {
            const int cap = f.kind ? DT_MAX_BSTAGES : DT_MAX_STAGES;
            f.nstage = ns < 2 ? 2 : (ns > cap ? cap : ns);
            break;
        }
        if (f.msub > 1) msub_cap = f.msub - 1;
        else tptv -= f.vpt;
    }
    f.total = f.hdr_bytes + (size_t)f.ntiles * f.nchunks * f.tile_bytes;
    return f;
}

// per-template max |w| (I8 scales) and the global max (F16 format choice)
__global__ void filters_maxabs_kernel(const float* __restrict__ LF, int M, size_t per_template, uint8_t* __restrict__ blob) {
    uint32_t* hdr = reinterpret_cast<uint32_t*>(blob);
    uint32_t* tmax = hdr + 32;   // scale slots double as max|w| bit patterns until published
    const int m = blockIdx.y;
    const float* w = LF + (size_t)m * per_template;
    uint32_t mx = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < per_template; i += (size_t)gridDim.x * blockDim.x)
        mx = max(mx, __float_as_uint(fabsf(w[i])));
    for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if ((threadIdx.x & 31) == 0) { atomicMax(&tmax[m], mx); atomicMax(&hdr[1], mx); }
}
__global__ void filters_publish_kernel(uint8_t* __restrict__ blob, int plane_format, FilterLayout f) {
    uint32_t* hdr = reinterpret_cast<uint32_t*>(blob);
    if (threadIdx.x == 0 && blockIdx.x == 0)
        hdr[0] = f.kind ? (uint32_t)(100 + f.P) : pick_format(hdr[1], plane_format == TSR_PLANES_BF16 ? 1 : (plane_format == TSR_PLANES_FP16 ? 2 : 0));
    if (f.kind) {   // turn max|w| into the quantisation step of each template
        const float qmax = f.P == 3 ? 8355711.f : 32639.f;   // largest value of P balanced base-256 digits
        float* scale = reinterpret_cast<float*>(hdr + 32);
        for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < f.M; m += gridDim.x * blockDim.x) {
            float mx = __uint_as_float(reinterpret_cast<uint32_t*>(scale)[m]);
            scale[m] = mx > 0.f ? mx / qmax : 1.f;
        }
    }
}

// One thread per weight element of a (tile, chunk): column n = (group v, frame tau), v = m*vpt + plane.
__global__ void prepare_filters_kernel(const float* __restrict__ LF, uint8_t* __restrict__ blob, FilterLayout f,
                                       int plane_format) {
    // One thread per 16-byte row of a filter tile (16 int8 digits or 8 16-bit values of one column n and one k-group):
    // the index arithmetic is paid once per row, the loads are 64 / 32 contiguous bytes, consecutive threads write
    // consecutive rows.  (The EM M-step rebuilds its "filters" -- the affinities -- every iteration.)
    const uint32_t* hdr = reinterpret_cast<const uint32_t*>(blob);
    const uint32_t fmt = pick_format(hdr[1], plane_format == TSR_PLANES_BF16 ? 1 : (plane_format == TSR_PLANES_FP16 ? 2 : 0));
    const uint32_t* tmaxbits = hdr + 32;
    uint8_t* tiles = blob + f.hdr_bytes;
    const int ept = f.kind ? 16 : 8;                    // elements per thread
    const int nkg = f.fps / ept;                        // k-groups per chunk
    const size_t nrows = (size_t)f.ntiles * f.nchunks * nkg * f.NT;
    const float qmax = f.P == 3 ? 8355711.f : 32639.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nrows; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % f.NT);
        size_t r = i / f.NT;
        const int kg = (int)(r % nkg); r /= nkg;
        const int c = (int)(r % f.nchunks);
        const int nt = (int)(r / f.nchunks);
        const int dd0 = kg * ept;                       // first feature of the row within the chunk
        int phi, d0;                                    // frame within the super-row, first feature within the frame
        if (f.kind) {   // chunk c = (cf * fold + phi) * 2 + half
            const int half = c & 1, cp = c >> 1;
            phi = cp % f.fold;
            d0 = (cp / f.fold) * 128 + half * 64 + dd0;
        } else {
            const int kf = c * f.fps + dd0;             // position in the contraction (fold == 1); rows never straddle frames
            const int dpad = f.rowwords * 32;
            phi = kf / dpad; d0 = kf - phi * dpad;
        }
        const int vl = n / f.Lf, tauf = n - vl * f.Lf;
        const int tau = tauf * f.fold + phi;
        const int v = nt * f.tptv + vl;
        const int m = v / f.vpt, plane = v - m * f.vpt;
        const bool live = vl < f.tptv && m < f.M && phi < f.fold && tau < f.L;
        const float* src = LF + ((size_t)(live ? m : 0) * f.L + (live ? tau : 0)) * f.D;
        uint8_t* tile = tiles + ((size_t)nt * f.nchunks + c) * f.tile_bytes;
        if (f.kind) {
            float sc = 1.f;
            if (live) {
                const float mx = __uint_as_float(tmaxbits[m]);
                sc = mx > 0.f ? mx / qmax : 1.f;
            }
            uint32_t w4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                int digit = 0;
                if (live && d0 + e < f.D) {
                    float qf = rintf(src[d0 + e] / sc);
                    qf = fminf(fmaxf(qf, -qmax), qmax);
                    int q = (int)qf;
                    for (int pl = 0; pl <= plane; ++pl) {      // balanced base-256 digits, least significant first
                        digit = ((q + 128) & 255) - 128;
                        q = (q - digit) >> 8;
                    }
                }
                w4[e >> 2] |= (uint32_t)(digit & 255) << (8 * (e & 3));
            }
            *reinterpret_cast<uint4*>(tile + ((size_t)kg * f.NT + n) * 16) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        } else {
            uint32_t h4[4] = {0u, 0u, 0u, 0u}, l4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float w = (live && d0 + e < f.D) ? src[d0 + e] : 0.f;
                uint16_t hi, lo;
                split_hi_lo((double)w, fmt, hi, lo);
                h4[e >> 1] |= (uint32_t)hi << (16 * (e & 1));
                l4[e >> 1] |= (uint32_t)lo << (16 * (e & 1));
            }
            const size_t o = ((size_t)kg * f.NT + n) * 16;
            *reinterpret_cast<uint4*>(tile + o) = make_uint4(h4[0], h4[1], h4[2], h4[3]);
            *reinterpret_cast<uint4*>(tile + (size_t)f.NT * 128 + o) = make_uint4(l4[0], l4[1], l4[2], l4[3]);
        }
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
    int ntile_group;        // N tiles whose filter tiles share the L2 at one time (unit order: group, row unit, tile in group)
    int step_rows;          // TN kernel: frames per step (UMMA N, a multiple of 16)
    int nissue;             // I8: MMA issuer warps (2 when the whole filter bank is L2 resident, else 1 -- measured, see tsr_detect)
    FilterLayout f;
    long long* trace;   // optional debug timeline of CTA 0: [stage iteration][8] clock64 stamps
    int dbg;            // debug experiments (timing only, wrong results): 1 = every phase reads the strip at row 0
    unsigned long long* keys;   // optional [R]: running max over templates of (score + cs[m]), see detect_key
    const float* cs;            // optional [M] constants added before the max (detect_float_max :258, get_train_data.py:964-992)
};

// (value, template) packed so that an unsigned 64-bit max picks the larger value and, among equal values, the
// LOWER template index (np.maximum chain / strict > of detect_float_max :275): the fused mixture max needs no
// [R][M] score matrix and no second pass -- units holding different template tiles combine with atomicMax.
__device__ __forceinline__ unsigned long long detect_key(float v, int m) {
    uint32_t u = __float_as_uint(v);
    u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;              // monotone float -> uint
    return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)m);
}
// The timeline stamps and debug switches are compiled in only with -DTSR_DEBUG_TRACE (TSR_NVCC_EXTRA=-DTSR_DEBUG_TRACE
// python template-speech-recognition_b200/build.py --force; tools/trace_detect.py needs it): even predicated-off they
// sit in the issue loops, where every instruction shows up in the kernel time.
#ifdef TSR_DEBUG_TRACE
#define DT_DBG(x) (A.dbg & (x))
#define DT_TRACE_ITERS 256
#define DT_TRACE(itv, slot)                                                                          \
    do {                                                                                             \
        if (A.trace && blockIdx.x == 0 && (itv) < DT_TRACE_ITERS && lane == 0) A.trace[(itv) * 8 + (slot)] = clock64(); \
    } while (0)

// debug: phase stamps of the drain passes of step 1, rows 192.. of the trace buffer
#define DT_TRACE_PASS(slot)                                                                                        \
    do {                                                                                                           \
        if (A.trace && blockIdx.x == 0 && stepno == 1 && quad == 0 && lane == 0 && passno < 60)                    \
            A.trace[(192 + passno) * 8 + (slot)] = clock64();                                                      \
    } while (0)
#else
#define DT_DBG(x) 0
#define DT_TRACE_ITERS 256
#define DT_TRACE(itv, slot) do { } while (0)
#define DT_TRACE_PASS(slot) do { } while (0)
#endif

struct DetectSmem {
    uint32_t b, stage, outacc, bars, stage_bytes, a_bytes, a_sub, strip_bytes, total;
};
// F16: nstage stages of [msub A sub-tiles | filter tile].  I8: two A strips, then a ring of nstage filter tiles.
__host__ __device__ inline DetectSmem detect_smem_layout(const FilterLayout& f) {
    DetectSmem S;
    S.a_sub = 16384;   // F16: 128 rows x 128 bytes per sub-tile and stage
    if (f.kind) {
        S.strip_bytes = 8u * (uint32_t)f.kgs;
        S.a_bytes = 2 * S.strip_bytes;
        S.stage_bytes = (uint32_t)f.tile_bytes;
        S.b = S.a_bytes;                               // start of the tile ring
        S.stage = S.b + f.nstage * S.stage_bytes;
    } else {
        S.strip_bytes = 0;
        S.a_bytes = f.msub * S.a_sub;
        S.stage_bytes = S.a_bytes + (uint32_t)f.tile_bytes;
        S.b = S.a_bytes;                               // within a stage
        S.stage = f.nstage * S.stage_bytes;            // epilogue staging [128][TC+1] words
    }
    S.stage = (S.stage + 127) & ~127u;
    S.outacc = S.stage + (((128 + DT_TAU_CHUNK) * (DT_TAU_CHUNK + 1) * 4 + 127) & ~127u);
    S.bars = S.outacc + (((f.msub * 128 + f.Lc + 1) * (f.tptv | 1) * 4 + 127) & ~127u);
    S.total = S.bars + 256;
    return S;
}

// 16 bits -> 16 bytes in {0,1} (uint8 A operand of the I8 path)
__device__ __forceinline__ uint4 expand16_u8(uint32_t b) {
    uint4 r;
    r.x = ((b & 0xFu) * 0x00204081u) & 0x01010101u;
    r.y = (((b >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
    r.z = (((b >> 8) & 0xFu) * 0x00204081u) & 0x01010101u;
    r.w = (((b >> 12) & 0xFu) * 0x00204081u) & 0x01010101u;
    return r;
}

// Unit u -> (N tile, row unit).  Units are ordered (tile group, row unit, tile within group) so that the CTAs
// running at any one time share the filter tiles of a few N tiles (they stay L2 resident: with many
// templates the whole filter bank is far larger than the L2) while the frame bits stream once per group.
__device__ __forceinline__ void detect_unit_coords(long long u, int ntiles, int U, int G, int& nt, long long& ur) {
    const long long per_group = (long long)U * G;
    const long long ng = u / per_group, rem = u - ng * per_group;
    const int left = ntiles - (int)ng * G;
    const int gt = left < G ? left : G;
    ur = rem / gt;
    nt = (int)ng * G + (int)(rem - ur * gt);
}

// NIS = MMA issuer warps (I8 only): a compile-time constant, because the issue loop is the critical path of the kernel and
// a run-time stride / barrier count in it cost 6-20 % (measured)
// UNI: take the warp index through a broadcast shuffle (uniform issue loop).  It wins everywhere except for L = 1 (the EM
// contractions), where the frame-bit expanders set the pace and the extra spills it causes there cost 5 %.  UNI = false is
// the L = 1 variant: plain warp index and expander groups that take alternate strips (see the expander role).
template <bool I8, int NIS, bool UNI>
__global__ void __launch_bounds__(448 + 32 * (NIS - 1), 1) detect_tc_kernel(DetectArgs A) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const FilterLayout& f = A.f;
    const DetectSmem S = detect_smem_layout(f);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S.bars);   // [DT_MAX_BSTAGES]  (I8: filter tile landed)
    uint64_t* empty_bar = full_bar + DT_MAX_BSTAGES;                   // [DT_MAX_BSTAGES]
    uint64_t* afull_bar = empty_bar + DT_MAX_BSTAGES;                  // [2]  I8: A strip expanded
    uint64_t* aempty_bar = afull_bar + 2;                              // [2]
    uint64_t* tfull_bar = aempty_bar + 2;                              // [1]
    uint64_t* tempty_bar = tfull_bar + 1;                              // [DT_MSUB_MAX]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + DT_MSUB_MAX);

    // warp index through a broadcast shuffle: tells the compiler it is warp-uniform, so the role branches are uniform and the
    // issuer's descriptor arithmetic can stay in uniform registers (fewer R2UR moves on the critical issue path)
    const int warp = UNI ? __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0) : (int)(threadIdx.x >> 5), lane = threadIdx.x & 31;
    const uint32_t fmt = reinterpret_cast<const uint32_t*>(A.blob)[0];
    const int L = f.L, NT = f.NT, msub = f.msub, nchunks = f.nchunks, tptv = f.tptv, nstage = f.nstage;
    const int step_rows = msub * 128;
    const long long nunits = (long long)f.ntiles * A.units_per_ntile;

    if (threadIdx.x == 0) {
        // I8: two MMA issuer warps (13, 14) share every pipeline slot, so the consumer-side barriers take two commits
        constexpr uint32_t nis = NIS;
        for (int s = 0; s < DT_MAX_BSTAGES; ++s) { mbar_init(&full_bar[s], I8 ? 1 : DT_PRODUCERS / 2 + 1); mbar_init(&empty_bar[s], nis); }
        for (int s = 0; s < 2; ++s) { mbar_init(&afull_bar[s], (I8 && !UNI) ? DT_PRODUCERS / 2 : DT_PRODUCERS); mbar_init(&aempty_bar[s], nis); }
        mbar_init(tfull_bar, nis);
        for (int s = 0; s < DT_MSUB_MAX; ++s) mbar_init(&tempty_bar[s], 128);
        mbar_fence_init();
    }
    if (warp == 13) tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp >= 1 && warp <= 8) {
        // ============================ frame-bit expanders ============================
        // Two groups of 4 warps take alternate pipeline stages (group = stage parity); a thread expands
        // one row of each 128-row sub-tile per stage it owns.  The proxy fence + mbarrier arrive that end a
        // stage cost several hundred cycles of latency; alternating groups overlaps them with the other
        // group's expansion work.
        if constexpr (I8) {
            // I8: all 8 warps expand one A strip per 128-feature frame chunk: rows = the step's frames plus
            // fold-1 more, 8 k-groups of 16 bytes each, k-group stride kgs.  Row r of phase phi's A tile is
            // strip row r + phi, so the strip is expanded once and read by all `fold` phases.
            // ALT (the L = 1 / EM variant): the two groups of 4 warps take alternate strips, group g always strip buffer g.
            // With L = 1 a strip feeds only 8 MMAs, and the chain wait -> expand -> proxy fence -> arrive of one strip
            // (~2 000 cycles, mostly latency) set the kernel's pace when all 8 warps worked on the same strip.
            constexpr bool ALT = !UNI;
            constexpr int XT = ALT ? 128 : 256;              // threads per strip
            const int grp = (threadIdx.x - 32) >> 7;         // ALT: group of this thread
            const int tid = ALT ? (int)((threadIdx.x - 32) & 127) : (int)(threadIdx.x - 32);
            const int nw2 = f.rowwords >> 1;                 // 64-bit words per frame
            const int srows = step_rows + f.fold - 1;        // live strip rows
            constexpr int RPT = (DT_MSUB_MAX * 128 + XT - 1) / XT + 1;   // strip rows per thread (<= 384 + fold - 1 rows)
            uint32_t ita = 0;
            for (long long u = blockIdx.x; u < nunits; u += gridDim.x) {
                int nt_unused; long long ur;
                detect_unit_coords(u, f.ntiles, A.units_per_ntile, A.ntile_group, nt_unused, ur);
                const long long o0 = ur * A.rows_per_unit;
                long long o1 = o0 + A.rows_per_unit; if (o1 > A.R) o1 = A.R;
                long long in_end = o1 + L - 1; if (in_end > A.R) in_end = A.R;
                for (int step = 0; step < A.steps_per_unit; ++step) {
                    const long long g0 = o0 + (long long)step * step_rows;
                    for (int cf = 0; cf < f.ncf; ++cf, ++ita) {
                        if (ALT && (int)(ita & 1) != grp) continue;
                        uint2 w[RPT], wh[RPT];
#pragma unroll
                        for (int k = 0; k < RPT; ++k) {
                            const int r = tid + XT * k;
                            const long long fr = g0 + r;
                            const bool ok = r < srows && fr < in_end;
                            const uint2* src = reinterpret_cast<const uint2*>(A.bits + (ok ? fr : 0) * f.rowwords);
                            w[k] = (ok && 2 * cf < nw2) ? __ldg(src + 2 * cf) : make_uint2(0u, 0u);
                            wh[k] = (ok && 2 * cf + 1 < nw2) ? __ldg(src + 2 * cf + 1) : make_uint2(0u, 0u);
                        }
                        const int buf = ita & 1;
                        mbar_wait_relaxed(&aempty_bar[buf], ((ita >> 1) & 1) ^ 1);
                        if (warp == 1) DT_TRACE(ita, 0);
                        uint8_t* strip = smem + buf * S.strip_bytes;
#pragma unroll
                        for (int k = 0; k < RPT; ++k) {
                            const int r = tid + XT * k;
                            if (r < f.strip_rows) {
                                uint8_t* dst = strip + r * 16;
                                const uint32_t kgs = (uint32_t)f.kgs;
                                *reinterpret_cast<uint4*>(dst) = expand16_u8(w[k].x & 0xFFFF);
                                *reinterpret_cast<uint4*>(dst + kgs) = expand16_u8(w[k].x >> 16);
                                *reinterpret_cast<uint4*>(dst + 2 * kgs) = expand16_u8(w[k].y & 0xFFFF);
                                *reinterpret_cast<uint4*>(dst + 3 * kgs) = expand16_u8(w[k].y >> 16);
                                *reinterpret_cast<uint4*>(dst + 4 * kgs) = expand16_u8(wh[k].x & 0xFFFF);
                                *reinterpret_cast<uint4*>(dst + 5 * kgs) = expand16_u8(wh[k].x >> 16);
                                *reinterpret_cast<uint4*>(dst + 6 * kgs) = expand16_u8(wh[k].y & 0xFFFF);
                                *reinterpret_cast<uint4*>(dst + 7 * kgs) = expand16_u8(wh[k].y >> 16);
                            }
                        }
                        fence_proxy_async_smem();
                        mbar_arrive(&afull_bar[buf]);
                        if (warp == 1) DT_TRACE(ita, 1);
                    }
                }
            }
        } else {
        const int grp = (warp - 1) / 4;                  // 0 or 1
        const int tg = threadIdx.x - 32 - grp * 128;     // 0..127 within the group == row of a sub-tile
        const uint32_t one = fmt_one(fmt);
        const int nw2 = f.rowwords >> 1;                  // 64-bit words per row
        const int w2ps = I8 ? 2 : 1;                      // 64-bit words per stage
        uint32_t it = 0;                                  // global chunk counter (pipeline position)
        int stage = 0; uint32_t ph = 0;
        for (long long u = blockIdx.x; u < nunits; u += gridDim.x) {
            int nt_unused; long long ur;
            detect_unit_coords(u, f.ntiles, A.units_per_ntile, A.ntile_group, nt_unused, ur);
            const long long o0 = ur * A.rows_per_unit;
            long long o1 = o0 + A.rows_per_unit; if (o1 > A.R) o1 = A.R;
            long long in_end = o1 + L - 1; if (in_end > A.R) in_end = A.R;
            for (int step = 0; step < A.steps_per_unit; ++step) {
                long long grow[DT_MSUB_MAX];
#pragma unroll
                for (int h = 0; h < DT_MSUB_MAX; ++h)                              // row tg of sub-tile h
                    grow[h] = h < msub ? o0 + (long long)step * step_rows + h * 128 + tg : in_end;   // in_end: reads as zeros
                // 64-bit word w of super-row g lives in frame g + w / nw2 at word w % nw2 (frames beyond the
                // unit's input range read as zero: they only meet zero weights or invalid outputs)
                auto fetch = [&](int h, int wq) -> uint2 {
                    const int phi = wq / nw2;
                    const long long fr = grow[h] + phi;
                    return (phi < f.fold && fr < in_end)
                               ? __ldg(reinterpret_cast<const uint2*>(A.bits + fr * f.rowwords) + (wq - phi * nw2))
                               : make_uint2(0u, 0u);
                };
                for (int c = 0; c < nchunks; ++c, ++it) {
                    if ((int)(it & 1) == grp) {
                        uint2 w[DT_MSUB_MAX], wh[DT_MSUB_MAX];
#pragma unroll
                        for (int h = 0; h < DT_MSUB_MAX; ++h) {
                            if (h < msub) {
                                w[h] = fetch(h, c * w2ps);
                                wh[h] = I8 ? fetch(h, c * w2ps + 1) : make_uint2(0u, 0u);
                            }
                        }
                        mbar_wait_relaxed(&empty_bar[stage], ph ^ 1);
                        if (warp == 1 || warp == 5) DT_TRACE(it, 0);
#pragma unroll
                        for (int h = 0; h < DT_MSUB_MAX; ++h) {
                            if (h < msub) {
                                uint8_t* dst = smem + stage * S.stage_bytes + h * S.a_sub + tg * 16;
                                if (I8) {   // 128 features -> 8 k-groups of 16 uint8
                                    *reinterpret_cast<uint4*>(dst) = expand16_u8(w[h].x & 0xFFFF);
                                    *reinterpret_cast<uint4*>(dst + 2048) = expand16_u8(w[h].x >> 16);
                                    *reinterpret_cast<uint4*>(dst + 4096) = expand16_u8(w[h].y & 0xFFFF);
                                    *reinterpret_cast<uint4*>(dst + 6144) = expand16_u8(w[h].y >> 16);
                                    *reinterpret_cast<uint4*>(dst + 8192) = expand16_u8(wh[h].x & 0xFFFF);
                                    *reinterpret_cast<uint4*>(dst + 10240) = expand16_u8(wh[h].x >> 16);
                                    *reinterpret_cast<uint4*>(dst + 12288) = expand16_u8(wh[h].y & 0xFFFF);
                                    *reinterpret_cast<uint4*>(dst + 14336) = expand16_u8(wh[h].y >> 16);
                                } else {    // 64 features -> 8 k-groups of 8 16-bit floats
#pragma unroll
                                    for (int kg = 0; kg < 4; ++kg)
                                        *reinterpret_cast<uint4*>(dst + kg * 2048) = expand_byte((w[h].x >> (8 * kg)) & 0xFF, one);
#pragma unroll
                                    for (int kg = 0; kg < 4; ++kg)
                                        *reinterpret_cast<uint4*>(dst + (kg + 4) * 2048) = expand_byte((w[h].y >> (8 * kg)) & 0xFF, one);
                                }
                            }
                        }
                        fence_proxy_async_smem();
                        mbar_arrive(&full_bar[stage]);
                        if (warp == 1 || warp == 5) DT_TRACE(it, 1);
                    }
                    if (++stage == nstage) { stage = 0; ph ^= 1; }
                }
            }
        }
        }   // F16 expanders
    } else if (warp == 0) {
        // ============================ filter-tile loader ============================
        if (lane == 0) {
            int stage = 0; uint32_t ph = 0;
            uint32_t it = 0;
            for (long long u = blockIdx.x; u < nunits; u += gridDim.x) {
                int nt; long long ur_unused;
                detect_unit_coords(u, f.ntiles, A.units_per_ntile, A.ntile_group, nt, ur_unused);
                const uint8_t* src = A.blob + f.hdr_bytes + (size_t)nt * nchunks * f.tile_bytes;
                for (int step = 0; step < A.steps_per_unit; ++step)
                    for (int c = 0; c < nchunks; ++c, ++it) {
                        mbar_wait_relaxed(&empty_bar[stage], ph ^ 1);
                        DT_TRACE(it, 2);
                        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)f.tile_bytes);
                        bulk_g2s(smem + (I8 ? S.b + stage * S.stage_bytes : stage * S.stage_bytes + S.b),
                                 src + (size_t)c * f.tile_bytes, (uint32_t)f.tile_bytes, &full_bar[stage]);
                        if (++stage == nstage) { stage = 0; ph ^= 1; }
                    }
            }
        }
    } else if (warp == 13 || (NIS == 2 && warp == 14)) {
        // ================================ MMA issuers ================================
        // warp-uniform loop, one elected lane issues (see umma_f16_elect).  One thread gets a tcgen05.mma out only
        // every ~75-90 cycles plus the barrier work between filter tiles (tools/ts_probe.cu), which left the I8
        // kernel at 141 cycles per 120-cycle MMA: two warps on different SMSPs each issue the MMAs of their own row
        // sub-tiles (sub = warp - 13, +2, ...) and both commit to every barrier.
        const int iss = warp - 13;
        const uint32_t idesc = I8 ? ((2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24))
                                  : umma_idesc_f32acc(fmt, 128, NT);
        const uint32_t desc_hi = (uint32_t)(128 >> 4) | (1u << 14);                 // SBO = 128 B, version 1
        const uint64_t a_hi = ((uint64_t)desc_hi << 32) | ((uint64_t)(2048 >> 4) << 16);   // LBO = 2048 B
        const uint64_t b_hi = ((uint64_t)desc_hi << 32) | ((uint64_t)NT << 16);            // LBO = NT*16 B
        int stage = 0; uint32_t ph = 0;
        uint32_t stepno = 0, it = 0;
        if constexpr (I8) {
            // per frame chunk cf: wait for its A strip, then fold x 2 filter tiles (64 features = 2 MMAs of
            // K = 32 per sub-tile each); phase phi reads the strip `phi` rows further down
            const uint64_t a_hi8 = ((uint64_t)desc_hi << 32) | ((uint64_t)((uint32_t)f.kgs >> 4) << 16);   // LBO = k-group stride
            const uint32_t kgs16 = (uint32_t)f.kgs >> 4;
            uint32_t ita = 0;
            for (long long u = blockIdx.x; u < nunits; u += gridDim.x)
                for (int step = 0; step < A.steps_per_unit; ++step, ++stepno) {
                    for (int cf = 0; cf < f.ncf; ++cf, ++ita) {
                        const int buf = ita & 1;
                        mbar_wait(&afull_bar[buf], (ita >> 1) & 1);
                        tc_fence_after();
                        const uint32_t sbase = smem_u32(smem + buf * S.strip_bytes) >> 4;
                        for (int phi = 0; phi < f.fold; ++phi)
                            for (int half = 0; half < 2; ++half, ++it) {
                                mbar_wait(&full_bar[stage], ph);
                                if (iss == 0) DT_TRACE(it, 4);
                                tc_fence_after();
                                const uint32_t bbase = smem_u32(smem + S.b + stage * S.stage_bytes) >> 4;
                                for (int sub = iss; sub < msub; sub += NIS) {
                                    if ((cf | phi | half) == 0) {
                                        mbar_wait(&tempty_bar[sub], (stepno & 1) ^ 1);
                                        tc_fence_after();
                                    }
                                    const uint32_t d_tmem = tmem_base + sub * NT;
                                    const uint32_t arow = sbase + (uint32_t)(sub * 128 + (DT_DBG(1) ? 0 : phi));
#pragma unroll
                                    for (int j = 0; j < 2; ++j) {   // K = 32 per MMA: k-groups 2*(2*half+j), +1
                                        const uint64_t adesc = a_hi8 | (uint64_t)(arow + (uint32_t)(2 * (2 * half + j)) * kgs16);
                                        const uint64_t bdesc = b_hi | (uint64_t)(bbase + j * 2 * NT);
                                        umma_i8_elect(d_tmem, adesc, bdesc, idesc, (cf | phi | half | j) != 0 ? 1u : 0u);
                                    }
                                }
                                umma_commit_elect(&empty_bar[stage]);
                                if (iss == 0) DT_TRACE(it, 5);
                                if (++stage == nstage) { stage = 0; ph ^= 1; }
                            }
                        umma_commit_elect(&aempty_bar[buf]);
                    }
                    umma_commit_elect(tfull_bar);
                }
        } else
        for (long long u = blockIdx.x; u < nunits; u += gridDim.x)
            for (int step = 0; step < A.steps_per_unit; ++step, ++stepno) {
                for (int c = 0; c < nchunks; ++c, ++it) {
                    mbar_wait(&full_bar[stage], ph);
                    DT_TRACE(it, 4);
                    tc_fence_after();
                    const uint32_t abase = smem_u32(smem + stage * S.stage_bytes) >> 4;
                    const uint32_t bbase = abase + (S.b >> 4);
                    for (int sub = 0; sub < msub; ++sub) {
                        if (c == 0) {
                            mbar_wait(&tempty_bar[sub], (stepno & 1) ^ 1);
                            tc_fence_after();
                        }
                        const uint32_t d_tmem = tmem_base + sub * NT;
                        if (I8) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {   // K = 32 per MMA: 2 x 16-byte k-groups, 128 features per stage
                                const uint64_t adesc = a_hi | (uint64_t)(abase + sub * (16384 >> 4) + ks * (4096 >> 4));
                                const uint64_t bdesc = b_hi | (uint64_t)(bbase + ks * 2 * NT);
                                umma_i8_elect(d_tmem, adesc, bdesc, idesc, (c | ks) != 0 ? 1u : 0u);
                            }
                        } else {
#pragma unroll
                            for (int plane = 0; plane < 2; ++plane)
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) {
                                    const uint64_t adesc = a_hi | (uint64_t)(abase + sub * (16384 >> 4) + ks * (4096 >> 4));
                                    const uint64_t bdesc = b_hi | (uint64_t)(bbase + plane * NT * 8 + ks * 2 * NT);
                                    umma_f16_elect(d_tmem, adesc, bdesc, idesc, (c | plane | ks) != 0 ? 1u : 0u);
                                }
                        }
                    }
                    umma_commit_elect(&empty_bar[stage]);
                    DT_TRACE(it, 5);
                    if (++stage == nstage) { stage = 0; ph ^= 1; }
                }
                umma_commit_elect(tfull_bar);
            }
    } else if (warp >= 9 && warp <= 12) {
        // ================================= epilogue =================================
        // accumulators are fp32 (F16 kind) or int32 (I8 kind); staging and diagonal sums use the same type
        typedef typename std::conditional<I8, int, float>::type acc_t;
        constexpr int PITCH = DT_TAU_CHUNK + 1;                      // odd word pitch: conflict-free rows and diagonals
        acc_t* stg = reinterpret_cast<acc_t*>(smem + S.stage);       // skewed staging [128 + TC - 1][PITCH]
        acc_t* outacc = reinterpret_cast<acc_t*>(smem + S.outacc);   // [step_rows + Lc][opitch]
        const int opitch = tptv | 1;   // odd row pitch: the per-row accumulations of a warp (one row per lane) hit distinct banks
        const float* scale = reinterpret_cast<const float*>(A.blob + 128);
        const int quad = warp & 3;                                    // TMEM lane quadrant of this warp
        const int e = quad * 32 + lane;                               // 0..127 == TMEM lane == row of the sub-tile
        // Columns of a group are the folded frames tau' = 0..Lf-1; column tau' belongs to the out row
        // fold*tau' rows above.  They are staged TCf columns at a time, TCf chosen so that the skew
        // fold*(TCf-1) stays within the DT_TAU_CHUNK-1 extra staging rows.
        const int Lf = f.Lf, fold = f.fold, Lc = f.Lc;
        int TCf = (DT_TAU_CHUNK - 1) / fold + 1;
        if (TCf > Lf) TCf = Lf;
        const int skew = fold * (TCf - 1);                 // row shift of the last column of a chunk
        const int srows = 128 + skew;                       // staging rows in use
        const int nacc = (step_rows + Lc) * opitch;
        const int vpt = f.vpt, tpt = tptv / vpt;
        // Skewed staging: thread e writes C[e][c0+tr] to row q' = e - fold*tr + skew, column tr, so that a
        // diagonal of the tile becomes a ROW of the staging buffer and the diagonal sum a plain row sum (one
        // load + one add per element, no masks).  Cells no (e,tr) maps to -- the two triangles at the top
        // and bottom -- are zeroed once here and never written again.
        for (int i = e; i < (128 + DT_TAU_CHUNK) * PITCH; i += 128) stg[i] = 0;
        // the extra rows (q' >= 128) are spread over the four warps: lanes 0..9 of each take one
        const int xrow = lane < 10 ? 128 + quad * 10 + lane : -1;
        uint32_t stepno = 0;
        for (long long u = blockIdx.x; u < nunits; u += gridDim.x) {
            int nt; long long ur;
            detect_unit_coords(u, f.ntiles, A.units_per_ntile, A.ntile_group, nt, ur);
            const long long o0 = ur * A.rows_per_unit;
            long long o1 = o0 + A.rows_per_unit; if (o1 > A.R) o1 = A.R;
            for (int i = e; i < nacc; i += 128) outacc[i] = 0;
            named_bar_sync(1, 128);
            for (int step = 0; step < A.steps_per_unit; ++step, ++stepno) {
                mbar_wait_relaxed(tfull_bar, stepno & 1);
                if (quad == 0) DT_TRACE(stepno, 6);
                int passno = 0; (void)passno;
                tc_fence_after();
                for (int sub = 0; sub < msub; ++sub) {
                    const uint32_t taddr = tmem_base + sub * NT + ((uint32_t)(quad * 32) << 16);
                    if (Lc < DT_TAU_CHUNK) {
                        // ---- whole groups per pass: gp = floor(40 / Lf) groups = up to 40 TMEM columns ----
                        const int gp = DT_TAU_CHUNK / Lf;
                        for (int vl0 = 0; vl0 < tptv; vl0 += gp) {
                            const int g = (tptv - vl0) < gp ? (tptv - vl0) : gp;
                            const int cn = g * Lf;                                 // live columns of this pass
                            uint32_t v[DT_TAU_CHUNK];
#pragma unroll
                            for (int t0 = 0; t0 < DT_TAU_CHUNK; t0 += 8) {
                                if (t0 < cn) {                                     // warp-uniform
                                    tmem_ld8(taddr + vl0 * Lf + t0, *reinterpret_cast<uint32_t(*)[8]>(&v[t0]));
                                } else {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) v[t0 + i] = 0u;
                                }
                            }
                            DT_TRACE_PASS(0);
                            tmem_ld_wait();
                            DT_TRACE_PASS(1);
                            // column i (frame tr = i % Lf of group i / Lf) goes to staging row e + Lc - fold*tr.
                            // Compile-time Lf (the common fold choices) gives straight-line code with
                            // immediate offsets; other values take the counter-driven generic path.
                            auto store_skewed = [&](auto lfc) {
                                constexpr int LFC = decltype(lfc)::value;
                                uint32_t* sp = reinterpret_cast<uint32_t*>(stg) + (e + Lc) * PITCH;
                                if constexpr (LFC > 0) {
                                    const int rstep = fold * PITCH;
#pragma unroll
                                    for (int i = 0; i < DT_TAU_CHUNK; ++i)
                                        if ((i / LFC) * LFC + LFC <= DT_TAU_CHUNK)            // whole groups only (static)
                                            sp[i - (i % LFC) * rstep] = i < cn ? v[i] : 0u;
                                } else {
                                    int roff = 0, tr = 0;
#pragma unroll
                                    for (int i = 0; i < DT_TAU_CHUNK; ++i) {
                                        sp[roff + i] = i < cn ? v[i] : 0u;
                                        if (++tr == Lf) { tr = 0; roff = 0; } else { roff -= fold * PITCH; }
                                    }
                                }
                            };
                            switch (Lf) {
                                case 40: store_skewed(std::integral_constant<int, 40>{}); break;
                                case 20: store_skewed(std::integral_constant<int, 20>{}); break;
                                case 10: store_skewed(std::integral_constant<int, 10>{}); break;
                                case 8: store_skewed(std::integral_constant<int, 8>{}); break;
                                case 5: store_skewed(std::integral_constant<int, 5>{}); break;
                                case 4: store_skewed(std::integral_constant<int, 4>{}); break;
                                case 2: store_skewed(std::integral_constant<int, 2>{}); break;
                                case 1: store_skewed(std::integral_constant<int, 1>{}); break;
                                default: store_skewed(std::integral_constant<int, 0>{}); break;
                            }
                            if (vl0 + g >= tptv) {   // accumulator drained: hand it back to the MMA warp
                                tc_fence_before();
                                mbar_arrive(&tempty_bar[sub]);
                            }
                            DT_TRACE_PASS(2);
                            named_bar_sync(1, 128);
                            DT_TRACE_PASS(3);
                            // row q' holds, per group, the Lf terms of out index q' (staging row == out index)
                            auto sum_rows = [&](auto lfc) {
                                constexpr int LFC = decltype(lfc)::value;
                                const acc_t* r0p = stg + e * PITCH;
                                const acc_t* r1p = stg + (xrow < 0 ? 0 : xrow) * PITCH;
                                const bool x_ok = xrow >= 0 && xrow <= 128 + Lc;
                                acc_t* o0p = outacc + (sub * 128 + e) * opitch + vl0;
                                acc_t* o1p = outacc + (sub * 128 + (x_ok ? xrow : 0)) * opitch + vl0;
                                if constexpr (LFC > 0) {
                                    // all row sums, then all accumulator loads, then all stores: the read-modify-write
                                    // chains of the groups overlap instead of paying one smem round trip each
                                    constexpr int NG = DT_TAU_CHUNK / LFC;
                                    constexpr int BG = NG < 8 ? NG : 8;      // groups per batch (register budget)
#pragma unroll
                                    for (int g0 = 0; g0 < NG; g0 += BG) {
                                        acc_t s0[BG], s1[BG], a0[BG], a1[BG];
#pragma unroll
                                        for (int k = 0; k < BG; ++k) {
                                            s0[k] = 0; s1[k] = 0;
                                            if (g0 + k < NG) {
#pragma unroll
                                                for (int tr = 0; tr < LFC; ++tr) {
                                                    s0[k] += r0p[(g0 + k) * LFC + tr];
                                                    s1[k] += r1p[(g0 + k) * LFC + tr];
                                                }
                                            }
                                        }
#pragma unroll
                                        for (int k = 0; k < BG; ++k) {
                                            a0[k] = g0 + k < g ? o0p[g0 + k] : 0;
                                            a1[k] = (g0 + k < g && x_ok) ? o1p[g0 + k] : 0;
                                        }
#pragma unroll
                                        for (int k = 0; k < BG; ++k) {
                                            if (g0 + k < g) {
                                                o0p[g0 + k] = a0[k] + s0[k];
                                                if (x_ok) o1p[g0 + k] = a1[k] + s1[k];
                                            }
                                        }
                                    }
                                } else {
                                    acc_t s0 = 0, s1 = 0;
                                    int tr = 0, gi = 0;
#pragma unroll
                                    for (int i = 0; i < DT_TAU_CHUNK; ++i) {
                                        s0 += r0p[i]; s1 += r1p[i];
                                        if (++tr == Lf) {                              // end of a group: warp-uniform
                                            if (gi < g) { o0p[gi] += s0; if (x_ok) o1p[gi] += s1; }
                                            s0 = 0; s1 = 0; tr = 0; ++gi;
                                        }
                                    }
                                }
                            };
                            switch (Lf) {
                                case 40: sum_rows(std::integral_constant<int, 40>{}); break;
                                case 20: sum_rows(std::integral_constant<int, 20>{}); break;
                                case 10: sum_rows(std::integral_constant<int, 10>{}); break;
                                case 8: sum_rows(std::integral_constant<int, 8>{}); break;
                                case 5: sum_rows(std::integral_constant<int, 5>{}); break;
                                case 4: sum_rows(std::integral_constant<int, 4>{}); break;
                                case 2: sum_rows(std::integral_constant<int, 2>{}); break;
                                case 1: sum_rows(std::integral_constant<int, 1>{}); break;
                                default: sum_rows(std::integral_constant<int, 0>{}); break;
                            }
                            DT_TRACE_PASS(4);
                            named_bar_sync(1, 128);
                            DT_TRACE_PASS(5);
                            ++passno;
                        }
                    } else
                    for (int vl = 0; vl < tptv; ++vl) {
                        for (int c0 = 0; c0 < Lf; c0 += TCf) {         // chunk of columns [c0, c0+cl)
                            const int cl = (Lf - c0) < TCf ? (Lf - c0) : TCf;
                            // all TMEM loads of the chunk in flight at once, one wait; columns past the chunk
                            // end are zeroed by select so that the sums below need no bounds checks
                            uint32_t v[DT_TAU_CHUNK];
#pragma unroll
                            for (int t0 = 0; t0 < DT_TAU_CHUNK; t0 += 8) {
                                if (t0 < cl) {                                     // warp-uniform
                                    tmem_ld8(taddr + vl * Lf + c0 + t0, *reinterpret_cast<uint32_t(*)[8]>(&v[t0]));
                                } else {
#pragma unroll
                                    for (int i = 0; i < 8; ++i) v[t0 + i] = 0u;
                                }
                            }
                            tmem_ld_wait();
                            {
                                uint32_t* sp = reinterpret_cast<uint32_t*>(stg) + (e + skew) * PITCH;
                                const int dstep = 1 - fold * PITCH;                 // next column, `fold` rows up
#pragma unroll
                                for (int i = 0; i < DT_TAU_CHUNK; ++i)
                                    if (i < TCf) sp[i * dstep] = i < cl ? v[i] : 0u;   // warp-uniform guard
                            }
                            if (vl == tptv - 1 && c0 + cl >= Lf) {   // accumulator drained: hand it back to the MMA warp
                                tc_fence_before();
                                mbar_arrive(&tempty_bar[sub]);
                            }
                            named_bar_sync(1, 128);
                            // row q' of the staging buffer = sum over the chunk's columns of C[q'+fold*tr-skew][c0+tr]:
                            // the partial diagonal sum of out index q = q' + (Lc - skew - fold*c0).
                            {
                                const int qoff = Lc - skew - fold * c0;
                                acc_t s0 = 0, s1 = 0;
                                const acc_t* r0p = stg + e * PITCH;
                                const acc_t* r1p = stg + (xrow < 0 ? 0 : xrow) * PITCH;
#pragma unroll
                                for (int tr = 0; tr < DT_TAU_CHUNK; ++tr)
                                    if (tr < TCf) { s0 += r0p[tr]; s1 += r1p[tr]; }   // warp-uniform guard
                                const int q0 = e + qoff;
                                if (q0 >= 0) outacc[(sub * 128 + q0) * opitch + vl] += s0;
                                const int q1 = xrow + qoff;
                                if (xrow >= 0 && xrow < srows && q1 >= 0) outacc[(sub * 128 + q1) * opitch + vl] += s1;
                            }
                            named_bar_sync(1, 128);
                        }
                    }
                }
                // rows [0, step_rows) of outacc are complete: out row g = step_row0 - Lc + oi
                const long long gbase = o0 + (long long)step * step_rows - Lc;
                for (int i = e; i < step_rows * tpt; i += 128) {
                    const int oi = i / tpt, ml = i - oi * tpt;
                    const long long g = gbase + oi;
                    const int m = nt * tpt + ml;
                    if (A.scores && g >= o0 && g < o1 && m < f.M) {
                        const int tl = (int)(g % A.T);
                        float val = 0.f;
                        if (tl <= A.T - L) {
                            const acc_t* a = outacc + oi * opitch + ml * vpt;
                            if (I8) {   // recombine the base-256 digit planes exactly, then scale once
                                double q = (double)a[0] + 256.0 * (double)a[1];
                                if (vpt == 3) q += 65536.0 * (double)a[2];
                                val = (float)(q * (double)scale[m]);
                            } else {
                                val = (float)a[0];
                            }
                        }
                        A.scores[g * f.M + m] = val;
                    }
                }
                if (A.keys) {
                    // fused mixture max: a thread walks the templates of this tile for its rows (odd accumulator
                    // pitch: the lanes' rows sit in different banks) and publishes one key per row
                    for (int oi = e; oi < step_rows; oi += 128) {
                        const long long g = gbase + oi;
                        if (g >= o0 && g < o1) {
                            const bool valid = (int)(g % A.T) <= A.T - L;
                            unsigned long long key = 0;
                            for (int ml = 0; ml < tpt; ++ml) {
                                const int m = nt * tpt + ml;
                                if (m >= f.M) break;
                                float val = 0.f;
                                if (valid) {
                                    const acc_t* a = outacc + oi * opitch + ml * vpt;
                                    if (I8) {
                                        double q = (double)a[0] + 256.0 * (double)a[1];
                                        if (vpt == 3) q += 65536.0 * (double)a[2];
                                        val = (float)(q * (double)scale[m]);
                                    } else {
                                        val = (float)a[0];
                                    }
                                }
                                const unsigned long long k2 = detect_key(val + (A.cs ? __ldg(A.cs + m) : 0.f), m);
                                key = k2 > key ? k2 : key;
                            }
                            atomicMax(A.keys + g, key);
                        }
                    }
                }
                named_bar_sync(1, 128);
                // carry the Lc unfinished rows to the front and clear the rest.  dst = src - shift; ascending
                // passes of `shift` elements never read what an earlier pass wrote, one barrier per pass.
                {
                    const int keep = Lc * opitch, shift = step_rows * opitch;
                    for (int p0 = 0; p0 < keep; p0 += shift) {
                        const int p1 = (p0 + shift) < keep ? (p0 + shift) : keep;
                        for (int i = p0 + e; i < p1; i += 128) outacc[i] = outacc[i + shift];
                        named_bar_sync(1, 128);
                    }
                    for (int i = keep + e; i < nacc; i += 128) outacc[i] = 0;
                    named_bar_sync(1, 128);
                }
                if (quad == 0) DT_TRACE(stepno, 7);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 13) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------ TN kernel
// detect_tn_kernel (TSR_PLANES_INT8X2): the same contraction with the operand roles swapped,
//   D[(v,tau') : 128 TMEM lanes][frame g : <= 256 TMEM columns] += Wq[(v,tau')][k] * bit[g + phi][k],
// so that the diagonal of the sliding window runs along TMEM *columns*:  scores[g][v] = sum_tau' D[(v,tau')][g + fold*tau'].
// A group's Lf columns are Lf adjacent lanes (Lf a power of two), so an epilogue thread loads its own lane at the column
// offsets fold*tau', keeps the one that belongs to its tau' and a butterfly of log2(Lf) shuffles finishes the sum: no
// shared-memory staging, no per-step accumulator matrix.  What the 2 x 240-column accumulators of detect_tc_kernel could
// not have follows from that: one 128-lane x 256-column accumulator per step leaves room for a second one, the epilogue
// of step s runs while the MMAs of step s+1 fill the other (in detect_tc_kernel the tensor pipe idles ~12 % of every
// step while both accumulators drain), and ~100 KB of shared memory go to a deeper filter-tile ring instead.
//   TMEM: accumulator a = step parity at columns [256 a, 256 a + step_rows).  The epilogue zeroes an accumulator after
//   reading it (tcgen05.st) and every MMA accumulates, so the two issuer warps can take alternate filter tiles without
//   ordering their first MMAs.
//   Unfinished frames (the last Lc = fold*(Lf-1) of a step) leave their partial sums in a thread-private shared-memory
//   slot [Lc][128] and are completed by the same thread in the next step.
//   step_rows (UMMA N, a multiple of 16) is picked by the host so that the units' steps are evenly filled.
struct TnSmem { uint32_t strip_bytes, b, carry, bars, total; };
__host__ __device__ inline TnSmem tn_smem_layout(const FilterLayout& f) {
    TnSmem S;
    S.strip_bytes = 8u * (uint32_t)f.kgs;
    S.b = 2 * S.strip_bytes;
    S.carry = S.b + (uint32_t)f.nstage * 2u * (uint32_t)f.tile_bytes;
    S.bars = (S.carry + (uint32_t)f.Lc * 128u * 4u + 127u) & ~127u;
    S.total = S.bars + 512;
    return S;
}

__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8_zero(uint32_t taddr) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st_wait_all() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// x holds the 8 columns loaded at a clamped position; d = wanted start - loaded start.  Columns outside the slot read as 0.
__device__ __forceinline__ void tn_shift_window(uint32_t (&x)[8], int d) {
#pragma unroll
    for (int sft = 1; sft < 8; ++sft) {
        if (d == -sft) {
#pragma unroll
            for (int i = 7; i >= 0; --i) x[i] = i >= sft ? x[i - sft] : 0u;
        }
        if (d == sft) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = i + sft < 8 ? x[i + sft] : 0u;
        }
    }
}

template <int NIS>
__global__ void __launch_bounds__(448 + 32 * (NIS - 1), 1) detect_tn_kernel(DetectArgs A) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const FilterLayout& f = A.f;
    const TnSmem S = tn_smem_layout(f);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S.bars);   // [DT_MAX_BSTAGES] filter tile landed
    uint64_t* empty_bar = full_bar + DT_TN_MAX_STAGES;
    uint64_t* afull_bar = empty_bar + DT_TN_MAX_STAGES;                  // [2] strip expanded
    uint64_t* aempty_bar = afull_bar + 2;                              // [2]
    uint64_t* tfull_bar = aempty_bar + 2;                              // [2] accumulator complete
    uint64_t* tempty_bar = tfull_bar + 2;                              // [2] accumulator drained and zeroed
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int L = f.L, nstage = f.nstage, fold = f.fold;
    const int step_rows = A.step_rows;
    const long long nunits = (long long)f.ntiles * A.units_per_ntile;
    // Integer accumulation is order independent, so every CTA walks the contraction from its own starting chunk: at any
    // moment the SMs read different filter tiles (all CTAs fetching the same 8 KB at the same time load a few L2 slices
    // with 148 requests per line while the others idle).
    const int rot = DT_DBG(32) ? 0 : (int)(blockIdx.x % (unsigned)f.ncf);

    if (threadIdx.x == 0) {
        for (int s = 0; s < DT_TN_MAX_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NIS); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&afull_bar[s], DT_PRODUCERS); mbar_init(&aempty_bar[s], NIS);
            mbar_init(&tfull_bar[s], NIS); mbar_init(&tempty_bar[s], 128);
        }
        mbar_fence_init();
    }
    if (warp == 13) tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp >= 1 && warp <= 8) {
        // ============================ frame-bit expanders ============================
        // one strip per 128-feature chunk of a frame: rows = the step's frames plus fold-1 more, 8 k-groups of 16 bytes
        // (k-group stride kgs); phase phi of the contraction reads it `phi` rows further down
        const int tid = (int)threadIdx.x - 32;
        const int nw2 = f.rowwords >> 1;
        const int srows = step_rows + fold - 1;
        constexpr int RPT = 2;                       // strip_rows <= 256 + 255
        uint32_t ita = 0;
        for (long long u = blockIdx.x; u < nunits; u += gridDim.x) {
            int nt_unused; long long ur;
            detect_unit_coords(u, f.ntiles, A.units_per_ntile, A.ntile_group, nt_unused, ur);
            const long long o0 = ur * A.rows_per_unit;
            long long o1 = o0 + A.rows_per_unit; if (o1 > A.R) o1 = A.R;
            long long in_end = o1 + L - 1; if (in_end > A.R) in_end = A.R;
            for (int step = 0; step < A.steps_per_unit; ++step) {
                const long long g0 = o0 + (long long)step * step_rows;
                for (int cfi = 0; cfi < f.ncf; ++cfi, ++ita) {
                    int cf = cfi + rot; if (cf >= f.ncf) cf -= f.ncf;
                    uint2 w[RPT], wh[RPT];
#pragma unroll
                    for (int k = 0; k < RPT; ++k) {
                        const int r = tid + 256 * k;
                        const long long fr = g0 + r;
                        const bool ok = r < srows && fr < in_end;
                        const uint2* src = reinterpret_cast<const uint2*>(A.bits + (ok ? fr : 0) * f.rowwords);
                        w[k] = (ok && 2 * cf < nw2) ? __ldg(src + 2 * cf) : make_uint2(0u, 0u);
                        wh[k] = (ok && 2 * cf + 1 < nw2) ? __ldg(src + 2 * cf + 1) : make_uint2(0u, 0u);
                    }
                    const int buf = ita & 1;
                    mbar_wait_relaxed(&aempty_bar[buf], ((ita >> 1) & 1) ^ 1);
                    uint8_t* strip = smem + buf * S.strip_bytes;
#pragma unroll
                    for (int k = 0; k < RPT; ++k) {
                        const int r = tid + 256 * k;
                        if (r < f.strip_rows && !DT_DBG(8)) {
                            uint8_t* dst = strip + r * 16;
                            const uint32_t kgs = (uint32_t)f.kgs;
                            *reinterpret_cast<uint4*>(dst) = expand16_u8(w[k].x & 0xFFFF);
                            *reinterpret_cast<uint4*>(dst + kgs) = expand16_u8(w[k].x >> 16);
                            *reinterpret_cast<uint4*>(dst + 2 * kgs) = expand16_u8(w[k].y & 0xFFFF);
                            *reinterpret_cast<uint4*>(dst + 3 * kgs) = expand16_u8(w[k].y >> 16);
                            *reinterpret_cast<uint4*>(dst + 4 * kgs) = expand16_u8(wh[k].x & 0xFFFF);
                            *reinterpret_cast<uint4*>(dst + 5 * kgs) = expand16_u8(wh[k].x >> 16);
                            *reinterpret_cast<uint4*>(dst + 6 * kgs) = expand16_u8(wh[k].y & 0xFFFF);
                            *reinterpret_cast<uint4*>(dst + 7 * kgs) = expand16_u8(wh[k].y >> 16);
                        }
                    }
                    fence_proxy_async_smem();
                    mbar_arrive(&afull_bar[buf]);
                }
            }
        }
    } else if (warp == 0) {
        // ============================ filter-tile loader ============================
        if (lane == 0) {
            int stage = 0; uint32_t ph = 0;
            for (long long u = blockIdx.x; u < nunits; u += gridDim.x) {
                int nt; long long ur_unused;
                detect_unit_coords(u, f.ntiles, A.units_per_ntile, A.ntile_group, nt, ur_unused);
                const uint8_t* src = A.blob + f.hdr_bytes + (size_t)nt * f.nchunks * f.tile_bytes;
                // One thread pays ~300 cycles of dependent latencies per copy (try_wait, expect_tx, UBLKCP): a stage is
                // therefore the 16 KB pair of tiles of one phase (480 cycles of MMAs), not a single 8 KB tile.
                const uint32_t sbytes = 2u * (uint32_t)f.tile_bytes;
                const int nst = f.ncf * fold;
                for (int step = 0; step < A.steps_per_unit; ++step)
                    for (int ci = 0; ci < nst; ++ci) {
                        int c = ci + rot * fold; if (c >= nst) c -= nst;
                        mbar_wait_relaxed(&empty_bar[stage], ph ^ 1);
                        if (DT_DBG(4)) { mbar_arrive(&full_bar[stage]); } else {
                        mbar_arrive_expect_tx(&full_bar[stage], sbytes);
                        bulk_g2s(smem + S.b + stage * sbytes, src + (size_t)c * sbytes, sbytes, &full_bar[stage]);
                        }
                        if (++stage == nstage) { stage = 0; ph ^= 1; }
                    }
            }
        }
    } else if (warp == 13 || (NIS == 2 && warp == 14)) {
        // ================================ MMA issuers ================================
        // A = filter tile (int8, 128 columns x 32 per MMA), B = strip rows phi .. phi + step_rows (uint8 {0,1}), D = int32.
        // With two issuers, issuer i takes the 64-feature tile i of every stage; every MMA accumulates (zeroed accumulators).
        const int iss = warp - 13;
        const uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(step_rows >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t desc_hi = (uint32_t)(128 >> 4) | (1u << 14);                                  // SBO = 128 B, version 1
        const uint64_t a_hi = ((uint64_t)desc_hi << 32) | ((uint64_t)128 << 16);                     // LBO = 128 rows * 16 B
        const uint64_t b_hi = ((uint64_t)desc_hi << 32) | ((uint64_t)((uint32_t)f.kgs >> 4) << 16);  // LBO = k-group stride
        const uint32_t kgs16 = (uint32_t)f.kgs >> 4;
        int stage = 0; uint32_t ph = 0;
        uint32_t stepno = 0, ita = 0;
        for (long long u = blockIdx.x; u < nunits; u += gridDim.x)
            for (int step = 0; step < A.steps_per_unit; ++step, ++stepno) {
                const uint32_t acc = stepno & 1;
                mbar_wait(&tempty_bar[acc], (stepno >> 1) & 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int cfi = 0; cfi < f.ncf; ++cfi, ++ita) {
                    int cf = cfi + rot; if (cf >= f.ncf) cf -= f.ncf;
                    // 32-feature K-steps of this chunk that hold features at all (the last chunk of D = 3600: 1 of 4)
                    const int feats = f.D - cf * 128;
                    const int ksteps = feats >= 128 ? 4 : (feats + 31) >> 5;
                    const int buf = ita & 1;
                    mbar_wait(&afull_bar[buf], (ita >> 1) & 1);
                    tc_fence_after();
                    const uint32_t sbase = smem_u32(smem + buf * S.strip_bytes) >> 4;
                    for (int phi = 0; phi < fold; ++phi) {
                        mbar_wait(&full_bar[stage], ph);
                        tc_fence_after();
                        const uint32_t abase = smem_u32(smem + S.b + stage * 2u * (uint32_t)f.tile_bytes) >> 4;
#pragma unroll
                        for (int half = (NIS == 2 ? iss : 0); half < 2; half += NIS) {   // two issuers: one 64-feature tile each
#pragma unroll
                            for (int j = 0; j < 2; ++j) {   // K = 32 per MMA: k-groups 2j, 2j+1 of the tile, 2(2 half + j), +1 of the strip
                                const uint64_t adesc = a_hi | (uint64_t)(abase + (uint32_t)half * 512u + (uint32_t)j * 2 * 128);
                                const uint64_t bdesc = b_hi | (uint64_t)(sbase + (uint32_t)(DT_DBG(1) ? 0 : phi) + (uint32_t)(2 * (2 * half + j)) * kgs16);
                                if (2 * half + j < ksteps) umma_i8_elect(d_tmem, adesc, bdesc, idesc, 1u);
                            }
                        }
                        umma_commit_elect(&empty_bar[stage]);
                        if (++stage == nstage) { stage = 0; ph ^= 1; }
                    }
                    umma_commit_elect(&aempty_bar[buf]);
                }
                umma_commit_elect(&tfull_bar[acc]);
            }
    } else if (warp >= 9 && warp <= 12) {
        // ================================= epilogue =================================
        const int quad = warp & 3;
        const int e = quad * 32 + lane;                       // TMEM lane == filter column of the tile
        const int Lf = f.Lf, Lc = f.Lc;
        const int tr = e & (Lf - 1), vl = e / Lf;             // tau', group
        const float* scale = reinterpret_cast<const float*>(A.blob + 128);
        int* carry = reinterpret_cast<int*>(smem + S.carry) + e;      // [Lc][128], column e is private to this thread
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        // both accumulators start zeroed
        for (int c = 0; c < 512; c += 8) tmem_st8_zero(tmem_base + lane_addr + c);
        tmem_st_wait_all();
        tc_fence_before();
        mbar_arrive(&tempty_bar[0]);
        mbar_arrive(&tempty_bar[1]);
        const bool small_r = A.R < (1ll << 31);
        const int Lc8 = (Lc + 7) & ~7;
        uint32_t stepno = 0;
        for (long long u = blockIdx.x; u < nunits; u += gridDim.x) {
            int nt; long long ur;
            detect_unit_coords(u, f.ntiles, A.units_per_ntile, A.ntile_group, nt, ur);
            const long long o0 = ur * A.rows_per_unit;
            long long o1 = o0 + A.rows_per_unit; if (o1 > A.R) o1 = A.R;
            const int v = nt * f.tptv + vl;
            const int m = v >> 1, plane = v & 1;
            const bool live = vl < f.tptv && m < f.M;
            const float sc_m = live ? scale[m] : 0.f;
            const float cs_m = (live && A.cs) ? __ldg(A.cs + m) : 0.f;
            for (int step = 0; step < A.steps_per_unit; ++step, ++stepno) {
                const uint32_t acc = stepno & 1;
                mbar_wait_relaxed(&tfull_bar[acc], (stepno >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + lane_addr + acc * 256;
                const long long F0 = o0 + (long long)step * step_rows;
                for (int c0 = -Lc8; c0 < step_rows; c0 += 8) {
                    if (DT_DBG(16)) break;
                    const long long gb = F0 + c0;
                    if (gb + 8 <= o0 || gb >= o1 + Lc) continue;                  // nothing owned or carried in this block
                    int val[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) val[i] = 0;
                    // window t = columns [c0 + fold t, +8) of this lane; the lane keeps the window of its own tau'.  Columns
                    // [step_rows, 256) of an accumulator slot are never written and read as zero; a window that sticks out
                    // of [0, 256) is loaded from the clamped position and shifted (boundary blocks only).
                    for (int t = 0; t < Lf; t += 2) {
                        uint32_t x0[8], x1[8];
                        const int col0 = c0 + fold * t, col1 = col0 + fold;
                        const bool h0 = col0 + 8 > 0 && col0 < step_rows;
                        const bool h1 = t + 1 < Lf && col1 + 8 > 0 && col1 < step_rows;
                        const int cc0 = col0 < 0 ? 0 : (col0 > 248 ? 248 : col0), cc1 = col1 < 0 ? 0 : (col1 > 248 ? 248 : col1);
                        if (h0) tmem_ld8(taddr + cc0, x0);
                        if (h1) tmem_ld8(taddr + cc1, x1);
                        if (h0 || h1) tmem_ld_wait();
                        if (h0) {
                            if (cc0 != col0) tn_shift_window(x0, col0 - cc0);
#pragma unroll
                            for (int i = 0; i < 8; ++i) val[i] = tr == t ? (int)x0[i] : val[i];
                        }
                        if (h1) {
                            if (cc1 != col1) tn_shift_window(x1, col1 - cc1);
#pragma unroll
                            for (int i = 0; i < 8; ++i) val[i] = tr == t + 1 ? (int)x1[i] : val[i];
                        }
                    }
                    for (int off = 1; off < Lf; off <<= 1) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) val[i] += __shfl_xor_sync(0xffffffffu, val[i], off);
                    }
                    int tl = 0;
                    {
                        const long long gq = gb < 0 ? 0 : gb;
                        tl = small_r ? (int)((unsigned)gq % (unsigned)A.T) : (int)(gq % A.T);
                        if (gb < 0) tl += (int)gb;                                   // only blocks that are skipped below
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int c = c0 + i;
                        const long long g = gb + i;
                        int s = val[i];
                        if (c < -Lc) continue;
                        if (c < 0) s += carry[(c + Lc) * 128];
                        if (c >= step_rows - Lc) { carry[(c - (step_rows - Lc)) * 128] = s; continue; }
                        if (g < o0 || g >= o1) continue;                            // warp-uniform
                        int tli = tl + i; while (tli >= A.T) tli -= A.T;
                        const int other = __shfl_xor_sync(0xffffffffu, s, Lf);      // the other digit plane of this template
                        float sc = 0.f;
                        if (tli <= A.T - L) {
                            // q = lo + 256 hi in fp32 (the I2F.F64 -> DFMA -> DMUL -> F2F chain per output stalled these warps): |hi| < 2^24 converts
                            // exactly, so the result carries three fp32 roundings (~2e-7 relative), far below the 1.5e-5
                            // quantisation step of the 16-bit weights
                            const float lo = (float)(plane ? other : s), hi = (float)(plane ? s : other);
                            sc = __fmaf_rn(hi, 256.f, lo) * sc_m;
                        }
                        if (A.scores && live && plane == 0 && tr == 0) A.scores[g * f.M + m] = sc;
                        if (A.keys) {
                            // max over the warp's templates: one REDUX on the monotone bit pattern; the lowest lane that holds
                            // the maximum belongs to the lowest template index (first maximum wins, like detect_key)
                            uint32_t uk = __float_as_uint(sc + cs_m);
                            uk ^= (uk >> 31) ? 0xFFFFFFFFu : 0x80000000u;
                            if (!live) uk = 0u;
                            const uint32_t umax = __reduce_max_sync(0xffffffffu, uk);
                            const uint32_t who = __ballot_sync(0xffffffffu, live && uk == umax);
                            if (lane == 0 && who) {
                                const int bl = __ffs(who) - 1;
                                const int mb = (nt * f.tptv + (quad * 32 + bl) / Lf) >> 1;
                                atomicMax(A.keys + g, ((unsigned long long)umax << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)mb));
                            }
                        }
                    }
                }
                // hand the accumulator back zeroed
                for (int c = 0; c < step_rows; c += 8) tmem_st8_zero(taddr + c);
                tmem_st_wait_all();
                tc_fence_before();
                mbar_arrive(&tempty_bar[acc]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 13) {
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

// Real-valued features: separate multiply and add in (tau, d) order (detect_float :206-216) -- the reference's exact
// float32 arithmetic, so the result is bit-identical to it.  That fixes the order of the L*D additions of every
// output: an output is one dependent chain, and what a kernel can still do is keep several chains in flight per thread
// and feed them from shared memory.  CTA = 128 consecutive frame rows x DF_TM templates; per (tau, 32-feature chunk)
// the 128 x 32 window of features is staged coalesced (pitch 33: row-per-thread reads are conflict free) next to the
// DF_TM weight rows (broadcast reads); a thread owns one row and DF_TM independent accumulators: 1 + DF_TM shared loads
// for 2 * DF_TM arithmetic instructions per feature, no global load in the inner loop.
// TM = templates per thread (independent chains; their weights are broadcast LDS.128), DC = features per staged step.
// TM = 8 when the grid still fills the SMs several times over, else 4 (more CTAs); DC = 64 for launches of at most one CTA
// per SM (fewer steps: a step costs two barriers and exposes part of a global-load latency, which is what a single-utterance
// call consists of), else 32 (128 instead of 243 registers).
template <int TM, int DC>
__global__ void __launch_bounds__(128) detect_float_kernel(const float* __restrict__ Fm, const float* __restrict__ LF,
                                                           const float* __restrict__ init, float* __restrict__ scores, long long R,
                                                           int T, int D, int M, int L) {
    __shared__ float sx[128 * (DC + 1)];
    __shared__ __align__(16) float sw[DC * TM];                // [feature c][template j]
    const int tid = threadIdx.x;
    const long long g0 = (long long)blockIdx.x * 128;
    const int m0 = blockIdx.y * TM;
    const long long g = g0 + tid;
    const bool valid = g < R && (int)(g % T) <= T - L;
    float acc[TM];
#pragma unroll
    for (int j = 0; j < TM; ++j) acc[j] = (init && m0 + j < M) ? init[m0 + j] : 0.f;     // detect_float_max starts at cs[m] (:258)
    // (tau, chunk) steps in the reference's order; the next step's 128 x DC window and weight rows are fetched into
    // registers while the current one is being consumed from shared memory.  Per feature a thread issues 1 + TM/4 shared
    // loads (its x, float4s of weights) for 2 TM arithmetic instructions.
    const int nd = (D + DC - 1) / DC;
    const int nstep = L * nd;
    constexpr int NW = TM * DC / 128;                           // weights per thread and step
    float px[DC], pw[NW];
    auto fetch = [&](int step) {
        const int tau = step / nd, d0 = (step - tau * nd) * DC;
        const int dc = D - d0 < DC ? D - d0 : DC;
#pragma unroll
        for (int i = 0; i < DC; ++i) {                      // element e = tid + 128 i of the window: row e / DC, column e % DC
            const int e = tid + 128 * i, r = e / DC, c = e % DC;
            const long long row = g0 + r + tau;
            px[i] = (row < R && c < dc) ? __ldg(Fm + row * D + d0 + c) : 0.f;
        }
#pragma unroll
        for (int h = 0; h < NW; ++h) {                      // weight e = tid + 128 h: template e / DC, feature e % DC
            const int e = tid + 128 * h, j = e / DC, c = e % DC;
            pw[h] = (m0 + j < M && c < dc) ? __ldg(LF + ((size_t)(m0 + j) * L + tau) * D + d0 + c) : 0.f;
        }
    };
    fetch(0);
    for (int step = 0; step < nstep; ++step) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < DC; ++i) {
            const int e = tid + 128 * i;
            sx[(e / DC) * (DC + 1) + (e % DC)] = px[i];
        }
#pragma unroll
        for (int h = 0; h < NW; ++h) {
            const int e = tid + 128 * h;
            sw[(e % DC) * TM + (e / DC)] = pw[h];
        }
        __syncthreads();
        if (step + 1 < nstep) fetch(step + 1);
        const float* xr = sx + tid * (DC + 1);
#pragma unroll
        for (int c = 0; c < DC; ++c) {                     // columns past the end of the row hold zeros: x * w = +-0, acc unchanged
            const float x = xr[c];
#pragma unroll
            for (int q = 0; q < TM / 4; ++q) {
                const float4 w = *reinterpret_cast<const float4*>(sw + c * TM + 4 * q);
                acc[4 * q + 0] = __fadd_rn(acc[4 * q + 0], __fmul_rn(w.x, x)); acc[4 * q + 1] = __fadd_rn(acc[4 * q + 1], __fmul_rn(w.y, x));
                acc[4 * q + 2] = __fadd_rn(acc[4 * q + 2], __fmul_rn(w.z, x)); acc[4 * q + 3] = __fadd_rn(acc[4 * q + 3], __fmul_rn(w.w, x));
            }
        }
    }
    if (g < R) {
#pragma unroll
        for (int j = 0; j < TM; ++j)
            if (m0 + j < M) scores[g * M + m0 + j] = valid ? acc[j] : 0.f;
    }
}

// best/arg over templates, first maximum wins (strict >, detect_float_max :268-277).
// Eight lanes share a row: lane q takes templates 4q..4q+3 (+32, +64, ...) as one 16-byte load, so a
// warp reads 4 rows x 128 B contiguously; (value, index) pairs are then combined with three shuffles,
// the lower index winning ties -- the same result as the sequential strict-> scan.
__global__ void __launch_bounds__(256) detect_reduce_kernel(const float* __restrict__ scores, const float* __restrict__ cs,
                                                            long long R, int T, int M, int L, int tail_mode,
                                                            float* __restrict__ best, uint16_t* __restrict__ arg) {
    const int q = threadIdx.x & 7;
    const bool vec = (M & 3) == 0 && (reinterpret_cast<uintptr_t>(scores) & 15) == 0;
    const long long stride = (long long)gridDim.x * (blockDim.x >> 3);
    // two rows per 8-lane group and iteration (two independent 16-byte loads in flight per lane); same trip count for
    // every lane (shuffles below)
    const long long nloop = (R + 2 * stride - 1) / (2 * stride);
    long long g0 = blockIdx.x * (long long)(blockDim.x >> 3) + (threadIdx.x >> 3);
    for (long long it = 0; it < nloop; ++it, g0 += 2 * stride) {
        float bv[2];
        int bm[2];
        bool live[2], valid[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long g = g0 + u * stride;
            live[u] = g < R;
            const int tl = live[u] ? (R < (1ll << 31) ? (int)((unsigned)g % (unsigned)T) : (int)(g % T)) : 0;
            valid[u] = live[u] && tl <= T - L;
            bv[u] = -INFINITY;
            bm[u] = 0x7fffffff;
        }
        if (vec && M <= 32) {          // one float4 per lane covers the row: both rows' loads issue before either is used
            float4 x[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid[u] && 4 * q < M) x[u] = __ldcs(reinterpret_cast<const float4*>(scores + (g0 + u * stride) * M + 4 * q));
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if ((valid[u] || (live[u] && tail_mode == 1)) && 4 * q < M) {
                    const float v[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float y = v[i] + (cs ? __ldg(cs + 4 * q + i) : 0.f);
                        if (y > bv[u]) { bv[u] = y; bm[u] = 4 * q + i; }
                    }
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const long long g = g0 + u * stride;
                if (valid[u] || (live[u] && tail_mode == 1)) {
                    for (int m0 = 4 * q; m0 < M; m0 += 32) {
                        float v[4] = {0.f, 0.f, 0.f, 0.f};
                        if (valid[u]) {
                            if (vec) {
                                const float4 x = __ldcs(reinterpret_cast<const float4*>(scores + g * M + m0));
                                v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
                            } else {
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    if (m0 + i < M) v[i] = scores[g * M + m0 + i];
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (m0 + i < M) {
                                const float y = v[i] + (cs ? __ldg(cs + m0 + i) : 0.f);
                                if (y > bv[u]) { bv[u] = y; bm[u] = m0 + i; }
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int off = 4; off >= 1; off >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv[u], off);
                const int om = __shfl_xor_sync(0xffffffffu, bm[u], off);
                if (ov > bv[u] || (ov == bv[u] && om < bm[u])) { bv[u] = ov; bm[u] = om; }
            }
            if (live[u] && q == 0) {
                const long long g = g0 + u * stride;
                if (!(valid[u] || tail_mode == 1)) bv[u] = 0.f;
                if (bm[u] == 0x7fffffff) bm[u] = 0;
                if (best) best[g] = bv[u];
                if (arg) arg[g] = (uint16_t)bm[u];
            }
        }
    }
}

// keys -> (best, arg)
__global__ void detect_decode_kernel(const unsigned long long* __restrict__ keys, long long R, float* __restrict__ best,
                                     uint16_t* __restrict__ arg) {
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < R; g += (long long)gridDim.x * blockDim.x) {
        const unsigned long long k = keys[g];
        uint32_t u = (uint32_t)(k >> 32);
        u ^= (u >> 31) ? 0x80000000u : 0xFFFFFFFFu;
        if (best) best[g] = __uint_as_float(u);
        if (arg) arg[g] = (uint16_t)(0xFFFFFFFFu - (uint32_t)k);
    }
}

}  // namespace tsr

using namespace tsr;

static long long* g_detect_trace = nullptr;   // debug hook
static int g_detect_dbg = 0;

extern "C" {

/* Debug only (not in the public header): device buffer of DT_TRACE_ITERS*8 int64 clock stamps. */
void tsr_debug_set_detect_trace(long long* dev_buf) { g_detect_trace = dev_buf; }
void tsr_debug_set_detect_flags(int flags) { g_detect_dbg = flags; }

size_t tsr_filter_blob_bytes(int M, int L, int D, int plane_format) {
    if (M < 1 || L < 1 || L > 256 || D < 1 || plane_format < 0 || plane_format > 4) return 0;
    return filter_layout(M, L, D, plane_format).total;
}

int tsr_prepare_filters(const float* LF, int M, int L, int D, int plane_format, void* blob, void* stream) {
    TSR_REQUIRE(LF && blob, "tsr_prepare_filters: null pointer");
    TSR_REQUIRE(M >= 1 && M <= 65535, "tsr_prepare_filters: need 1 <= M <= 65535 templates (uint16 argmax), got %d", M);
    TSR_REQUIRE(L >= 1 && L <= 256, "tsr_prepare_filters: template length %d outside [1,256]", L);
    TSR_REQUIRE(D >= 1, "tsr_prepare_filters: D < 1");
    TSR_REQUIRE(plane_format >= 0 && plane_format <= 4, "tsr_prepare_filters: plane_format must be one of TSR_PLANES_*");
    int rc = require_sm100("tsr_prepare_filters");
    if (rc) return rc;
    FilterLayout f = filter_layout(M, L, D, plane_format);
    cudaStream_t st = (cudaStream_t)stream;
    TSR_CUDA(cudaMemsetAsync(blob, 0, f.hdr_bytes, st));
    dim3 g(16, M);
    filters_maxabs_kernel<<<g, 256, 0, st>>>(LF, M, (size_t)L * D, (uint8_t*)blob);
    prepare_filters_kernel<<<sm_count() * 8, 256, 0, st>>>(LF, (uint8_t*)blob, f, plane_format);
    filters_publish_kernel<<<(M + 255) / 256, 256, 0, st>>>((uint8_t*)blob, plane_format, f);
    return check_launch("tsr_prepare_filters");
}

// TN kernel: partition the rows into units and the units into evenly filled steps of step_rows <= 256 frames
static int detect_tn_launch(DetectArgs& A, cudaStream_t st) {
    const FilterLayout& f = A.f;
    const TnSmem S = tn_smem_layout(f);
    if (S.total > 227 * 1024) {
        set_error("tsr_detect: L=%d needs %u B of shared memory (> 227 KB)", f.L, S.total);
        return TSR_ERR_UNSUPPORTED;
    }
    const int nsm = sm_count();
    const long long R = A.R;
    const int min_rows = 128;                      // narrower MMAs are issue bound; also keeps step_rows > Lc + 8
    long long best_cost = -1;
    int best_U = 1, best_steps = 1, best_sr = DT_TN_STEP_MAX;
    const long long maxU = (R + min_rows - 1) / min_rows;
    for (long long U = 1; U <= maxU && U <= 8192; ++U) {
        const long long need = (R + U - 1) / U + f.Lc;
        const long long steps = (need + DT_TN_STEP_MAX - 1) / DT_TN_STEP_MAX;
        long long sr = (((need + steps - 1) / steps) + 15) & ~15LL;
        if (sr < min_rows) sr = min_rows;
        const long long waves = ((long long)f.ntiles * U + nsm - 1) / nsm;
        const long long cost = waves * steps * (sr + 8);        // + a per-step constant (accumulator hand-over)
        // ties go to more units as long as they fit one wave: every SM gets a unit and each unit fewer padded rows
        if (best_cost < 0 || cost < best_cost || (cost == best_cost && (long long)f.ntiles * U <= nsm)) {
            best_cost = cost; best_U = (int)U; best_steps = (int)steps; best_sr = (int)sr;
        }
    }
    A.units_per_ntile = best_U;
    A.rows_per_unit = (int)((R + best_U - 1) / best_U);
    A.steps_per_unit = best_steps;
    A.step_rows = best_sr;
    const size_t set_bytes = (size_t)f.nchunks * f.tile_bytes;
    long long G = (long long)(64u << 20) / (long long)(set_bytes ? set_bytes : 1);
    if (G < 1) G = 1;
    if (G > f.ntiles) G = f.ntiles;
    A.ntile_group = (int)G;
    A.nissue = 2;
#ifdef TSR_DEBUG_TRACE
    if (const char* ov = getenv("TSR_DEBUG_DETECT_FLAGS")) A.dbg = atoi(ov);   // timing experiments, wrong results
#endif
#ifdef TSR_DEBUG_TRACE
    if (const char* ov = getenv("TSR_DEBUG_DETECT_NISSUE")) A.nissue = ov[0] == '1' ? 1 : 2;   // experiments only
#endif
    const long long nunits = (long long)f.ntiles * best_U;
    const int blocks = (int)(nunits < nsm ? nunits : nsm);
    if (A.nissue == 2) {
        TSR_CUDA(cudaFuncSetAttribute(detect_tn_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.total));
        detect_tn_kernel<2><<<blocks, DT_THREADS, S.total, st>>>(A);
    } else {
        TSR_CUDA(cudaFuncSetAttribute(detect_tn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.total));
        detect_tn_kernel<1><<<blocks, DT_THREADS - 32, S.total, st>>>(A);
    }
    return check_launch("tsr_detect");
}

static int detect_impl(const uint32_t* featbits, int B, int T, int D, int rowwords, const void* filter_blob,
                       const float* LF_or_null, int M, int L, int plane_format, float* scores, unsigned long long* keys,
                       const float* cs, int impl, void* stream) {
    if (B == 0) return TSR_OK;
    TSR_REQUIRE(featbits && (scores || keys), "tsr_detect: null pointer");
    TSR_REQUIRE(B >= 0 && T >= 1 && D >= 1 && M >= 1 && L >= 1, "tsr_detect: bad sizes");
    TSR_REQUIRE(rowwords == tsr_rowwords(D), "tsr_detect: rowwords must be tsr_rowwords(D)");
    TSR_REQUIRE(impl == TSR_IMPL_TENSOR || impl == TSR_IMPL_SIMT, "tsr_detect: unknown impl %d", impl);
    int rc = require_sm100("tsr_detect");
    if (rc) return rc;
    const long long R = (long long)B * T;
    if (L > T) {   // no valid offset: all-zero output, like the reference's zero-initialised array
        TSR_REQUIRE(scores, "tsr_detect_best: template longer than the utterances");
        TSR_CUDA(cudaMemsetAsync(scores, 0, sizeof(float) * R * M, (cudaStream_t)stream));
        return TSR_OK;
    }
    if (impl == TSR_IMPL_SIMT) {
        TSR_REQUIRE(scores, "tsr_detect(simt): needs the score matrix");
        TSR_REQUIRE(LF_or_null, "tsr_detect(simt): needs the fp32 filters");
        long long blocks = (R * M + 127) / 128;
        if (blocks > (long long)sm_count() * 32) blocks = (long long)sm_count() * 32;
        detect_simt_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(featbits, LF_or_null, scores, R, T, D, rowwords, M, L);
        return check_launch("tsr_detect(simt)");
    }
    TSR_REQUIRE(filter_blob, "tsr_detect: needs the prepared filter blob");
    TSR_REQUIRE(L <= 256, "tsr_detect: template length %d > 256 unsupported on the tensor path", L);
    TSR_REQUIRE(plane_format >= 0 && plane_format <= 4, "tsr_detect: plane_format must be one of TSR_PLANES_*");
    DetectArgs A;
    A.bits = featbits; A.blob = (const uint8_t*)filter_blob; A.scores = scores; A.R = R; A.T = T;
    A.keys = keys; A.cs = cs;
    A.f = filter_layout(M, L, D, plane_format);
    A.trace = g_detect_trace;
    A.dbg = g_detect_dbg;
    if (A.f.kind == 2) return detect_tn_launch(A, (cudaStream_t)stream);
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
        long long steps = (rows + A.f.Lc + step_rows - 1) / step_rows;
        long long waves = ((long long)A.f.ntiles * U + nsm - 1) / nsm;
        long long cost = waves * steps;
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_U = (int)U; }
    }
    A.units_per_ntile = best_U;
    {   // N tiles per group: their filter tiles (one full set each) should fit in about half of the 126 MB L2
        const size_t set_bytes = (size_t)A.f.nchunks * A.f.tile_bytes;
        long long G = (long long)(64u << 20) / (long long)(set_bytes ? set_bytes : 1);
        if (G < 1) G = 1;
        if (G > A.f.ntiles) G = A.f.ntiles;
        A.ntile_group = (int)G;
        // Two issuer warps lift config 3 (32 templates, bank L2 resident) from 141 to 128 cycles per MMA, 2.62 -> 2.49 ms,
        // but slow the 1000-template stream (bank 438 MB, streamed through L2 group by group) from 264 to 290 ms.
        A.nissue = (size_t)A.f.ntiles * set_bytes <= ((size_t)64 << 20) ? 2 : 1;
#ifdef TSR_DEBUG_TRACE
        if (const char* ov = getenv("TSR_DEBUG_DETECT_NISSUE")) A.nissue = ov[0] == '2' ? 2 : 1;   // experiments only
#endif
    }
    A.rows_per_unit = (int)((R + best_U - 1) / best_U);
    A.steps_per_unit = (A.rows_per_unit + A.f.Lc + step_rows - 1) / step_rows;
    const long long nunits = (long long)A.f.ntiles * best_U;
    int blocks = (int)(nunits < nsm ? nunits : nsm);
    if (A.f.kind) {
        auto launch = [&](auto kern, int threads) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.total);
            if (e == cudaSuccess) kern<<<blocks, threads, S.total, (cudaStream_t)stream>>>(A);
            return e;
        };
        // the second issuer is warp 14: launched only when used
        if (A.nissue == 2) TSR_CUDA(L > 1 ? launch(detect_tc_kernel<true, 2, true>, DT_THREADS) : launch(detect_tc_kernel<true, 2, false>, DT_THREADS));
        else TSR_CUDA(L > 1 ? launch(detect_tc_kernel<true, 1, true>, DT_THREADS - 32) : launch(detect_tc_kernel<true, 1, false>, DT_THREADS - 32));
    } else {
        TSR_CUDA(cudaFuncSetAttribute(detect_tc_kernel<false, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.total));
        detect_tc_kernel<false, 1, true><<<blocks, DT_THREADS - 32, S.total, (cudaStream_t)stream>>>(A);
    }
    return check_launch("tsr_detect");
}

int tsr_detect(const uint32_t* featbits, int B, int T, int D, int rowwords, const void* filter_blob,
               const float* LF_or_null, int M, int L, int plane_format, float* scores, int impl, void* stream) {
    if (B == 0) return TSR_OK;
    TSR_REQUIRE(scores, "tsr_detect: null pointer");
    return detect_impl(featbits, B, T, D, rowwords, filter_blob, LF_or_null, M, L, plane_format, scores, nullptr, nullptr, impl, stream);
}

int tsr_detect_best(const uint32_t* featbits, int B, int T, int D, int rowwords, const void* filter_blob, int M, int L,
                    int plane_format, const float* cs_or_null, float* scores_or_null, void* keys_workspace, float* best,
                    uint16_t* arg, void* stream) {
    if (B == 0) return TSR_OK;
    TSR_REQUIRE(featbits && filter_blob && keys_workspace && (best || arg), "tsr_detect_best: null pointer");
    TSR_REQUIRE(B >= 0 && T >= 1 && L >= 1 && L <= T && M >= 1 && M <= 65535, "tsr_detect_best: bad sizes (need L <= T, M <= 65535)");
    int rc = require_sm100("tsr_detect_best");
    if (rc) return rc;
    const long long R = (long long)B * T;
    cudaStream_t st = (cudaStream_t)stream;
    TSR_CUDA(cudaMemsetAsync(keys_workspace, 0, (size_t)R * 8, st));
    rc = detect_impl(featbits, B, T, D, rowwords, filter_blob, nullptr, M, L, plane_format, scores_or_null,
                     (unsigned long long*)keys_workspace, cs_or_null, TSR_IMPL_TENSOR, stream);
    if (rc) return rc;
    long long blocks = (R + 255) / 256;
    if (blocks > (long long)sm_count() * 8) blocks = (long long)sm_count() * 8;
    detect_decode_kernel<<<(unsigned)blocks, 256, 0, st>>>((const unsigned long long*)keys_workspace, R, best, arg);
    return check_launch("tsr_detect_best");
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
    long long blocks = (R + 63) / 64;          // 2 x 32 rows per 256-thread CTA and iteration
    if (blocks > (long long)sm_count() * 32) blocks = (long long)sm_count() * 32;
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
    const long long rb = (R + 127) / 128;
    const int tm = (rb * ((M + 7) / 8) >= 4LL * sm_count()) ? 8 : 4;
    TSR_REQUIRE(rb < (1ll << 31) && (M + tm - 1) / tm <= 65535, "tsr_detect_float: problem too large");
    dim3 grid((unsigned)rb, (unsigned)((M + tm - 1) / tm));
    cudaStream_t st = (cudaStream_t)stream;
    // measured (64 x 300 x 320 features, 8 templates): <4,32> 0.79 ms, <8,32> 0.89 ms (150 CTAs: too few), <4,64> 1.11 ms (243
    // registers); one utterance x 160 bands x 1 template: <4,32> 383 us, <4,64> 221 us
    if (tm == 8) detect_float_kernel<8, 32><<<grid, 128, 0, st>>>(Fm, LF, init_or_null, scores, R, T, D, M, L);
    else if (D > 32 && rb * grid.y <= (long long)sm_count()) detect_float_kernel<4, 64><<<grid, 128, 0, st>>>(Fm, LF, init_or_null, scores, R, T, D, M, L);
    else detect_float_kernel<4, 32><<<grid, 128, 0, st>>>(Fm, LF, init_or_null, scores, R, T, D, M, L);
    return check_launch("tsr_detect_float");
}

}  // extern "C"
