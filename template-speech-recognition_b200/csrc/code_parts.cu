// code_parts.cu -- K2: Bernoulli part coding of every ph x pw x 8 patch of the packed edge map,
// fused with the edge-count gate and the arg-max (code_parts.pyx:70-135, get_train_data.py:2621).
//
//   score[loc][k] = c[k] + sum_{a,b,z} X[i+a][j+b][z] * W[k][a][b][z],   W = lp - lip, c = sum lip
//
// Tensor path (TSR_IMPL_TENSOR): one persistent CTA per SM, warp specialised.
//   producers (4 warps)  expand packed edge bytes of a 128-location tile (+halo) into bf16 {0,1}
//                        rows in shared memory: one 16-byte row per (t,f) position, positions
//                        flattened p = t*F + f, so location p's patch row `a` is the 16*pw bytes
//                        starting at position p + a*F.
//   MMA issuers (4 warps, one per SMSP; tile i -> issuer i % 4 -> accumulator i % 4; a single thread gets a
//                        tcgen05.mma out only every ~75-90 cycles, a 56-cycle MMA needs more than one issuer)
//                        tcgen05.mma M=128 x N=N_pad x K=16, A straight out of that buffer with
//                        a K-major no-swizzle descriptor whose core-matrix rows are the 16-byte
//                        positions (row stride 16 B = one frequency bin, SBO = 128 B, LBO = the
//                        distance to the next (a,b) group): the im2col is never materialised.
//                        hi and lo 16-bit weight planes accumulate into the same TMEM columns, after
//                        one extra K-step (constant-ones A) that sets the accumulator to c[k].
//   epilogue (2 x 4 warps, even / odd tiles) tcgen05.ld the 128 x N_pad fp32 tile, running arg-max (lowest k
//                        wins ties), read the exact edge counts from two extra weight columns
//                        (all-ones / mask, columns N_pad-2 / N_pad-1), apply the gate, store uint8 codes;
//                        no shared-memory reads (they stall behind the MMA operand traffic).
// Four TMEM accumulator buffers (two per epilogue set) and a 4-deep A ring overlap the three roles: an
// accumulator is held by the epilogue for about as long as the MMAs need to fill it, so with only two the MMA
// warps idled half the time waiting for a free one.
//
// Int8 form (TSR_IMPL_TENSOR_INT8, template parameter I8): same roles and the same in-place im2col, but one fixed-point
// scale for all parts, both base-255 weight digits of a position inside one 16-byte K group (A row = [x | 255 x]),
// kind::i8 MMAs of K = 32 (14 per tile instead of 27), an integer arg-max that also tracks the runner-up, and
// code_parts_refine_kernel, which re-evaluates every location whose runner-up lies inside the rounding bound with the
// reference's own float64 sum: the codes equal the reference's arg-max everywhere.  See prepare_parts_i8_kernel.
#include <atomic>

#include "common.cuh"

namespace tsr {

#define CP_MAX_CSTEPS 4      // int8 form: K-steps that load the per-part constant (each carries up to +-1 003 935 units)
#define CP_I8_QMAX 32512     // 127 + 255 * 127: two balanced base-255 digits
#define CP_I8_CSTEP_CAP 1003935   // 255 * 31 * 127 (+ a last digit of weight 1)

struct PartsHeaderI8 {       // written by prepare_parts_i8_kernel
    double scale;            // weight units: w ~ q * scale
    int cmid;                // c[k] ~ (cq[k] + cmid) * scale
    int ncsteps, qmax, E;
};
struct PartsLayout {
    int K, ph, pw, npad, nsteps, nq;
    size_t off_w, off_cstep, off_c, off_wf, off_mask, off_hdr, total;
    // int8 form (TSR_IMPL_TENSOR_INT8): B operand, constant steps, the integer table / float64 log tables of the refinement
    size_t off_q, off_cstep8, off_qtab, off_cq, off_lp64, off_lip64, off_hdr8;
};

__host__ __device__ inline PartsLayout parts_layout(int K, int ph, int pw) {
    PartsLayout L;
    L.K = K; L.ph = ph; L.pw = pw;
    L.nq = ph * pw;
    L.npad = ((K + 2) + 15) & ~15;
    L.nsteps = (L.nq + 1) / 2;
    L.off_w = 0;
    L.off_cstep = (size_t)2 * L.nsteps * L.npad * 32;   // one more K-step of weights: c[k] in three 16-bit terms
    L.off_c = L.off_cstep + (size_t)L.npad * 32;
    L.off_wf = L.off_c + (size_t)L.npad * 4;
    L.off_mask = L.off_wf + (size_t)L.nq * 8 * L.npad * 4;
    L.off_hdr = L.off_mask + 8;          // [0] plane format, [1] max |w| bits
    L.off_q = (L.off_mask + 16 + 127) & ~(size_t)127;                 // [step][kgroup][row][16 B]: 8 low digits | 8 high digits
    L.off_cstep8 = L.off_q + (size_t)L.nsteps * L.npad * 32;          // [cstep][kgroup][row][16 B], contiguous with the steps
    L.off_qtab = L.off_cstep8 + (size_t)CP_MAX_CSTEPS * L.npad * 32;  // int16 [nq*8][npad]
    L.off_cq = L.off_qtab + (size_t)L.nq * 8 * L.npad * 2;            // int32 [npad]
    L.off_lp64 = (L.off_cq + (size_t)L.npad * 4 + 15) & ~(size_t)15;  // double [K][nq*E]
    L.off_lip64 = L.off_lp64 + (size_t)K * L.nq * 8 * 8;
    L.off_hdr8 = L.off_lip64 + (size_t)K * L.nq * 8 * 8;
    L.total = L.off_hdr8 + 32;
    return L;
}

// ------------------------------------------------------------------------------ prepare
__global__ void parts_maxabs_kernel(const double* __restrict__ lp, const double* __restrict__ lip, size_t n,
                                    uint8_t* __restrict__ blob, PartsLayout L) {
    uint32_t* hdr = reinterpret_cast<uint32_t*>(blob + L.off_hdr);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        atomic_max_abs(&hdr[1], (float)(lp[i] - lip[i]));
    // c[k] = sum lip rides through the tensor cores as well (see prepare_parts_kernel): fp16 planes only if it fits
    const size_t per = n / L.K;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < (size_t)L.K; k += (size_t)gridDim.x * blockDim.x) {
        double c = 0.0;
        for (size_t i = 0; i < per; ++i) c += lip[k * per + i];
        atomic_max_abs(&hdr[1], (float)(c * (TSR_FP16_MAX_WEIGHT / 60000.0)));
    }
}

// One thread per (n, q, z) weight: n < K parts, n == N_pad-2 all-ones count column, n == N_pad-1 mask column.
__global__ void prepare_parts_kernel(const double* __restrict__ lp, const double* __restrict__ lip, int E,
                                     const uint8_t* __restrict__ mask, uint8_t* __restrict__ blob, PartsLayout L,
                                     int plane_format) {
    uint32_t* hdr = reinterpret_cast<uint32_t*>(blob + L.off_hdr);
    const uint32_t fmt = pick_format(hdr[1], plane_format);
    const int total = L.npad * L.nsteps * 16;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int n = i / (L.nsteps * 16);
        int rem = i - n * (L.nsteps * 16);
        int s = rem >> 4, e = rem & 15;
        int q = 2 * s + (e >> 3), z = e & 7;
        double w = 0.0;
        if (q < L.nq && z < E) {
            if (n < L.K) {
                size_t idx = ((size_t)n * L.nq + q) * E + z;
                w = lp[idx] - lip[idx];
            } else if (n == L.npad - 2) {
                w = 1.0;
            } else if (n == L.npad - 1) {
                w = (mask && mask[q]) ? 1.0 : 0.0;
            }
        }
        uint16_t hi, lo;
        split_hi_lo(w, fmt, hi, lo);
        // [plane][step][kgroup][row][8]: both planes accumulate into the same TMEM columns
        size_t o = (((size_t)s * 2 + (e >> 3)) * L.npad + n) * 16 + z * 2;
        *reinterpret_cast<uint16_t*>(blob + L.off_w + o) = hi;
        *reinterpret_cast<uint16_t*>(blob + L.off_w + (size_t)L.nsteps * L.npad * 32 + o) = lo;
        if (q < L.nq) reinterpret_cast<float*>(blob + L.off_wf)[((size_t)q * 8 + z) * L.npad + n] = (float)w;
    }
    // c[k] = sum lip (float64, reference order), one thread per part
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < L.npad; n += gridDim.x * blockDim.x) {
        double c = 0.0;
        if (n < L.K)
            for (int i = 0; i < L.nq * E; ++i) c += lip[(size_t)n * L.nq * E + i];
        // padded columns (count / mask / zero columns) get -inf: they can never win the arg-max
        reinterpret_cast<float*>(blob + L.off_c)[n] = n < L.K ? (float)c : -INFINITY;
        // tensor path: c[k] = t0 + t1 + t2 (16-bit terms, >= 24 significant bits) sits in elements 0..2 of an extra
        // K-step whose A operand is a constant [1,1,1,0,...] row, so the accumulator starts at c[k] and the epilogue
        // needs no shared-memory reads (they stall behind the MMA operand traffic).  Padded columns start at 0.
        uint16_t t[3] = {0, 0, 0};
        if (n < L.K) {
            uint16_t lo_unused;
            double r = c;
            for (int j = 0; j < 3; ++j) {
                split_hi_lo(r, fmt, t[j], lo_unused);
                r -= fmt == TSR_FMT_FP16 ? (double)__half2float(__ushort_as_half(t[j])) : (double)__bfloat162float(__ushort_as_bfloat16(t[j]));
            }
        }
        uint16_t* row = reinterpret_cast<uint16_t*>(blob + L.off_cstep + (size_t)n * 16);          // k-group 0
        uint16_t* row1 = reinterpret_cast<uint16_t*>(blob + L.off_cstep + ((size_t)L.npad + n) * 16);  // k-group 1
        for (int j = 0; j < 8; ++j) { row[j] = j < 3 ? t[j] : (uint16_t)0; row1[j] = 0; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long m = 0;
        if (mask)
            for (int q = 0; q < L.nq; ++q)
                if (mask[q]) m |= 1ull << q;
        reinterpret_cast<unsigned long long*>(blob + L.off_mask)[0] = m;
    }
}
// runs after prepare_parts_kernel (same stream): publish the chosen format
__global__ void parts_publish_format_kernel(uint8_t* __restrict__ blob, PartsLayout L, int plane_format) {
    uint32_t* hdr = reinterpret_cast<uint32_t*>(blob + L.off_hdr);
    hdr[0] = pick_format(hdr[1], plane_format);
}


// Int8 form of the same model (TSR_IMPL_TENSOR_INT8), one CTA.  All parts share one fixed-point scale so that the
// integer accumulators of different parts compare directly:
//   q[k][a][b][z] = rint(W / scale) in [-qmax, qmax], qmax <= 32512, as two balanced base-255 digits (d0 + 255 d1);
//   the kernel's A row of a position is [x_z | 255 x_z] (16 bytes), B holds [d0_z | d1_z]: both digits accumulate into
//   the SAME int32 column, one MMA of K = 32 covers two positions.
//   c[k] -> cq[k] = rint(c / scale) - cmid (cmid = the mid-range: it cannot change the arg-max), loaded by ncsteps
//   constant K-steps: A = [255 x 31, 1], B digits e with cq = sum_j (255 * sum_i e_ji + r_j).
// |score_k / scale - cmid - acc_k| <= (set bits + 1) / 2, so two parts whose accumulators differ by more than
// set bits + 1 are ordered like their exact scores; the kernel flags everything else for the float64 refinement.
// qmax also keeps |acc| * 256 + 255 inside int32 for ANY input (the epilogue packs (acc, 255 - k) into one key).
__global__ void __launch_bounds__(256) prepare_parts_i8_kernel(const double* __restrict__ lp, const double* __restrict__ lip, int E,
                                                               const uint8_t* __restrict__ mask, uint8_t* __restrict__ blob, PartsLayout L) {
    __shared__ double s_c[256];
    __shared__ double s_red[3][256];
    __shared__ PartsHeaderI8 s_h;
    const int tid = threadIdx.x, per = L.nq * E;
    double mw = 0.0;
    for (int i = tid; i < L.K * per; i += 256) {
        const double w = fabs(lp[i] - lip[i]);
        if (w > mw && w < 1e300) mw = w;
    }
    double c = 0.0;
    if (tid < L.K)
        for (int i = 0; i < per; ++i) c += lip[(size_t)tid * per + i];
    s_c[tid] = c;
    s_red[0][tid] = mw;
    s_red[1][tid] = tid < L.K ? c : 1e300;
    s_red[2][tid] = tid < L.K ? c : -1e300;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) {
            s_red[0][tid] = fmax(s_red[0][tid], s_red[0][tid + o]);
            s_red[1][tid] = fmin(s_red[1][tid], s_red[1][tid + o]);
            s_red[2][tid] = fmax(s_red[2][tid], s_red[2][tid + o]);
        }
        __syncthreads();
    }
    if (tid == 0) {
        double maxw = s_red[0][0], cmin = s_red[1][0], cmax = s_red[2][0];
        if (!(maxw > 0.0)) maxw = 1.0;
        if (!(cmax - cmin < 1e300)) { cmin = 0.0; cmax = 0.0; }          // non-finite constants: nothing sensible to keep
        const double R = (cmax - cmin) * 0.5 / maxw + 1.0;                // half range of c in units of max |w| (+ rounding slack)
        double qm = CP_I8_QMAX;
        qm = fmin(qm, floor((8388608.0 - 256.0) / ((double)per + R)));
        qm = fmin(qm, floor((double)CP_MAX_CSTEPS * CP_I8_CSTEP_CAP / R));
        if (qm < 1.0) qm = 1.0;
        s_h.qmax = (int)qm;
        s_h.scale = maxw / qm;
        s_h.cmid = (int)llrint((cmax + cmin) * 0.5 / s_h.scale);
        s_h.E = E;
        s_h.ncsteps = 1;
    }
    __syncthreads();
    const double scale = s_h.scale;
    const int cmid = s_h.cmid, qmax = s_h.qmax;
    // constants: cq, their digits in CP_MAX_CSTEPS steps, the number of steps in use
    int mycq = 0;
    if (tid < L.npad) {
        if (tid < L.K) {
            long long v = llrint(s_c[tid] / scale) - (long long)cmid;
            const long long cap = (long long)CP_MAX_CSTEPS * CP_I8_CSTEP_CAP;
            mycq = (int)(v > cap ? cap : (v < -cap ? -cap : v));
        }
        reinterpret_cast<int*>(blob + L.off_cq)[tid] = mycq;
        int rem = mycq;
        for (int j = 0; j < CP_MAX_CSTEPS; ++j) {
            int piece = rem > CP_I8_CSTEP_CAP ? CP_I8_CSTEP_CAP : (rem < -CP_I8_CSTEP_CAP ? -CP_I8_CSTEP_CAP : rem);
            rem -= piece;
            int S = (piece >= 0 ? piece + 127 : piece - 127) / 255;       // round to nearest
            const int r = piece - 255 * S;                                // in [-127, 127]
            int8_t* g0 = reinterpret_cast<int8_t*>(blob + L.off_cstep8 + ((size_t)j * 2 * L.npad + tid) * 16);
            int8_t* g1 = g0 + (size_t)L.npad * 16;
            for (int e = 0; e < 31; ++e) {
                const int d = S > 127 ? 127 : (S < -127 ? -127 : S);
                S -= d;
                (e < 16 ? g0 : g1)[e & 15] = (int8_t)d;
            }
            g1[15] = (int8_t)r;
        }
    }
    {
        const int a = mycq < 0 ? -mycq : mycq;
        const int need = a == 0 ? 1 : (a + CP_I8_CSTEP_CAP - 1) / CP_I8_CSTEP_CAP;
        atomicMax(&s_h.ncsteps, need);
    }
    // weights: B operand digits, the int16 table of the refinement, the float64 log tables in reference order
    const int total = L.npad * L.nsteps * 16;
    for (int i = tid; i < total; i += 256) {
        const int n = i / (L.nsteps * 16);
        const int rem = i - n * (L.nsteps * 16);
        const int s = rem >> 4, e = rem & 15;
        const int q = 2 * s + (e >> 3), z = e & 7;
        int qv = 0;
        if (q < L.nq && z < E) {
            if (n < L.K) {
                const size_t idx = ((size_t)n * L.nq + q) * E + z;
                const long long v = llrint((lp[idx] - lip[idx]) / scale);
                qv = (int)(v > qmax ? qmax : (v < -qmax ? -qmax : v));
            } else if (n == L.npad - 2) {
                qv = 1;
            } else if (n == L.npad - 1) {
                qv = (mask && mask[q]) ? 1 : 0;
            }
        }
        const int d1 = (qv >= 0 ? qv + 127 : qv - 127) / 255, d0 = qv - 255 * d1;
        int8_t* row = reinterpret_cast<int8_t*>(blob + L.off_q + (((size_t)s * 2 + (e >> 3)) * L.npad + n) * 16);
        row[z] = (int8_t)d0;
        row[8 + z] = (int8_t)d1;
        if (q < L.nq) reinterpret_cast<int16_t*>(blob + L.off_qtab)[((size_t)q * 8 + z) * L.npad + n] = (int16_t)qv;
    }
    for (int i = tid; i < L.K * per; i += 256) {
        reinterpret_cast<double*>(blob + L.off_lp64)[i] = lp[i];
        reinterpret_cast<double*>(blob + L.off_lip64)[i] = lip[i];
    }
    __syncthreads();
    if (tid == 0) *reinterpret_cast<PartsHeaderI8*>(blob + L.off_hdr8) = s_h;
}

struct SpreadConst {      // per-launch constants of the fused spreading (divisions by F and block_len as multiply-high)
    uint32_t magicF, magicB;
    int T, F, block_len, tfull, tail_lo, last;
};
__host__ inline SpreadConst spread_const(int T, int F, int block_len) {
    SpreadConst c;
    c.T = T; c.F = F; c.block_len = block_len;
    c.magicF = (uint32_t)(0xFFFFFFFFull / (unsigned)F) + 1u;
    c.magicB = block_len > 0 ? (uint32_t)(0xFFFFFFFFull / (unsigned)block_len) + 1u : 0u;
    c.tfull = block_len > 0 ? (T / block_len) * block_len : T;
    c.tail_lo = block_len > 0 && T > block_len ? T - block_len : 0;
    c.last = T * F - 1;
    return c;
}
struct CodeArgs {
    const uint8_t* bits;
    const uint8_t* blob;
    uint8_t* codes;
    float* scores;
    uint16_t* counts;
    int B, T, F, To, Fo;
    int threshold, mask_threshold;
    int spread_radius, block_len;   // fused edge spreading in the producers (radius 1, unit direction table); 0 = input used as is
    SpreadConst sc;
    PartsLayout L;
    // int8 form, the refinement's work lists: ws[0] = undecided locations with exactly two candidates (entries of two
    // words from ws[4]: location, part | part << 8; at most ws_cap), ws[1] = the other undecided locations (one word
    // each, after the entries)
    unsigned int* ws;
    unsigned int ws_cap;
    int dbg;            // debug experiments: 1 = epilogue skips TMEM reads, 2 = MMA warp skips tcgen05.mma
    long long* trace;   // optional debug timeline of CTA 0: [tile][8] clock64 stamps (tools/trace_code_parts.py)
};
// The timeline stamps and debug switches are compiled in only with -DTSR_DEBUG_TRACE (tools/trace_code_parts.py needs
// it): even predicated-off they sit in the issue loops, where every instruction shows up in the kernel time.
#ifdef TSR_DEBUG_TRACE
#define CP_DBG(x) (A.dbg & (x))
#define CP_TRACE_TILES 96
#define CP_TRACE(slot)                                                                              \
    do {                                                                                            \
        if (A.trace && blockIdx.x == 0 && it < CP_TRACE_TILES && lane == 0) A.trace[it * 8 + (slot)] = clock64(); \
    } while (0)
#else
#define CP_DBG(x) 0
#define CP_TRACE_TILES 96
#define CP_TRACE(slot) do { } while (0)
#endif

// K1a fused into the producers: radius-1 line spread of the packed edge byte at flattened position p of one utterance
// (edge_signal_proc.py:1397-1417 with the unit direction table of :1751-1821: orientations 0,1 along time, 2,3 along
// frequency, 4,5 along the diagonal, 6,7 along the anti-diagonal), applied inside each block_len-frame block as
// _edge_map_threshold_segments does (:1617-1663); the frames of a ragged last block belong to the window of the final
// block_len frames (:1595-1633).  Nine byte loads per position instead of one, no extra launch and no second pass
// over the edge map.
// x / d for x < 2^31 with magic = floor((2^32 - 1) / d) + 1: the multiply-high is x / d or one more; one fix-up
__device__ __forceinline__ int fast_div(int x, int d, uint32_t magic, int& rem) {
    if (magic == 0u) { rem = 0; return x; }               // d == 1 (2^32 / 1 does not fit the 32-bit magic)
    int q = (int)__umulhi((uint32_t)x, magic);
    rem = x - q * d;
    if (rem < 0) { rem += d; --q; }
    return q;
}
__device__ __forceinline__ uint32_t spread1_byte(const uint8_t* __restrict__ src, int p, const SpreadConst& C) {
    int f, r;
    const int t = fast_div(p, C.F, C.magicF, f);
    bool up, dn;
    if (C.block_len <= 0) { up = t > 0; dn = t + 1 < C.T; }
    else if (t >= C.tfull) { up = t - 1 >= C.tail_lo; dn = t + 1 < C.T; }        // ragged tail: window of the final block_len frames
    else { fast_div(t, C.block_len, C.magicB, r); up = r != 0; dn = r != C.block_len - 1; }
    // All nine loads are unconditional (addresses kept inside the utterance with four clamps) so that they issue back
    // to back; the validity of a neighbour is applied as a mask afterwards.
    const uint32_t mu = up ? 0xFFu : 0u, md = dn ? 0xFFu : 0u, ml = f > 0 ? 0xFFu : 0u, mr = f + 1 < C.F ? 0xFFu : 0u;
    const int pu = up ? p - C.F : p, pd = dn ? p + C.F : p;          // a valid neighbour row keeps its true address
    const uint32_t c = src[p], u = src[pu], d = src[pd], l = src[max(p - 1, 0)], rr = src[min(p + 1, C.last)];
    const uint32_t ul = src[max(pu - 1, 0)], ur = src[min(pu + 1, C.last)], dl = src[max(pd - 1, 0)], dr = src[min(pd + 1, C.last)];
    return c | (((u & mu) | (d & md)) & 0x03u) | (((l & ml) | (rr & mr)) & 0x0Cu) | (((ul & mu & ml) | (dr & md & mr)) & 0x30u) |
           (((dl & md & ml) | (ur & mu & mr)) & 0xC0u);
}

// --------------------------------------------------------------------------- SIMT path
// One warp per location, lanes over parts; walks the set bits of the window in (a,b,z) order and
// adds fp32 W rows from shared memory.  Cross-check / comparison kernel (per-stage choice).
__global__ void __launch_bounds__(256) code_parts_simt_kernel(CodeArgs A) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const PartsLayout& L = A.L;
    float* wf = reinterpret_cast<float*>(smem);                 // [nq*8][npad]
    float* cs = wf + (size_t)L.nq * 8 * L.npad;                 // [npad]
    {
        const float* gw = reinterpret_cast<const float*>(A.blob + L.off_wf);
        const float* gc = reinterpret_cast<const float*>(A.blob + L.off_c);
        for (int i = threadIdx.x; i < L.nq * 8 * L.npad; i += blockDim.x) wf[i] = gw[i];
        for (int i = threadIdx.x; i < L.npad; i += blockDim.x) cs[i] = gc[i];
    }
    const unsigned long long maskbits = *reinterpret_cast<const unsigned long long*>(A.blob + L.off_mask);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long nloc = (long long)A.B * A.To * A.Fo;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long loc = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; loc < nloc; loc += nwarps) {
        int j = (int)(loc % A.Fo);
        long long tmp = loc / A.Fo;
        int i = (int)(tmp % A.To);
        int b = (int)(tmp / A.To);
        const uint8_t* base = A.bits + ((size_t)b * A.T + i) * A.F + j;
        float acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = (lane + 32 * u < L.K) ? cs[lane + 32 * u] : 0.f;
        int count = 0, mcount = 0;
        for (int q = 0; q < L.nq; ++q) {
            int a = q / L.pw, c = q - a * L.pw;
            uint32_t byte = base[(size_t)a * A.F + c];          // warp-uniform address: broadcast load
            int pc = __popc(byte);
            count += pc;
            if ((maskbits >> q) & 1ull) mcount += pc;
            while (byte) {
                int z = __ffs(byte) - 1;
                byte &= byte - 1;
                const float* wrow = wf + (size_t)(q * 8 + z) * L.npad;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (32 * u < L.K) {
                        int k = lane + 32 * u;
                        if (k < L.K) acc[u] = __fadd_rn(acc[u], wrow[k]);
                    }
            }
        }
        // arg-max, lowest k wins ties
        float best = -INFINITY;
        int bestk = 0x7fffffff;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            int k = lane + 32 * u;
            if (k < L.K && acc[u] > best) { best = acc[u]; bestk = k; }
        }
        for (int off = 16; off > 0; off >>= 1) {
            float ob = __shfl_xor_sync(0xffffffffu, best, off);
            int ok = __shfl_xor_sync(0xffffffffu, bestk, off);
            if (ob > best || (ob == best && ok < bestk)) { best = ob; bestk = ok; }
        }
        bool gate = count >= A.threshold && (A.mask_threshold < 0 || mcount > A.mask_threshold);
        if (lane == 0) {
            if (A.codes) A.codes[loc] = gate ? (uint8_t)(bestk + 1) : 0;
            if (A.counts) A.counts[loc] = (uint16_t)count;
        }
        if (A.scores) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                int k = lane + 32 * u;
                if (k < L.K) A.scores[loc * L.K + k] = acc[u];
            }
        }
    }
}

// ------------------------------------------------------------------------- tensor path
#define CP_NSTAGE 4
#define CP_NISSUE 4
#define CP_THREADS (384 + 32 * CP_NISSUE)   // hi+lo form: warps 0-3 / 4-7 epilogue (even / odd tiles), 8-11 producers, 12.. MMA issuers (tile % CP_NISSUE; highest ids win arbitration)
// Roles (the same for both forms).  Measured for the int8 form, which issues 14 MMAs per tile instead of 27 and is bound
// by its epilogue: two producer warps + two issuers with the other four warps retired (fewer polling warps) 1.09 ms
// against 0.72; a third epilogue set in their place 1.07 ms -- a tile's MMAs leave ONE thread at ~125 cycles each, so
// the four issuers (four tiles in flight) are what overlaps them.
template <bool I8> struct CodeRoles {
    static constexpr int kEpiSets = 2;
    static constexpr int kProdWarp0 = 4 * kEpiSets;
    static constexpr int kProdWarps = 4;
    static constexpr int kIssueWarp0 = kProdWarp0 + kProdWarps;
    static constexpr int kIssueWarps = CP_NISSUE;
    static constexpr int kProdSlots = 512 / (32 * kProdWarps);      // positions a producer thread keeps in registers
    static_assert(32 * (kIssueWarp0 + kIssueWarps) <= CP_THREADS, "roles must fit the CTA");
};
#define CP_MAX_STEPS 32
#define CP_NACC 4         // barrier slots; TMEM accumulators in use: 4 x 128 columns when N_pad <= 128, else 2 x 256

struct CodeSmem {   // offsets into dynamic shared memory
    uint32_t w, cstep, ones, a, bars, stage_bytes, npos, total;
};
__host__ __device__ inline CodeSmem code_smem_layout(const PartsLayout& L, int F, bool i8 = false) {
    CodeSmem S;
    S.w = 0;
    S.cstep = (uint32_t)((i8 ? 1 : 2) * L.nsteps * L.npad * 32);   // the c[k] K-step(s), contiguous with the planes in the blob
    S.ones = (S.cstep + (i8 ? CP_MAX_CSTEPS : 1) * L.npad * 32 + 127) & ~127u;   // their constant A operand (128 rows, 4 KB)
    S.a = S.ones + 4096;
    S.npos = 128 + (L.ph - 1) * F + (L.pw - 1) + 2;
    S.stage_bytes = (S.npos * 16 + 127) & ~127u;
    S.bars = S.a + CP_NSTAGE * S.stage_bytes;
    S.total = S.bars + 192 + CP_MAX_STEPS * 4;   // barriers | generic-path A descriptor words
    return S;
}

// (utterance, tile inside the utterance) of a tile index that advances by a constant stride: no 64-bit divisions in the
// per-tile loops
struct TileWalk {
    int b, q, sb, sq, tpu;
    __device__ __forceinline__ TileWalk(long long tile0, long long stride, int tiles_per_utt) : tpu(tiles_per_utt) {
        b = (int)(tile0 / tpu); q = (int)(tile0 % tpu); sb = (int)(stride / tpu); sq = (int)(stride % tpu);
    }
    __device__ __forceinline__ void next() {
        b += sb; q += sq;
        if (q >= tpu) { q -= tpu; ++b; }
    }
};

// NSTEPS_CT > 0: K-step count known at compile time (13 for the 5x5 parts of every reference config): the
// issue loop is fully unrolled with the descriptor words in registers.  NSTEPS_CT == 0: generic, smem table.
// I8: the int8 form (prepare_parts_i8_kernel): kind::i8 MMAs of K = 32 -- both weight digits of two positions per
// MMA, half the MMAs of the 16-bit hi+lo form at the same cycles each -- exact int32 accumulators, integer arg-max
// with the runner-up tracked; locations whose runner-up is inside the rounding bound are appended to a work list and
// settled in float64 by code_parts_refine_kernel.
template <int NSTEPS_CT, bool I8>
__global__ void __launch_bounds__(CP_THREADS, 1) code_parts_tc_kernel(CodeArgs A, int tiles_per_utt, long long ntiles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    using R = CodeRoles<I8>;
    const PartsLayout& L = A.L;
    const CodeSmem S = code_smem_layout(L, A.F, I8);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S.bars);       // [CP_NSTAGE]
    uint64_t* empty_bar = full_bar + CP_NSTAGE;                            // [CP_NSTAGE]
    uint64_t* tfull_bar = empty_bar + CP_NSTAGE;                           // [CP_NACC]
    uint64_t* tempty_bar = tfull_bar + CP_NACC;                            // [CP_NACC]
    uint64_t* w_bar = tempty_bar + CP_NACC;                                // [1]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_bar + 1);

    // warp index through a broadcast shuffle: the compiler then knows the role branches are warp-uniform and keeps the
    // issuers' descriptor arithmetic in uniform registers (no R2UR moves, no elect on the critical issue path)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t fmt = reinterpret_cast<const uint32_t*>(A.blob + L.off_hdr)[0];
    const uint32_t nacc = L.npad <= 128 ? 4u : 2u, acc_cols = 512u / nacc;

    if (threadIdx.x == 0) {
        for (int s = 0; s < CP_NSTAGE; ++s) { mbar_init(&full_bar[s], 32 * R::kProdWarps); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < CP_NACC; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 128); }
        mbar_init(w_bar, 1);
        mbar_fence_init();
    }
    if (warp == R::kIssueWarp0) {
        tmem_alloc<512>(tmem_ptr);
    }
    {   // constant A operand of the c[k] step: k-group 0 rows = [1,1,1,0,0,0,0,0], k-group 1 (2 KB further) = 0
        // (int8 form: 31 elements of 255 and a last one of 1)
        const uint32_t one = fmt_one(fmt);
        for (int i = threadIdx.x; i < 256; i += blockDim.x) {
            if (I8) reinterpret_cast<uint4*>(smem + S.ones)[i] = make_uint4(~0u, ~0u, ~0u, i < 128 ? ~0u : 0x01FFFFFFu);
            else reinterpret_cast<uint4*>(smem + S.ones)[i] = i < 128 ? make_uint4(one | (one << 16), one, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // weights + c: bulk (TMA) copies, landed once per CTA
    const PartsHeaderI8* h8 = reinterpret_cast<const PartsHeaderI8*>(A.blob + L.off_hdr8);
    const int ncsteps = I8 ? h8->ncsteps : 1;
    const uint32_t wbytes = S.cstep + ncsteps * L.npad * 32;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(w_bar, wbytes);
        const uint32_t half = (uint32_t)(L.nsteps * L.npad * 32);
        if (I8) {
            bulk_g2s(smem + S.w, A.blob + L.off_q, half, w_bar);
            bulk_g2s(smem + S.cstep, A.blob + L.off_cstep8, ncsteps * L.npad * 32, w_bar);
        } else {
            bulk_g2s(smem + S.w, A.blob + L.off_w, half, w_bar);
            bulk_g2s(smem + S.w + half, A.blob + L.off_w + half, half, w_bar);
            bulk_g2s(smem + S.cstep, A.blob + L.off_cstep, L.npad * 32, w_bar);
        }
    }

    if (warp >= R::kProdWarp0 && warp < R::kIssueWarp0) {
        // ================================ producers ================================
        const int pt = threadIdx.x - 32 * R::kProdWarp0;
        constexpr int PT = 32 * R::kProdWarps, NJ = R::kProdSlots;
        const uint32_t one = fmt_one(fmt);
        const int plimit = A.T * A.F;
        const int npos = (int)S.npos;
        uint32_t it = 0;
        // The packed bytes of a tile are fetched TWO tiles ahead into alternating register sets: with one set the global
        // load latency (~900 cycles) sat inside every iteration, which is what bounds the int8 form once its MMAs take
        // half the time.
        uint32_t preA[NJ], preB[NJ];
        const uint32_t nap = 40u;
        TileWalk wf(blockIdx.x, gridDim.x, tiles_per_utt);       // the tile being fetched
        TileWalk wt(blockIdx.x, gridDim.x, tiles_per_utt);       // the tile being written
        auto fetch = [&](uint32_t (&pre)[NJ], long long tile) {
            if (tile < ntiles) {
                const int p0 = wf.q * 128;
                const uint8_t* src = A.bits + (size_t)wf.b * A.T * A.F;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    int pos = pt + PT * j, p = p0 + pos;
                    pre[j] = (pos < npos && p < plimit) ? (A.spread_radius ? spread1_byte(src, p, A.sc) : (uint32_t)src[p]) : 0u;
                }
            }
            wf.next();
        };
        auto put = [&](const uint32_t (&pre)[NJ]) {
            const int stage = it % CP_NSTAGE;
            const uint32_t ph = (it / CP_NSTAGE) & 1;
            mbar_wait_relaxed(&empty_bar[stage], ph ^ 1, nap);
            if (warp == R::kProdWarp0 && !CP_DBG(4)) CP_TRACE(0);
            uint8_t* dst = smem + S.a + stage * S.stage_bytes;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                int pos = pt + PT * j;
                if (pos < npos) *reinterpret_cast<uint4*>(dst + pos * 16) = I8 ? expand_byte_u8x2(pre[j]) : expand_byte(pre[j], one);
            }
            if (npos > 512) {   // very wide inputs: remaining positions straight from global
                const int p0 = wt.q * 128;
                const uint8_t* src = A.bits + (size_t)wt.b * A.T * A.F;
                for (int pos = pt + 512; pos < npos; pos += PT) {
                    int p = p0 + pos;
                    const uint32_t byte = p < plimit ? (A.spread_radius ? spread1_byte(src, p, A.sc) : (uint32_t)src[p]) : 0u;
                    *reinterpret_cast<uint4*>(dst + pos * 16) = I8 ? expand_byte_u8x2(byte) : expand_byte(byte, one);
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(&full_bar[stage]);
            if (warp == R::kProdWarp0 && !CP_DBG(4)) CP_TRACE(1);
            wt.next();
            ++it;
        };
        long long tile = blockIdx.x;
        fetch(preA, tile);
        fetch(preB, tile + gridDim.x);
        while (tile < ntiles) {
            put(preA);
            fetch(preA, tile + 2ll * gridDim.x);
            tile += gridDim.x;
            if (tile >= ntiles) break;
            put(preB);
            fetch(preB, tile + 2ll * gridDim.x);
            tile += gridDim.x;
        }
    } else if (warp >= R::kIssueWarp0 + R::kIssueWarps) {
        // spare warps (int8 form): nothing to do
    } else if (warp >= R::kIssueWarp0) {
        // ================================ MMA issuers ================================
        // One thread can issue a tcgen05.mma only every ~75-90 cycles (tools/ts_probe.cu: 90.7 cycles per MMA from
        // one warp, 60.4 from two, against 59 cycles of tensor work at N = 112), so two warps on different SMSPs
        // issue alternate tiles into their own accumulators; each tile's MMAs stay in one thread's program order.
        // The whole warp runs this loop in lock step; one elected lane issues each tcgen05 op.  The
        // issuing warp executes ~10 instructions per MMA: at 56 tensor cycles per MMA anything more
        // (divisions, table walks, 64-way predicated unrolls) makes the issue loop the bottleneck.
        //   A low word = (patch group offset in 16-byte units) | LBO << 16   (+ stage base per tile)
        //   B low word = plane/step address >> 4 | LBO << 16
        mbar_wait(w_bar, 0);
        const uint32_t idesc = I8 ? ((2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(L.npad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24))   // s32 += u8 x s8
                                  : umma_idesc_f32acc(fmt, 128, L.npad);
        const uint64_t desc_hi = (uint64_t)((uint32_t)(128 >> 4) | (1u << 14)) << 32;   // SBO = 128 B, version 1
        const uint32_t b_base = (smem_u32(smem + S.w) >> 4) | ((uint32_t)L.npad << 16);
        const uint32_t b_step = 2 * L.npad;
        const uint64_t ones_desc = umma_desc_kmajor_noswz(smem_u32(smem + S.ones), 2048, 128);
        const uint64_t cstep_desc = umma_desc_kmajor_noswz(smem_u32(smem + S.cstep), (uint32_t)L.npad * 16, 128);
        constexpr int NS = NSTEPS_CT > 0 ? NSTEPS_CT : 1;
        uint32_t a_lo[NS];
        uint32_t* a_tab = reinterpret_cast<uint32_t*>(smem + S.bars + 192);   // generic path: [nsteps]
        {
            int pa = 0, pb = 0;
            for (int s = 0; s < L.nsteps; ++s) {
                const uint32_t off0 = pa * A.F + pb;
                int pb1 = pb + 1, pa1 = pa;
                if (pb1 == L.pw) { pb1 = 0; ++pa1; }
                const uint32_t off1 = (2 * s + 1 < L.nq) ? pa1 * A.F + pb1 : off0 + 1;
                const uint32_t w = off0 | ((off1 - off0) << 16);
                if (NSTEPS_CT > 0) {
#pragma unroll
                    for (int k = 0; k < NS; ++k)
                        if (k == s) a_lo[k] = w;
                } else if (lane == 0) {
                    a_tab[s] = w;
                }
                pb = pb1 + 1; pa = pa1;
                if (pb == L.pw) { pb = 0; ++pa; }
            }
            __syncwarp();
        }
        // With two 256-column accumulators (N_pad > 128) only two issuers run: an accumulator must belong to ONE issuer,
        // whose program order is what makes its tempty parity waits see every second release.
        const int iss = warp - R::kIssueWarp0, nissue = nacc == 4u ? R::kIssueWarps : 2;
        uint32_t it = iss;
        for (long long tile = iss < nissue ? blockIdx.x + (long long)iss * gridDim.x : ntiles; tile < ntiles; tile += (long long)nissue * gridDim.x, it += nissue) {
            const int stage = it % CP_NSTAGE;
            const uint32_t ph = (it / CP_NSTAGE) & 1;
            const int acc = it % nacc;
            const uint32_t aph = (it / nacc) & 1;
            mbar_wait(&full_bar[stage], ph);
            if (!CP_DBG(4)) CP_TRACE(2);
            mbar_wait(&tempty_bar[acc], aph ^ 1);
            if (!CP_DBG(4)) CP_TRACE(3);
            tc_fence_after();
            const uint32_t a_base = smem_u32(smem + S.a + stage * S.stage_bytes) >> 4;
            const uint32_t d_tmem = tmem_base + acc * acc_cols;
            if (!CP_DBG(2)) {
                // accumulator := c[k] (constant-ones A x the c step), then the 2 x nsteps patch steps accumulate
                uint32_t b_lo = b_base;
                if (I8) {
                    for (int cs = 0; cs < ncsteps; ++cs)
                        umma_i8_elect(d_tmem, ones_desc, cstep_desc + (uint64_t)(cs * b_step), idesc, cs ? 1u : 0u);
                    if (NSTEPS_CT > 0) {
#pragma unroll
                        for (int s = 0; s < NS; ++s) {
                            umma_i8_elect(d_tmem, desc_hi | (uint64_t)(a_lo[s] + a_base), desc_hi | (uint64_t)b_lo, idesc, 1u);
                            b_lo += b_step;
                        }
                    } else {
                        for (int s = 0; s < L.nsteps; ++s) {
                            umma_i8_elect(d_tmem, desc_hi | (uint64_t)(a_tab[s] + a_base), desc_hi | (uint64_t)b_lo, idesc, 1u);
                            b_lo += b_step;
                        }
                    }
                } else {
                    umma_f16_elect(d_tmem, ones_desc, cstep_desc, idesc, 0u);
                    if (NSTEPS_CT > 0) {
#pragma unroll
                        for (int plane = 0; plane < 2; ++plane)
#pragma unroll
                            for (int s = 0; s < NS; ++s) {
                                umma_f16_elect(d_tmem, desc_hi | (uint64_t)(a_lo[s] + a_base), desc_hi | (uint64_t)b_lo, idesc, 1u);
                                b_lo += b_step;
                            }
                    } else {
                        for (int plane = 0; plane < 2; ++plane)
                            for (int s = 0; s < L.nsteps; ++s) {
                                umma_f16_elect(d_tmem, desc_hi | (uint64_t)(a_tab[s] + a_base), desc_hi | (uint64_t)b_lo, idesc, 1u);
                                b_lo += b_step;
                            }
                    }
                }
            }
            umma_commit_elect(&empty_bar[stage]);   // A stage reusable once these MMAs retire
            umma_commit_elect(&tfull_bar[acc]);     // accumulator ready for the epilogue
            if (!CP_DBG(4)) CP_TRACE(4);
        }
    } else {
        // ================================ epilogue ================================
        // two warp sets: warps 0-3 drain accumulator 0 (even tiles), warps 4-7 accumulator 1 (odd tiles)
        const int eset = warp >> 2;
        constexpr int NSETS = R::kEpiSets;
        const int e = (warp & 3) * 32 + lane;   // TMEM lane == tile row (lane quadrant = warp % 4)
        uint32_t it = eset;
        TileWalk we(blockIdx.x + (long long)eset * gridDim.x, (long long)NSETS * gridDim.x, tiles_per_utt);
        for (long long tile = blockIdx.x + (long long)eset * gridDim.x; tile < ntiles; tile += (long long)NSETS * gridDim.x, it += NSETS, we.next()) {
            const int acc = it % nacc;             // this set owns accumulators eset and (with four) eset + 2
            const uint32_t aph = (it / nacc) & 1;
            const int b = we.b;
            const int p = we.q * 128 + e;
            int f;
            const int t = fast_div(p, A.F, A.sc.magicF, f);
            const bool valid = t < A.To && f < A.Fo;
            const size_t loc = ((size_t)b * A.To + t) * A.Fo + f;
            mbar_wait_relaxed(&tfull_bar[acc], aph);
            if ((warp & 3) == 0) CP_TRACE(5);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * acc_cols + ((uint32_t)((warp & 3) * 32) << 16);
            // four independent running (max, argmax) chains over columns n = 4j + c of the K part columns (c[k] is
            // already in the accumulator).  Strict '>' within a chain keeps the lowest index; the merge breaks ties by
            // index.  No shared-memory access in here: LDS stalls for hundreds of cycles behind the MMA operand reads.
            float bv[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            int bk[4] = {0, 1, 2, 3};
            // int8 form: keys (acc << 8) | (255 - k) order like (acc, lowest k first); best and runner-up per chain
            int m1[4] = {INT_MIN, INT_MIN, INT_MIN, INT_MIN}, m2[4] = {INT_MIN, INT_MIN, INT_MIN, INT_MIN};
            float fcount = 0.f, fmcount = 0.f;
            int icount = 0, imcount = 0;
            float* srow = A.scores + loc * L.K;      // drop-in path only: the full score map
            // All TMEM loads of a 64-column block are issued back to back and waited for once.
            for (int nb = 0; nb < (CP_DBG(1) ? 0 : L.npad); nb += 64) {
                uint32_t v[64];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (nb + 16 * j < L.npad) tmem_ld16(taddr + nb + 16 * j, *reinterpret_cast<uint32_t(*)[16]>(&v[16 * j]));
                tmem_ld_wait();
                if (CP_DBG(4) && (warp & 3) == 0) CP_TRACE(nb ? 3 : 1);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = nb + 16 * j;
                    if (I8) {
                        if (n + 16 <= L.K) {                  // whole 16-column group of parts
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int key = (int)v[16 * j + i] * 256 + (255 - n - i);
                                m2[i & 3] = max(m2[i & 3], min(m1[i & 3], key));
                                m1[i & 3] = max(m1[i & 3], key);
                            }
                        } else if (n < L.K) {                 // the group that holds the last parts
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int key = n + i < L.K ? (int)v[16 * j + i] * 256 + (255 - n - i) : INT_MIN;
                                m2[i & 3] = max(m2[i & 3], min(m1[i & 3], key));
                                m1[i & 3] = max(m1[i & 3], key);
                            }
                        }
                        if (A.scores && valid && n < L.K) {   // 16-bit fixed-point scores (the 16-bit float planes carry 22 bits)
                            const double scale = h8->scale;
                            const int cmid = h8->cmid;
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (n + i < L.K) srow[n + i] = (float)((double)((int)v[16 * j + i] + cmid) * scale);
                        }
                        if (n + 16 == L.npad) { icount = (int)v[16 * j + 14]; imcount = (int)v[16 * j + 15]; }
                        continue;
                    }
                    if (n + 16 <= L.K) {                      // whole 16-column group of parts
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float sc = __uint_as_float(v[16 * j + i]);
                            if (sc > bv[i & 3]) { bv[i & 3] = sc; bk[i & 3] = n + i; }
                        }
                    } else if (n < L.K) {                     // the group that holds the last parts
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float sc = __uint_as_float(v[16 * j + i]);
                            if (n + i < L.K && sc > bv[i & 3]) { bv[i & 3] = sc; bk[i & 3] = n + i; }
                        }
                    }
                    if (A.scores && valid && n < L.K) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (n + i < L.K) srow[n + i] = __uint_as_float(v[16 * j + i]);
                    }
                    // exact integer edge counts from the all-ones / mask weight columns: the last two columns of the
                    // padded range, i.e. static register indices (a data-dependent index would push v[] to local memory)
                    if (n + 16 == L.npad) { fcount = __uint_as_float(v[16 * j + 14]); fmcount = __uint_as_float(v[16 * j + 15]); }
                }
                if (CP_DBG(4) && (warp & 3) == 0) CP_TRACE(nb ? 4 : 2);
            }
            float best = bv[0];
            int bestk = bk[0];
            if (I8) {
                // merge the four (best, runner-up) pairs; the arg-max is certain only if the runner-up's accumulator is
                // more than (set bits + 1) units behind (see prepare_parts_i8_kernel; one more unit of slack for the
                // float64 rounding of the reference's own sums)
                // (top three of the eight values: the third decides whether the first two are the only candidates)
                const int alo = min(m1[0], m1[1]), ac = max(m2[0], m2[1]), blo = min(m1[2], m1[3]), bc = max(m2[2], m2[3]);
                const int a1 = max(m1[0], m1[1]), a2 = max(alo, ac), a3 = max(min(alo, ac), min(m2[0], m2[1]));
                const int b1 = max(m1[2], m1[3]), b2 = max(blo, bc), b3 = max(min(blo, bc), min(m2[2], m2[3]));
                const int lo1 = min(a1, b1), c2 = max(a2, b2);
                const int t1 = max(a1, b1), t2 = max(lo1, c2), t3 = max(max(min(lo1, c2), min(a2, b2)), max(a3, b3));
                bestk = 255 - (t1 & 255);
                const int vbest = t1 >> 8, bnd = icount + 2;
                const bool gate8 = icount >= A.threshold && (A.mask_threshold < 0 || imcount > A.mask_threshold);
                if (A.ws != nullptr && valid && gate8 && vbest - (t2 >> 8) <= bnd) {
                    // Undecided.  If the third value is outside the bound and the first two come from different chains,
                    // they are provably the only candidates (anything a chain did not keep is below its runner-up): the
                    // refinement gets the pair.  Otherwise it gets the bare location and recomputes the candidate set.
                    const bool pair_only = vbest - (t3 >> 8) > bnd && ((t1 ^ t2) & 3) != 0;
                    const uint32_t idx = pair_only ? atomicAdd(A.ws, 1u) : A.ws_cap;
                    if (idx < A.ws_cap) {
                        A.ws[4 + 2 * (size_t)idx] = (uint32_t)loc;
                        A.ws[5 + 2 * (size_t)idx] = (uint32_t)(255 - (t1 & 255)) | ((uint32_t)(255 - (t2 & 255)) << 8);
                    } else {
                        A.ws[4 + 2 * (size_t)A.ws_cap + atomicAdd(A.ws + 1, 1u)] = (uint32_t)loc;
                    }
                }
            } else {
#pragma unroll
                for (int c = 1; c < 4; ++c)
                    if (bv[c] > best || (bv[c] == best && bk[c] < bestk)) { best = bv[c]; bestk = bk[c]; }
            }
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
            if ((warp & 3) == 0) CP_TRACE(6);
            if (valid) {
                const int count = I8 ? icount : (int)(fcount + 0.5f), mcount = I8 ? imcount : (int)(fmcount + 0.5f);
                const bool gate = count >= A.threshold && (A.mask_threshold < 0 || mcount > A.mask_threshold);
                if (A.codes) A.codes[loc] = gate ? (uint8_t)(bestk + 1) : 0;
                if (A.counts) A.counts[loc] = (uint16_t)count;
            }
            if ((warp & 3) == 0) CP_TRACE(7);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == R::kIssueWarp0) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// --------------------------------------------------------------- float64 refinement of the int8 form
// Settles the locations code_parts_tc_kernel<.., true> could not decide inside the rounding bound.  Its epilogue
// appends them to the call's work lists: with the two parts involved where those are provably the only ones within the
// bound of the best accumulator -- the reference's winner is one of them --, as a bare location otherwise.  Warp w takes
// entries w, w + nwarps, ..., queues (location, part) pairs and
// evaluates the queue one pair per lane with the reference's own arithmetic: the float64 sum of log_parts /
// log_invparts over (a, b, z) in that order starting from 0.0 (code_parts.pyx:123-131), first maximum wins (np.argmax,
// get_train_data.py:2621).  The codes of the int8 form are therefore the reference's codes, with no near-tie exception.
// For a bare location the warp first recomputes the exact integer accumulators from the int16 table and keeps every
// part within the bound.
#define CPR_WARPS 8
#define CPR_SLOTS 16
struct RefineWarp {
    uint32_t loc[CPR_SLOTS];
    uint32_t pair[32];                 // slot << 8 | part
    uint8_t pbytes[CPR_SLOTS][64];     // the patch bytes of the queued locations
};
__device__ __forceinline__ double refine_sum(const uint8_t* pb, int k, const double* __restrict__ lp, const double* __restrict__ lip,
                                             int nq, int E) {
    const double* a = lp + (size_t)k * nq * E;
    const double* b = lip + (size_t)k * nq * E;
    double s = 0.0;
    if (E == 8) {
#pragma unroll 4
        for (int q = 0; q < nq; ++q) {
            const uint32_t byte = pb[q];
            double t[8];
#pragma unroll
            for (int z = 0; z < 8; ++z) t[z] = __ldg(((byte >> z) & 1u) ? a + q * 8 + z : b + q * 8 + z);
#pragma unroll
            for (int z = 0; z < 8; ++z) s = __dadd_rn(s, t[z]);
        }
    } else {
        for (int q = 0; q < nq; ++q) {
            const uint32_t byte = pb[q];
            for (int z = 0; z < E; ++z) s = __dadd_rn(s, __ldg(((byte >> z) & 1u) ? a + q * E + z : b + q * E + z));
        }
    }
    return s;
}
__global__ void __launch_bounds__(CPR_WARPS * 32, 4) code_parts_refine_kernel(CodeArgs A) {
    __shared__ RefineWarp s_w[CPR_WARPS];
    const PartsLayout& L = A.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nfull = min(A.ws[0], A.ws_cap), nent = nfull + A.ws[1];
    RefineWarp* W = &s_w[warp];
    const int16_t* qt = reinterpret_cast<const int16_t*>(A.blob + L.off_qtab);
    const int* cq = reinterpret_cast<const int*>(A.blob + L.off_cq);
    const double* lp64 = reinterpret_cast<const double*>(A.blob + L.off_lp64);
    const double* lip64 = reinterpret_cast<const double*>(A.blob + L.off_lip64);
    const int E = reinterpret_cast<const PartsHeaderI8*>(A.blob + L.off_hdr8)->E;
    const unsigned full = 0xffffffffu;
    int nslots = 0, npairs = 0;        // warp-uniform queue state

    auto flush = [&]() {
        if (npairs) {
            double sum = -INFINITY;
            int k = 0x7fffffff, slot = -1;
            if (lane < npairs) {
                const uint32_t pr = W->pair[lane];
                slot = (int)(pr >> 8); k = (int)(pr & 255u);
                sum = refine_sum(W->pbytes[slot], k, lp64, lip64, L.nq, E);
            }
            __syncwarp();
            bool win = lane < npairs;
            for (int j = 0; j < npairs; ++j) {
                const double sj = __shfl_sync(full, sum, j);
                const int slotj = __shfl_sync(full, slot, j), kj = __shfl_sync(full, k, j);
                if (slotj == slot && (sj > sum || (sj == sum && kj < k))) win = false;
            }
            if (win) A.codes[W->loc[slot]] = (uint8_t)(k + 1);
        }
        npairs = 0; nslots = 0;
        __syncwarp();
    };

    const uint32_t gw = blockIdx.x * CPR_WARPS + warp, nw = gridDim.x * CPR_WARPS;
    for (uint32_t ent = gw; ent < nent; ent += nw) {
        const bool have_set = ent < nfull;
        const uint32_t* e = A.ws + 4 + (have_set ? 2 * (size_t)ent : 2 * (size_t)A.ws_cap + (ent - nfull));
        const uint32_t loc = e[0];
        if (nslots == CPR_SLOTS) flush();
        int slot = nslots;
        const int j = (int)(loc % (uint32_t)A.Fo);
        const uint32_t tmp = loc / (uint32_t)A.Fo;
        const int i = (int)(tmp % (uint32_t)A.To);
        const uint8_t* src = A.bits + (size_t)(tmp / (uint32_t)A.To) * A.T * A.F;
        const int p0 = i * A.F + j;
        int cnt = 0;
        for (int q = lane; q < L.nq; q += 32) {
            const int a = q / L.pw, p = p0 + a * A.F + (q - a * L.pw);
            const uint32_t byte = A.spread_radius ? spread1_byte(src, p, A.sc) : (uint32_t)src[p];
            W->pbytes[slot][q] = (uint8_t)byte;
            cnt += __popc(byte);
        }
        __syncwarp();
        // bal[u] bit l: part 32 u + l is a candidate
        uint32_t bal[8];
        if (have_set) {
            const uint32_t kk = e[1], ka = kk & 255u, kb = kk >> 8;
#pragma unroll
            for (int u = 0; u < 8; ++u)     // (a key that never saw a part decodes to part 255: dropped)
                bal[u] = ((ka >> 5) == (uint32_t)u && ka < (uint32_t)L.K ? 1u << (ka & 31u) : 0u) | ((kb >> 5) == (uint32_t)u && kb < (uint32_t)L.K ? 1u << (kb & 31u) : 0u);
        } else {
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(full, cnt, o);
            int acc[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] = 0;
            for (int q = 0; q < L.nq; ++q) {
                uint32_t byte = W->pbytes[slot][q];
                while (byte) {
                    const int z = __ffs(byte) - 1;
                    byte &= byte - 1;
                    const int16_t* row = qt + (size_t)(q * 8 + z) * L.npad;
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (32 * u < L.K) {
                            const int k = lane + 32 * u;
                            if (k < L.K) acc[u] += row[k];
                        }
                }
            }
            int vmax = INT_MIN;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int k = lane + 32 * u;
                if (k < L.K) { acc[u] += cq[k]; vmax = max(vmax, acc[u]); }
            }
            for (int o = 16; o > 0; o >>= 1) vmax = max(vmax, __shfl_xor_sync(full, vmax, o));
#pragma unroll
            for (int u = 0; u < 8; ++u) bal[u] = __ballot_sync(full, lane + 32 * u < L.K && vmax - acc[u] <= cnt + 2);
        }
        int total = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) total += __popc(bal[u]);
        if (total > 32) {               // (nearly) identical parts: evaluate in rounds of 32 without the queue
            flush();
            double best = -INFINITY;
            int bestk = 0x7fffffff;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (bal[u] == 0u) continue;
                double sum = -INFINITY;
                int k = 0x7fffffff;
                if ((bal[u] >> lane) & 1u) { k = lane + 32 * u; sum = refine_sum(W->pbytes[slot], k, lp64, lip64, L.nq, E); }
                __syncwarp();
                for (int o = 16; o > 0; o >>= 1) {
                    const double so = __shfl_xor_sync(full, sum, o);
                    const int ko = __shfl_xor_sync(full, k, o);
                    if (so > sum || (so == sum && ko < k)) { sum = so; k = ko; }
                }
                if (sum > best || (sum == best && k < bestk)) { best = sum; bestk = k; }
            }
            if (lane == 0) A.codes[loc] = (uint8_t)(bestk + 1);
            __syncwarp();
            continue;
        }
        if (npairs + total > 32) {      // no room: evaluate what is queued, this location moves to slot 0
            flush();
            for (int q = lane; q < L.nq; q += 32) W->pbytes[0][q] = W->pbytes[slot][q];
            slot = 0;
            __syncwarp();
        }
        int pos = npairs;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if ((bal[u] >> lane) & 1u) W->pair[pos + __popc(bal[u] & ((1u << lane) - 1u))] = ((uint32_t)slot << 8) | (uint32_t)(lane + 32 * u);
            pos += __popc(bal[u]);
        }
        if (lane == 0) W->loc[slot] = loc;
        npairs = pos; nslots = slot + 1;
        __syncwarp();
    }
    flush();
}

}  // namespace tsr

using namespace tsr;

static long long* g_code_trace = nullptr;
static int g_code_dbg = 0;   // debug hook, see tsr_debug_set_code_trace

extern "C" {

/* Debug only (not in the public header): device buffer of CP_TRACE_TILES*8 int64 clock stamps. */
void tsr_debug_set_code_trace(long long* dev_buf) { g_code_trace = dev_buf; }
void tsr_debug_set_code_flags(int flags) { g_code_dbg = flags; }

size_t tsr_parts_blob_bytes(int K, int ph, int pw) {
    if (K < 1 || ph < 1 || pw < 1) return 0;
    return parts_layout(K, ph, pw).total;
}

int tsr_prepare_parts(const double* log_parts, const double* log_invparts, int K, int ph, int pw, int E,
                      const uint8_t* mask_or_null, int plane_format, void* blob, void* stream) {
    TSR_REQUIRE(log_parts && log_invparts && blob, "tsr_prepare_parts: null pointer");
    TSR_REQUIRE(K >= 1 && K <= 238, "tsr_prepare_parts: need 1 <= num_parts <= 238 (got %d)", K);
    TSR_REQUIRE(ph >= 1 && pw >= 1 && ph * pw <= 64, "tsr_prepare_parts: part window %dx%d unsupported (ph*pw <= 64)", ph, pw);
    TSR_REQUIRE(E >= 1 && E <= 8, "tsr_prepare_parts: need 1 <= E <= 8 (got %d)", E);
    int rc = require_sm100("tsr_prepare_parts");
    if (rc) return rc;
    TSR_REQUIRE(plane_format >= 0 && plane_format <= 2, "tsr_prepare_parts: plane_format must be 0 (auto), 1 (bf16) or 2 (fp16)");
    PartsLayout L = parts_layout(K, ph, pw);
    cudaStream_t st = (cudaStream_t)stream;
    TSR_CUDA(cudaMemsetAsync((uint8_t*)blob + L.off_wf, 0, L.total - L.off_wf, st));
    parts_maxabs_kernel<<<32, 256, 0, st>>>(log_parts, log_invparts, (size_t)K * ph * pw * E, (uint8_t*)blob, L);
    prepare_parts_kernel<<<64, 256, 0, st>>>(log_parts, log_invparts, E, mask_or_null, (uint8_t*)blob, L, plane_format);
    parts_publish_format_kernel<<<1, 1, 0, st>>>((uint8_t*)blob, L, plane_format);
    prepare_parts_i8_kernel<<<1, 256, 0, st>>>(log_parts, log_invparts, E, mask_or_null, (uint8_t*)blob, L);
    return check_launch("tsr_prepare_parts");
}

int tsr_code_parts(const uint8_t* bits, const void* parts_blob, int K, int ph, int pw, int threshold,
                   int mask_threshold, uint8_t* codes, float* scores_or_null, uint16_t* counts_or_null, int B, int T,
                   int F, int edge_spread_radius, int edge_block_len, int impl, void* stream) {
    if (B == 0) return TSR_OK;
    TSR_REQUIRE(bits && parts_blob, "tsr_code_parts: null pointer");
    TSR_REQUIRE(codes || scores_or_null || counts_or_null, "tsr_code_parts: no output requested");
    TSR_REQUIRE(K >= 1 && K <= 238 && ph >= 1 && pw >= 1 && ph * pw <= 64, "tsr_code_parts: unsupported parts shape (K <= 238, ph*pw <= 64)");
    TSR_REQUIRE(B >= 0 && T >= ph && F >= pw, "tsr_code_parts: input (T=%d,F=%d) smaller than the part window (%d,%d)", T, F, ph, pw);
    TSR_REQUIRE(impl == TSR_IMPL_TENSOR || impl == TSR_IMPL_SIMT || impl == TSR_IMPL_TENSOR_INT8, "tsr_code_parts: unknown impl %d", impl);
    TSR_REQUIRE(edge_spread_radius == 0 || (edge_spread_radius == 1 && impl != TSR_IMPL_SIMT && edge_block_len >= 0),
                "tsr_code_parts: the fused edge spreading supports radius 1 on the tensor path (use tsr_spread_edges otherwise)");
    TSR_REQUIRE((long long)T * F < (1ll << 30), "tsr_code_parts: utterance too large (T*F >= 2^30); split in time");
    if (B == 0) return TSR_OK;
    int rc = require_sm100("tsr_code_parts");
    if (rc) return rc;
    CodeArgs A;
    A.bits = bits; A.blob = (const uint8_t*)parts_blob; A.codes = codes; A.scores = scores_or_null; A.counts = counts_or_null;
    A.B = B; A.T = T; A.F = F; A.To = T - ph + 1; A.Fo = F - pw + 1;
    A.threshold = threshold; A.mask_threshold = mask_threshold;
    A.spread_radius = edge_spread_radius; A.block_len = edge_block_len;
    A.sc = spread_const(T, F, edge_block_len);
    A.trace = g_code_trace;
    A.dbg = g_code_dbg;
    A.L = parts_layout(K, ph, pw);
    const int nsm = sm_count();
    if (impl == TSR_IMPL_SIMT) {
        size_t smem = ((size_t)A.L.nq * 8 + 1) * A.L.npad * 4;
        TSR_REQUIRE(smem <= 220 * 1024, "tsr_code_parts(simt): weights do not fit shared memory");
        TSR_CUDA(cudaFuncSetAttribute(code_parts_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int blocks = nsm * (smem <= 100 * 1024 ? 2 : 1);
        code_parts_simt_kernel<<<blocks, 256, smem, (cudaStream_t)stream>>>(A);
        return check_launch("tsr_code_parts(simt)");
    }
    const bool i8 = impl == TSR_IMPL_TENSOR_INT8;
    CodeSmem S = code_smem_layout(A.L, F, i8);
    if (S.total > 227 * 1024) {
        set_error("tsr_code_parts: K=%d parts %dx%d with F=%d needs %u B of shared memory (> 227 KB)", K, ph, pw, F, S.total);
        return TSR_ERR_UNSUPPORTED;
    }
    TSR_REQUIRE((uint32_t)((ph - 1) * F + pw) * 16 < (1u << 18), "tsr_code_parts: F too large for the patch descriptor");
    // int8 form with codes: a stream-ordered work list for the refinement (one entry per location at most)
    unsigned int* ws = nullptr;
    A.ws = nullptr; A.ws_cap = 0;
    if (i8 && codes) {
        const long long nloc = (long long)B * A.To * A.Fo;
        TSR_REQUIRE(nloc < (1ll << 32), "tsr_code_parts(int8): more than 2^32 locations per call; split the batch");
        // two-candidate entries for one location in eight (forty times what log-odds dictionaries produce), then one
        // word per location for whatever does not fit or has more candidates
        A.ws_cap = (unsigned int)(nloc / 8 + 1024);
        {   // keep freed blocks in the device's pool across synchronisations: the default threshold (0) hands them back to
            // the driver at every sync and the next call pays a fresh allocation
            static std::atomic<unsigned long long> pool_kept{0ull};     // one bit per device ordinal
            int dev = 0;
            TSR_CUDA(cudaGetDevice(&dev));
            if (dev >= 0 && dev < 64 && !((pool_kept.load(std::memory_order_relaxed) >> dev) & 1ull)) {
                cudaMemPool_t pool;
                unsigned long long keep = ~0ull;
                if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
                pool_kept.fetch_or(1ull << dev, std::memory_order_relaxed);
            }
        }
        TSR_CUDA(cudaMallocAsync((void**)&ws, (2 * (size_t)A.ws_cap + (size_t)nloc + 4) * sizeof(unsigned int), (cudaStream_t)stream));
        TSR_CUDA(cudaMemsetAsync(ws, 0, 16, (cudaStream_t)stream));
        A.ws = ws;
    }
    const int tiles_per_utt = (int)(((long long)A.To * F + 127) / 128);
    const long long ntiles = (long long)tiles_per_utt * B;
    int blocks = (int)(ntiles < nsm ? ntiles : nsm);
#define CP_LAUNCH(NS, I8)                                                                                                   \
    do {                                                                                                                    \
        TSR_CUDA(cudaFuncSetAttribute(code_parts_tc_kernel<NS, I8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.total)); \
        code_parts_tc_kernel<NS, I8><<<blocks, CP_THREADS, S.total, (cudaStream_t)stream>>>(A, tiles_per_utt, ntiles);      \
    } while (0)
    if (A.L.nsteps == 13) {
        if (i8) CP_LAUNCH(13, true); else CP_LAUNCH(13, false);
    } else {
        if (i8) CP_LAUNCH(0, true); else CP_LAUNCH(0, false);
    }
#undef CP_LAUNCH
    rc = check_launch("tsr_code_parts");
    if (rc && ws) cudaFreeAsync(ws, (cudaStream_t)stream);
    if (rc || !ws) return rc;
    // int8 form: the locations it could not decide inside the rounding bound are settled in float64 (reference arithmetic)
    if (!(g_code_dbg & 8)) {                                      // debug flag 8: skip the refinement
        code_parts_refine_kernel<<<4 * nsm, CPR_WARPS * 32, 0, (cudaStream_t)stream>>>(A);
        rc = check_launch("tsr_code_parts(refine)");
    }
    TSR_CUDA(cudaFreeAsync(ws, (cudaStream_t)stream));
    return rc;
}

}  // extern "C"
