// em.cu -- K4: Bernoulli-mixture EM sufficient statistics on bit-packed data
// (bernoulli_mixture.py:263-291 E-step, :177-193 M-step; bernoulli_em.py:36-49, 103-127).
//
// E-step:  S[n][k] = c[k] + sum_d x[n][d] * W[k][d],  W = logT - logInvT,  c[k] = sum_d logInvT[k][d]
//          -- the same bit x bf16(hi+lo) contraction as detection with L = 1 (detect.cu), followed by
//          a row kernel: ll = S + logw, log-sum-exp, posteriors, column sums, total log-likelihood.
// M-step:  numer[k][d] = sum_n aff[n][k] * x[n][d], accumulated per row split in a fixed order
//          (deterministic), then reduced; the caller all-reduces `stats` across ranks (NCCL).
#include "common.cuh"

extern "C" int tsr_prepare_filters(const float* LF, int M, int L, int D, int plane_format, void* blob, void* stream);
extern "C" size_t tsr_filter_blob_bytes(int M, int L, int D, int plane_format);
extern "C" int tsr_rowwords(int D);
extern "C" int tsr_detect(const uint32_t* featbits, int B, int T, int D, int rowwords, const void* filter_blob,
                          const float* LF_or_null, int M, int L, int plane_format, float* scores, int impl, void* stream);
#define EM_PLANES TSR_PLANES_INT8X3

namespace tsr {

#define EM_MAX_K 64
#define EM_MSPLIT 64

// Tensor-core M-step geometry: the transposed bit matrix XT [D][N_pad] is viewed as D*S rows of
// N_pad/S "features" (row d*S+s = segment s of feature d's column); the posteriors become K*S
// "templates" (template s*K+k = aff[segment s rows][k]); one detection-style contraction (L = 1,
// INT8X3, exact int32 accumulation) yields every (row, template) product and the S useful
// diagonal blocks are gathered into the numerators.  S only exists to spread the contraction
// over enough CTAs; the off-diagonal work it wastes is a few percent of one E-step.
struct EmGeom {
    int S;
    long long Npad, seg;
};
static EmGeom em_geom(long long N, int D, int K) {
    EmGeom g;
    long long want = (2LL * 148 * 256 + D - 1) / D;      // row-steps for ~2 waves of 148 CTAs
    int S = (int)(want < 1 ? 1 : (want > 16 ? 16 : want));
    while (S > 1 && (long long)S * 128 > N) --S;
    g.S = S;
    g.Npad = ((N + (long long)S * 128 - 1) / ((long long)S * 128)) * (long long)S * 128;
    g.seg = g.Npad / S;
    return g;
}

struct EmWorkspace {
    size_t off_w, off_c, off_blob, off_scores, off_partial, off_mlf, off_mblob, off_mscores, total;
};
static EmWorkspace em_workspace(long long N, int D, int K) {
    EmWorkspace w;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const EmGeom g = em_geom(N, D, K);
    w.off_w = 0;
    w.off_c = up((size_t)K * D * 4);
    w.off_blob = w.off_c + up((size_t)K * 4);
    w.off_scores = w.off_blob + up(tsr_filter_blob_bytes(K, 1, D, EM_PLANES));
    w.off_partial = w.off_scores + up((size_t)N * K * 4);
    w.off_mlf = w.off_partial + up((size_t)EM_MSPLIT * K * D * 4);
    w.off_mblob = w.off_mlf + up((size_t)K * g.Npad * 4);
    w.off_mscores = w.off_mblob + up(tsr_filter_blob_bytes(K * g.S, 1, (int)g.seg, TSR_PLANES_INT8X3));
    w.total = w.off_mscores + up((size_t)D * g.S * K * g.S * 4);
    return w;
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

// LF[(s*K + k)][n'] = aff[s*seg + n'][k]  (zero beyond N)
__global__ void em_aff_to_filters_kernel(const float* __restrict__ aff, long long N, int K, int S, long long seg,
                                         float* __restrict__ LF) {
    const long long total = (long long)K * S * seg;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long np = i % seg;
        const long long m = i / seg;
        const int sidx = (int)(m / K), k = (int)(m - (long long)sidx * K);
        const long long n = sidx * seg + np;
        LF[i] = n < N ? aff[n * K + k] : 0.f;
    }
}

// numer[k][d] += sum_s scores[(d*S + s)][(s*K + k)]
__global__ void em_gather_numer_kernel(const float* __restrict__ scores, int D, int K, int S, float* __restrict__ numer) {
    const long long total = (long long)K * D;
    const int M = K * S;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / D), d = (int)(i - (long long)k * D);
        float acc = 0.f;
        for (int sidx = 0; sidx < S; ++sidx) acc += scores[((size_t)d * S + sidx) * M + (size_t)sidx * K + k];
        numer[i] += acc;
    }
}

// W = logT - logInvT (one thread per element); c[k] by one warp per component, fixed order.
__global__ void em_weights_kernel(const float* __restrict__ logT, const float* __restrict__ logInvT,
                                  float* __restrict__ W, float* __restrict__ c, int K, int D) {
    const long long n = (long long)K * D;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        W[i] = logT[i] - logInvT[i];
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp < K) {
        double acc = 0.0;
        for (int d = lane; d < D; d += 32) acc += (double)logInvT[(size_t)warp * D + d];
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) c[warp] = (float)acc;
    }
}

// One thread per row: ll_k = S_k + c_k + logw_k; max, sum-exp, posteriors.  Block-level reduction
// of the column sums and of the log-likelihood, one atomicAdd per block per value.
__global__ void __launch_bounds__(256) em_posterior_kernel(const float* __restrict__ S, const float* __restrict__ c,
                                                           const float* __restrict__ logw, long long N, int K,
                                                           float* __restrict__ aff, float* __restrict__ scores_out,
                                                           float* __restrict__ respsum, double* __restrict__ ll_out,
                                                           int hard) {
    __shared__ float s_resp[EM_MAX_K];
    __shared__ double s_ll;
    if (threadIdx.x < EM_MAX_K) s_resp[threadIdx.x] = 0.f;
    if (threadIdx.x == 0) s_ll = 0.0;
    __syncthreads();
    const long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    float a[EM_MAX_K];
    double my_ll = 0.0;
    if (n < N) {
        float mx = -INFINITY, smx = -INFINITY;
        int arg = 0;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            float s = S[n * K + k] + c[k];
            if (scores_out) scores_out[n * K + k] = s;
            if (s > smx) { smx = s; arg = k; }          // hard assignment: argmax of the template scores (:273-275)
            float v = s + logw[k];
            a[k] = v;
            mx = fmaxf(mx, v);
        }
        float sum = 0.f;
        for (int k = 0; k < K; ++k) { a[k] = expf(a[k] - mx); sum += a[k]; }
        my_ll = (double)logf(sum) + (double)mx;            // :269-270
        const float inv = 1.f / sum;
        for (int k = 0; k < K; ++k) {
            float p = hard ? (k == arg ? 1.f : 0.f) : a[k] * inv;
            aff[n * K + k] = p;
            a[k] = p;
        }
    }
    // block reduction: warp shuffles then shared atomics (<= 8 warps per block)
    for (int k = 0; k < K; ++k) {
        float v = n < N ? a[k] : 0.f;
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_resp[k], v);
    }
    for (int off = 16; off > 0; off >>= 1) my_ll += __shfl_xor_sync(0xffffffffu, my_ll, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_ll, my_ll);
    __syncthreads();
    if (threadIdx.x < K) atomicAdd(&respsum[threadIdx.x], s_resp[threadIdx.x]);
    if (threadIdx.x == 0) atomicAdd(ll_out, s_ll);
}

// M-step partials.  CTA = (group of 8 feature words = one 32-byte sector per row, row split);
// warp w owns word 8*gx + w, lane = feature bit, and walks the split's rows in order adding the
// row's K posteriors where its bit is set.  partial[split][k][d].
__global__ void __launch_bounds__(256) em_mstep_partial_kernel(const uint32_t* __restrict__ X, const float* __restrict__ aff,
                                                               long long N, int D, int rowwords, int K,
                                                               float* __restrict__ partial) {
    __shared__ float s_aff[128 * EM_MAX_K];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int word = blockIdx.x * 8 + warp;
    const int split = blockIdx.y;
    const long long rows_per = (N + gridDim.y - 1) / gridDim.y;
    const long long n0 = split * rows_per;
    long long n1 = n0 + rows_per; if (n1 > N) n1 = N;
    float acc[EM_MAX_K];
#pragma unroll
    for (int k = 0; k < EM_MAX_K; ++k) acc[k] = 0.f;
    for (long long nb = n0; nb < n1; nb += 128) {
        const int cnt = (int)((n1 - nb) < 128 ? (n1 - nb) : 128);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * K; i += blockDim.x) s_aff[i] = aff[nb * K + i];
        __syncthreads();
        if (word < rowwords) {
            for (int r = 0; r < cnt; ++r) {
                const uint32_t x = __ldg(X + (nb + r) * rowwords + word);
                if ((x >> lane) & 1u) {
#pragma unroll
                    for (int k = 0; k < EM_MAX_K; ++k)
                        if (k < K) acc[k] += s_aff[r * K + k];
                }
            }
        }
    }
    const int d = word * 32 + lane;
    if (word < rowwords && d < D) {
#pragma unroll
        for (int k = 0; k < EM_MAX_K; ++k)
            if (k < K) partial[((size_t)split * K + k) * D + d] = acc[k];
    }
}

__global__ void em_mstep_reduce_kernel(const float* __restrict__ partial, int nsplit, int K, int D, float* __restrict__ numer) {
    const long long n = (long long)K * D;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int sp = 0; sp < nsplit; ++sp) s += partial[(size_t)sp * n + i];
        numer[i] += s;
    }
}

__global__ void em_finalize_kernel(const float* __restrict__ stats, long long N_total, int D, int K, float min_prob,
                                   float* __restrict__ templates, float* __restrict__ logT, float* __restrict__ logInvT,
                                   float* __restrict__ weights, float* __restrict__ logw) {
    const long long n = (long long)K * D;
    const float* respsum = stats + n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / D);
        const float w = respsum[k] / (float)N_total;               // np.mean(affinities, 0)  (:178)
        float t = stats[i] / (float)N_total;                       // /= num_data             (:180)
        t = t / w;                                                 // /= weights              (:181)
        t = fminf(fmaxf(t, min_prob), 1.f - min_prob);             // threshold_templates     (:192-193)
        templates[i] = t;
        logT[i] = logf(t);
        logInvT[i] = logf(1.f - t);
        if (i - (long long)k * D == 0) { weights[k] = w; logw[k] = logf(w); }
    }
}

}  // namespace tsr

using namespace tsr;

extern "C" {

size_t tsr_em_workspace_bytes(long long N, int D, int K) {
    if (N < 0 || D < 1 || K < 1 || K > EM_MAX_K) return 0;
    return em_workspace(N, D, K).total;
}

int tsr_em_estep(const uint32_t* Xbits, long long N, int D, int rowwords, int K, const float* logT,
                 const float* logInvT, const float* logw, float* aff, float* scores_or_null, float* stats,
                 double* stats_ll, int hard, void* workspace, int impl, void* stream) {
    TSR_REQUIRE(Xbits && logT && logInvT && logw && aff && stats && stats_ll && workspace, "tsr_em_estep: null pointer");
    TSR_REQUIRE(K >= 1 && K <= EM_MAX_K, "tsr_em_estep: need 1 <= num_mix <= %d (got %d)", EM_MAX_K, K);
    TSR_REQUIRE(N >= 0 && N < (1ll << 31) && D >= 1 && rowwords == tsr_rowwords(D), "tsr_em_estep: bad sizes");
    if (N == 0) return TSR_OK;
    int rc = require_sm100("tsr_em_estep");
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    EmWorkspace w = em_workspace(N, D, K);
    uint8_t* ws = (uint8_t*)workspace;
    float* W = (float*)(ws + w.off_w);
    float* c = (float*)(ws + w.off_c);
    float* S = (float*)(ws + w.off_scores);
    em_weights_kernel<<<148 * 2, 256, 0, st>>>(logT, logInvT, W, c, K, D);
    rc = check_launch("tsr_em_estep(weights)");
    if (rc) return rc;
    if (impl == TSR_IMPL_TENSOR) {
        rc = tsr_prepare_filters(W, K, 1, D, EM_PLANES, ws + w.off_blob, stream);
        if (rc) return rc;
    }
    rc = tsr_detect(Xbits, 1, (int)N, D, rowwords, ws + w.off_blob, W, K, 1, EM_PLANES, S, impl, stream);
    if (rc) return rc;
    em_posterior_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(S, c, logw, N, K, aff, scores_or_null,
                                                                      stats + (size_t)K * D, stats_ll, hard);
    return check_launch("tsr_em_estep(posterior)");
}

int tsr_em_segments(long long N, int D, int K) { return (N < 1 || D < 1 || K < 1) ? 0 : em_geom(N, D, K).S; }
long long tsr_em_npad(long long N, int D, int K) { return (N < 1 || D < 1 || K < 1) ? 0 : em_geom(N, D, K).Npad; }

int tsr_transpose_bits(const uint32_t* Xbits, long long N, int D, int rowwords, uint32_t* XTbits, long long N_pad,
                       void* stream) {
    TSR_REQUIRE(Xbits && XTbits, "tsr_transpose_bits: null pointer");
    TSR_REQUIRE(N >= 1 && D >= 1 && rowwords == tsr_rowwords(D) && N_pad >= N && N_pad % 64 == 0, "tsr_transpose_bits: bad sizes");
    int rc = require_sm100("tsr_transpose_bits");
    if (rc) return rc;
    transpose_bits_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(Xbits, N, D, rowwords, XTbits, N_pad / 32);
    return check_launch("tsr_transpose_bits");
}

int tsr_em_mstep(const uint32_t* Xbits, const uint32_t* XTbits_or_null, long long N, int D, int rowwords, int K,
                 const float* aff, float* stats, void* workspace, int impl, void* stream) {
    TSR_REQUIRE(Xbits && aff && stats && workspace, "tsr_em_mstep: null pointer");
    TSR_REQUIRE(K >= 1 && K <= EM_MAX_K, "tsr_em_mstep: need 1 <= num_mix <= %d (got %d)", EM_MAX_K, K);
    TSR_REQUIRE(N >= 0 && D >= 1 && rowwords == tsr_rowwords(D), "tsr_em_mstep: bad sizes");
    TSR_REQUIRE(impl == TSR_IMPL_TENSOR || impl == TSR_IMPL_SIMT, "tsr_em_mstep: unknown impl %d", impl);
    if (N == 0) return TSR_OK;
    int rc = require_sm100("tsr_em_mstep");
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    EmWorkspace w = em_workspace(N, D, K);
    uint8_t* ws = (uint8_t*)workspace;
    if (impl == TSR_IMPL_SIMT) {
        float* partial = (float*)(ws + w.off_partial);
        int nsplit = (int)(N < EM_MSPLIT * 128 ? (N + 127) / 128 : EM_MSPLIT);
        dim3 grid((rowwords + 7) / 8, nsplit);
        em_mstep_partial_kernel<<<grid, 256, 0, st>>>(Xbits, aff, N, D, rowwords, K, partial);
        rc = check_launch("tsr_em_mstep(partial)");
        if (rc) return rc;
        em_mstep_reduce_kernel<<<148 * 2, 256, 0, st>>>(partial, nsplit, K, D, stats);
        return check_launch("tsr_em_mstep(reduce)");
    }
    TSR_REQUIRE(XTbits_or_null, "tsr_em_mstep: the tensor path needs the transposed bits (tsr_transpose_bits)");
    const EmGeom g = em_geom(N, D, K);
    TSR_REQUIRE(g.seg < (1ll << 31) && (long long)D * g.S < (1ll << 31), "tsr_em_mstep: problem too large");
    float* LF = (float*)(ws + w.off_mlf);
    float* scores = (float*)(ws + w.off_mscores);
    em_aff_to_filters_kernel<<<148 * 8, 256, 0, st>>>(aff, N, K, g.S, g.seg, LF);
    rc = check_launch("tsr_em_mstep(filters)");
    if (rc) return rc;
    rc = tsr_prepare_filters(LF, K * g.S, 1, (int)g.seg, TSR_PLANES_INT8X3, ws + w.off_mblob, stream);
    if (rc) return rc;
    rc = tsr_detect(XTbits_or_null, 1, D * g.S, (int)g.seg, tsr_rowwords((int)g.seg), ws + w.off_mblob, LF, K * g.S, 1,
                    TSR_PLANES_INT8X3, scores, TSR_IMPL_TENSOR, stream);
    if (rc) return rc;
    em_gather_numer_kernel<<<148 * 2, 256, 0, st>>>(scores, D, K, g.S, stats);
    return check_launch("tsr_em_mstep(gather)");
}

int tsr_em_finalize(const float* stats, long long N_total, int D, int K, float min_prob, float* templates,
                    float* logT, float* logInvT, float* weights, float* logw, void* stream) {
    TSR_REQUIRE(stats && templates && logT && logInvT && weights && logw, "tsr_em_finalize: null pointer");
    TSR_REQUIRE(K >= 1 && K <= EM_MAX_K && D >= 1 && N_total >= 1, "tsr_em_finalize: bad sizes");
    int rc = require_sm100("tsr_em_finalize");
    if (rc) return rc;
    em_finalize_kernel<<<148 * 2, 256, 0, (cudaStream_t)stream>>>(stats, N_total, D, K, min_prob, templates, logT,
                                                                 logInvT, weights, logw);
    return check_launch("tsr_em_finalize");
}

}  // extern "C"
