// patches.cu -- the two small neighbours of part coding:
//   * the magnitude / auxiliary Bernoulli terms of code_parts_mag_aux (code_parts.pyx:138-242),
//     added on top of the edge log-likelihoods the coding kernel produces;
//   * get_parts_mask (code_parts.pyx:297-348): masked edge counts per window, ordered stream
//     compaction of the windows above threshold, gather of the 5x5x8 patches (bytes for the
//     reference signature, packed bit rows for the EM kernels).
// Integer / byte work and a few adds per location; nothing here is near a roofline.
#include "common.cuh"

namespace tsr {

// ------------------------------------------------------------------------------ mag / aux terms
// One thread per (location, part).  For window origin (i0, j0), code_parts.pyx:195-239:
//   sum_{i<ph} [ sum_{j<aux_dim} (X_aux[i0+i][j] ? lap : laip)[k][i*aux_dim+j]
//              + (any_{j<pw} X_mags[i0+i][((j0+j)*mag_dim)/F] ? lmp : lmip)[k][i]
//              + sum_{q<mag_dim} ((c-q)^2 > 1 ? lmip : lmp)[k][q+ph] ],   c = ((j0+pw/2)*mag_dim)/F
// gated by thresh_mask[i0+ph/2][j0+pw/2]; ungated locations get score 0 (the map stays all zero).
__global__ void __launch_bounds__(256) mag_aux_terms_kernel(float* __restrict__ scores, uint8_t* __restrict__ gate,
                                                            const uint8_t* __restrict__ X_mags,
                                                            const uint8_t* __restrict__ X_aux,
                                                            const uint8_t* __restrict__ tmask,
                                                            const double* __restrict__ lmp, const double* __restrict__ lmip,
                                                            const double* __restrict__ lap, const double* __restrict__ laip,
                                                            long long n, int T, int F, int K, int ph, int pw, int mag_dim,
                                                            int aux_dim) {
    const int To = T - ph + 1, Fo = F - pw + 1;
    const int magw = ph + mag_dim, auxw = ph * aux_dim;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx % K);
        long long loc = idx / K;
        const int j0 = (int)(loc % Fo);
        loc /= Fo;
        const int i0 = (int)(loc % To);
        const long long b = loc / To;
        const bool on = tmask[(b * T + i0 + ph / 2) * F + j0 + pw / 2] != 0;
        if (k == 0 && gate) gate[(b * To + i0) * Fo + j0] = on ? 1 : 0;
        if (!on) { scores[idx] = 0.f; continue; }
        double s = 0.0;
        const int c = ((j0 + pw / 2) * mag_dim) / F;
        for (int i = 0; i < ph; ++i) {
            const uint8_t* arow = X_aux + (b * T + i0 + i) * aux_dim;
            for (int j = 0; j < aux_dim; ++j) s += (arow[j] ? lap : laip)[(size_t)k * auxw + i * aux_dim + j];
            const uint8_t* mrow = X_mags + (b * T + i0 + i) * mag_dim;
            bool any = false;
            for (int j = 0; j < pw; ++j) any = any || mrow[((j0 + j) * mag_dim) / F] != 0;
            for (int q = 0; q < mag_dim; ++q) s += ((c - q) * (c - q) > 1 ? lmip : lmp)[(size_t)k * magw + q + ph];
            s += (any ? lmp : lmip)[(size_t)k * magw + i];
        }
        scores[idx] = (float)((double)scores[idx] + s);
    }
}

// ------------------------------------------------------------------------------ masked counts
// counts[b][i][j] = sum_{a<ph, c<pw} mask[a][c] * popcount(bits[b][i+a][j+c])   (window-relative mask)
__global__ void __launch_bounds__(256) patch_counts_kernel(const uint8_t* __restrict__ bits, const uint8_t* __restrict__ mask,
                                                           long long n, int T, int F, int ph, int pw,
                                                           uint32_t* __restrict__ counts) {
    const int To = T - ph + 1, Fo = F - pw + 1;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % Fo);
        long long r = idx / Fo;
        const int i = (int)(r % To);
        const long long b = r / To;
        uint32_t cnt = 0;
        for (int a = 0; a < ph; ++a) {
            const uint8_t* row = bits + (b * T + i + a) * F + j;
            for (int c = 0; c < pw; ++c)
                if (!mask || mask[a * pw + c]) cnt += __popc((uint32_t)row[c]);
        }
        counts[idx] = cnt;
    }
}

// ------------------------------------------------------------------------------ ordered compaction
// slot[idx] = number of selected locations before idx (row-major order, as the reference's
// patch_id counter), or -1.  Three passes over 1024-location chunks: chunk totals, scan of the
// totals (one CTA), slots.
#define PS_CHUNK 1024
__device__ __forceinline__ bool ps_selected(uint32_t c, uint32_t lo, uint32_t hi) { return c >= lo && c < hi; }

__global__ void __launch_bounds__(256) patch_chunk_totals_kernel(const uint32_t* __restrict__ counts, long long n, uint32_t lo,
                                                                 uint32_t hi, uint32_t* __restrict__ totals) {
    __shared__ uint32_t wsum[8];
    const long long base = (long long)blockIdx.x * PS_CHUNK;
    uint32_t c = 0;
    for (int q = threadIdx.x; q < PS_CHUNK; q += 256) {
        const long long idx = base + q;
        if (idx < n && ps_selected(counts[idx], lo, hi)) ++c;
    }
    for (int off = 16; off; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; ++w) t += wsum[w];
        totals[blockIdx.x] = t;
    }
}

// exclusive scan of totals[0..nchunks) in place; totals[nchunks] = grand total (single CTA)
__global__ void __launch_bounds__(1024) patch_scan_totals_kernel(uint32_t* __restrict__ totals, int nchunks) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nchunks; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < nchunks ? totals[i] : 0u;
        uint32_t x = v;
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
            if ((threadIdx.x & 31) >= off) x += y;
        }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t s = wsum[threadIdx.x];
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, s, off);
                if (threadIdx.x >= off) s += y;
            }
            wsum[threadIdx.x] = s;
        }
        __syncthreads();
        const uint32_t before = carry + (threadIdx.x >= 32 ? wsum[(threadIdx.x >> 5) - 1] : 0u) + x - v;
        if (i < nchunks) totals[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[nchunks] = carry;
}

__global__ void __launch_bounds__(256) patch_slots_kernel(const uint32_t* __restrict__ counts, long long n, uint32_t lo,
                                                          uint32_t hi, const uint32_t* __restrict__ totals,
                                                          int32_t* __restrict__ slot) {
    __shared__ uint32_t wsum[8];
    const long long base = (long long)blockIdx.x * PS_CHUNK;
    uint32_t run = totals[blockIdx.x];
    for (int q0 = 0; q0 < PS_CHUNK; q0 += 256) {       // consecutive threads = consecutive locations
        const long long idx = base + q0 + threadIdx.x;
        const bool sel = idx < n && ps_selected(counts[idx], lo, hi);
        const uint32_t ball = __ballot_sync(0xffffffffu, sel);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) wsum[w] = __popc(ball);
        __syncthreads();
        uint32_t before = run;
        for (int x = 0; x < w; ++x) before += wsum[x];
        uint32_t all = 0;
        for (int x = 0; x < 8; ++x) all += wsum[x];
        if (idx < n) slot[idx] = sel ? (int32_t)(before + __popc(ball & ((1u << lane) - 1u))) : -1;
        run += all;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------ gather
// One warp per selected location.  patch_set [n][ph][pw][E] bytes (0/1), S_patch [n][ph][pw],
// locs [n][3] = (b, i, j), patchbits [n][rowwords] with feature (a*pw + c)*E + z.
__global__ void __launch_bounds__(256) patch_gather_kernel(const uint8_t* __restrict__ bits, const double* __restrict__ S,
                                                           const int32_t* __restrict__ slot, long long n, int T, int F, int E,
                                                           int ph, int pw, int rowwords, uint8_t* __restrict__ patch_set,
                                                           double* __restrict__ S_patch, uint32_t* __restrict__ locs,
                                                           uint32_t* __restrict__ patchbits) {
    const int To = T - ph + 1, Fo = F - pw + 1;
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    const int np = ph * pw, D = np * E;
    for (long long idx = warp0; idx < n; idx += nwarp) {
        const int32_t s = slot[idx];
        if (s < 0) continue;
        const int j = (int)(idx % Fo);
        long long r = idx / Fo;
        const int i = (int)(r % To);
        const long long b = r / To;
        if (locs && lane == 0) { locs[3 * (size_t)s] = (uint32_t)b; locs[3 * (size_t)s + 1] = i; locs[3 * (size_t)s + 2] = j; }
        for (int p = lane; p < np; p += 32) {
            const int a = p / pw, c = p - a * pw;
            const size_t src = (size_t)(b * T + i + a) * F + j + c;
            const uint32_t v = bits[src];
            if (patch_set)
                for (int z = 0; z < E; ++z) patch_set[((size_t)s * np + p) * E + z] = (v >> z) & 1u;
            if (S_patch) S_patch[(size_t)s * np + p] = S[src];
        }
        if (patchbits)
            for (int w = lane; w < rowwords; w += 32) {
                uint32_t word = 0;
                for (int q = 0; q < 32; ++q) {
                    const int d = 32 * w + q;
                    if (d >= D) break;
                    const int p = d / E, z = d - p * E;
                    const int a = p / pw, c = p - a * pw;
                    word |= ((bits[(size_t)(b * T + i + a) * F + j + c] >> z) & 1u) << q;
                }
                patchbits[(size_t)s * rowwords + w] = word;
            }
    }
}

// ------------------------------------------------------------------------------ peak picking (N2)
// flags[b][t] = 1 iff 1 <= t <= l-2 and v[t] >= v[t-1], v[t] >= v[t+1], v[t] >= threshold, with
// l = length[b] - C0 + 1 (roc_functions.get_max_above_threshold, roc_functions.py:177-186).
__global__ void __launch_bounds__(256) detect_peaks_kernel(const float* __restrict__ v, const int* __restrict__ lengths,
                                                           long long n, int T, int C0, float threshold,
                                                           uint32_t* __restrict__ flags) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / T;
        const int t = (int)(idx - b * T);
        int l = (lengths ? lengths[b] : T) - C0 + 1;
        if (l > T) l = T;
        uint32_t f = 0;
        if (t >= 1 && t <= l - 2) {
            const float x = v[idx];
            f = (x >= v[idx - 1] && x >= v[idx + 1] && x >= threshold) ? 1u : 0u;
        }
        flags[idx] = f;
    }
}
__global__ void __launch_bounds__(256) peaks_gather_kernel(const float* __restrict__ v, const int32_t* __restrict__ slot,
                                                           long long n, int T, int32_t* __restrict__ peak_bt,
                                                           float* __restrict__ peak_val) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int32_t s = slot[idx];
        if (s < 0) continue;
        const long long b = idx / T;
        peak_bt[2 * (size_t)s] = (int32_t)b;
        peak_bt[2 * (size_t)s + 1] = (int32_t)(idx - b * T);
        if (peak_val) peak_val[s] = v[idx];
    }
}

// ------------------------------------------------------------------------------ N2: clustering of the peaks
// One thread per utterance walks its (few) peaks: cluster_times.pyx:133-219 followed by
// roc_functions.get_cluster_starts_ends (:187-212).  peak_bt rows are (b, t) sorted by b then t.  Peaks at least C0
// frames apart start a new cluster; a cluster whose extent exceeds C1 is cut at its widest gap (codes: 2 = cluster
// start, final; 1 = start that the loop may still split; 0 = continuation).  Distances wrap at 16 bits like the
// reference's uint16 arithmetic (frames are < 65536 on this path).
__global__ void cluster_peaks_kernel(const int32_t* __restrict__ peak_bt, int n, int B, float C0, float C1,
                                     int32_t* __restrict__ offsets, uint8_t* __restrict__ partition,
                                     int32_t* __restrict__ clusters, int32_t* __restrict__ ncluster) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > B) return;
    auto lower = [&](int key) {                       // first peak whose utterance index is >= key
        int lo = 0, hi = n;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (peak_bt[2 * mid] < key) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    const int first = lower(b);
    offsets[b] = first;
    if (b == B) return;
    const int m = lower(b + 1) - first;               // peaks of this utterance
    const int32_t* t = peak_bt + 2 * first + 1;       // t[2*i] = frame of peak i
    uint8_t* part = partition + first + b;            // m + 1 codes
    for (int i = 0; i <= m; ++i) part[i] = 0;
    part[0] = 1;
    part[m] = 2;
    for (int i = 1; i < m; ++i)
        if ((float)(t[2 * i] - t[2 * i - 2]) >= C0) part[i] = 1;
    int start = 0, widest = 0;
    while (start < m) {
        if (part[start + 1] > 0) {
            part[start] = 2;
            ++start;
            while (part[start] != 1 && start < m) ++start;
            continue;
        }
        unsigned int max_gap = 0, extent = 0;
        int k = 1;
        while (part[start + k] == 0) {
            const unsigned int gap = (unsigned int)(t[2 * (start + k)] - t[2 * (start + k - 1)]) & 0xFFFFu;
            if (gap > max_gap) { widest = start + k; max_gap = gap; }
            extent = (extent + gap) & 0xFFFFu;
            ++k;
        }
        if ((float)extent > C1) {
            part[widest] = ((float)(t[2 * (start + k - 1)] - t[2 * widest]) <= C1) ? 2 : 1;
            if ((float)(t[2 * (widest - 1)] - t[2 * start]) <= C1) {
                part[start] = 2;
                if (part[widest] == 1) start = widest;
            } else {
                part[start] = 1;
            }
        } else {
            part[start] = 2;
            start += k;
            while (part[start] != 1 && start < m) ++start;
        }
    }
    // (start, end) = (first peak, last peak + 1) of every cluster
    int nc = 0, i = 0;
    while (i < m) {
        const int s0 = t[2 * i];
        ++i;
        while (part[i] != 2) ++i;
        clusters[2 * (first + nc)] = s0;
        clusters[2 * (first + nc) + 1] = t[2 * (i - 1)] + 1;
        ++nc;
    }
    ncluster[b] = nc;
}

}  // namespace tsr

using namespace tsr;

extern "C" {

int tsr_code_parts_mag_aux_terms(float* scores, uint8_t* gate_or_null, const uint8_t* X_mags, const uint8_t* X_aux,
                                 const uint8_t* thresh_mask, const double* log_mag_parts, const double* log_mag_invparts,
                                 const double* log_aux_parts, const double* log_aux_invparts, int B, int T, int F, int K,
                                 int ph, int pw, int mag_dim, int aux_dim, void* stream) {
    if (B == 0) return TSR_OK;
    TSR_REQUIRE(scores && thresh_mask && log_mag_parts && log_mag_invparts, "tsr_code_parts_mag_aux_terms: null pointer");
    TSR_REQUIRE(mag_dim == 0 || X_mags, "tsr_code_parts_mag_aux_terms: null X_mags");
    TSR_REQUIRE(aux_dim == 0 || (X_aux && log_aux_parts && log_aux_invparts), "tsr_code_parts_mag_aux_terms: null auxiliary input");
    TSR_REQUIRE(B > 0 && K >= 1 && ph >= 1 && pw >= 1 && T >= ph && F >= pw && mag_dim >= 0 && aux_dim >= 0,
                "tsr_code_parts_mag_aux_terms: bad sizes");
    int rc = require_sm100("tsr_code_parts_mag_aux_terms");
    if (rc) return rc;
    const long long n = (long long)B * (T - ph + 1) * (F - pw + 1) * K;
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)sm_count() * 32) blocks = (long long)sm_count() * 32;
    mag_aux_terms_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(scores, gate_or_null, X_mags, X_aux, thresh_mask,
                                                                          log_mag_parts, log_mag_invparts, log_aux_parts,
                                                                          log_aux_invparts, n, T, F, K, ph, pw, mag_dim, aux_dim);
    return check_launch("tsr_code_parts_mag_aux_terms");
}

int tsr_patch_counts(const uint8_t* bits, const uint8_t* mask_or_null, int B, int T, int F, int ph, int pw, uint32_t* counts,
                     void* stream) {
    if (B == 0) return TSR_OK;
    TSR_REQUIRE(bits && counts, "tsr_patch_counts: null pointer");
    TSR_REQUIRE(B > 0 && ph >= 1 && pw >= 1 && T >= ph && F >= pw, "tsr_patch_counts: bad sizes");
    int rc = require_sm100("tsr_patch_counts");
    if (rc) return rc;
    const long long n = (long long)B * (T - ph + 1) * (F - pw + 1);
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)sm_count() * 32) blocks = (long long)sm_count() * 32;
    patch_counts_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(bits, mask_or_null, n, T, F, ph, pw, counts);
    return check_launch("tsr_patch_counts");
}

size_t tsr_patch_select_workspace_bytes(long long n_loc) {
    if (n_loc < 0) return 0;
    return (size_t)((n_loc + PS_CHUNK - 1) / PS_CHUNK + 1) * sizeof(uint32_t);
}

int tsr_patch_select(const uint32_t* counts, long long n_loc, unsigned int lo, unsigned int hi, int32_t* slot, uint32_t* total,
                     void* workspace, void* stream) {
    TSR_REQUIRE(total, "tsr_patch_select: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_loc == 0) {
        TSR_CUDA(cudaMemsetAsync(total, 0, sizeof(uint32_t), st));
        return TSR_OK;
    }
    TSR_REQUIRE(counts && slot && workspace, "tsr_patch_select: null pointer");
    TSR_REQUIRE(n_loc > 0 && n_loc < (1LL << 31), "tsr_patch_select: n_loc out of range");
    int rc = require_sm100("tsr_patch_select");
    if (rc) return rc;
    const int nchunks = (int)((n_loc + PS_CHUNK - 1) / PS_CHUNK);
    uint32_t* totals = (uint32_t*)workspace;
    patch_chunk_totals_kernel<<<nchunks, 256, 0, st>>>(counts, n_loc, lo, hi, totals);
    patch_scan_totals_kernel<<<1, 1024, 0, st>>>(totals, nchunks);
    patch_slots_kernel<<<nchunks, 256, 0, st>>>(counts, n_loc, lo, hi, totals, slot);
    TSR_CUDA(cudaMemcpyAsync(total, totals + nchunks, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    return check_launch("tsr_patch_select");
}

int tsr_patch_gather(const uint8_t* bits, const double* S_or_null, const int32_t* slot, int B, int T, int F, int E, int ph,
                     int pw, uint8_t* patch_set_or_null, double* S_patch_or_null, uint32_t* locs_or_null,
                     uint32_t* patchbits_or_null, void* stream) {
    if (B == 0) return TSR_OK;
    TSR_REQUIRE(bits && slot, "tsr_patch_gather: null pointer");
    TSR_REQUIRE(B > 0 && ph >= 1 && pw >= 1 && T >= ph && F >= pw && E >= 1 && E <= 8, "tsr_patch_gather: bad sizes");
    TSR_REQUIRE(!S_patch_or_null || S_or_null, "tsr_patch_gather: S_patch requested without S");
    int rc = require_sm100("tsr_patch_gather");
    if (rc) return rc;
    const long long n = (long long)B * (T - ph + 1) * (F - pw + 1);
    long long blocks = (n + 7) / 8;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    patch_gather_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(bits, S_or_null, slot, n, T, F, E, ph, pw,
                                                                         tsr_rowwords(ph * pw * E), patch_set_or_null,
                                                                         S_patch_or_null, locs_or_null, patchbits_or_null);
    return check_launch("tsr_patch_gather");
}

int tsr_detect_peaks(const float* track, const int* lengths_or_null, int B, int T, int C0, float threshold, uint32_t* flags,
                     void* stream) {
    if (B == 0 || T == 0) return TSR_OK;
    TSR_REQUIRE(track && flags, "tsr_detect_peaks: null pointer");
    TSR_REQUIRE(B > 0 && T > 0, "tsr_detect_peaks: bad sizes");
    int rc = require_sm100("tsr_detect_peaks");
    if (rc) return rc;
    const long long n = (long long)B * T;
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)sm_count() * 32) blocks = (long long)sm_count() * 32;
    detect_peaks_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(track, lengths_or_null, n, T, C0, threshold, flags);
    return check_launch("tsr_detect_peaks");
}

int tsr_peaks_gather(const float* track, const int32_t* slot, int B, int T, int32_t* peak_bt, float* peak_val_or_null,
                     void* stream) {
    if (B == 0 || T == 0) return TSR_OK;
    TSR_REQUIRE(track && slot && peak_bt, "tsr_peaks_gather: null pointer");
    int rc = require_sm100("tsr_peaks_gather");
    if (rc) return rc;
    const long long n = (long long)B * T;
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)sm_count() * 32) blocks = (long long)sm_count() * 32;
    peaks_gather_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(track, slot, n, T, peak_bt, peak_val_or_null);
    return check_launch("tsr_peaks_gather");
}

int tsr_cluster_peaks(const int32_t* peak_bt, int n, int B, float C0, float C1, int32_t* offsets, uint8_t* partition,
                      int32_t* clusters, int32_t* ncluster, void* stream) {
    TSR_REQUIRE(B >= 0 && n >= 0, "tsr_cluster_peaks: negative size");
    TSR_REQUIRE(offsets && (B == 0 || ncluster), "tsr_cluster_peaks: null pointer");
    TSR_REQUIRE(n == 0 || (peak_bt && clusters), "tsr_cluster_peaks: null pointer");
    TSR_REQUIRE(partition || (n == 0 && B == 0), "tsr_cluster_peaks: null pointer");
    int rc = require_sm100("tsr_cluster_peaks");
    if (rc) return rc;
    cluster_peaks_kernel<<<(B + 1 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(peak_bt, n, B, C0, C1, offsets, partition,
                                                                               clusters, ncluster);
    return check_launch("tsr_cluster_peaks");
}

}  // extern "C"
