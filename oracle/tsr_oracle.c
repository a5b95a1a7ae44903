/*
 * tsr_oracle.c -- CPU restatement of the template_speech_rec scoring hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package
 * (template-speech-recognition_b200/) may link, import or call this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / the timed CPU arm.
 *
 * Every function restates, loop for loop and in the same accumulation order and
 * arithmetic type, a function of the reference (paths relative to /root/reference).
 * Pinning: tests/test_oracle.py checks this file against the reference Cython compiled
 * unmodified into oracle/_ref (when present) and against tests/golden/ fixtures that
 * were generated from that compiled reference (tests/golden/make_golden.py).
 *
 * All arrays are C-contiguous.  No allocation happens here; callers pass outputs.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: the reference is built with
 * gcc -O2 and no FMA contraction, code_parts.pyx:2-3).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* ---- code_parts.pyx:18-31  _count_edges ------------------------------------------ */
static unsigned int count_edges(const uint8_t *X, int F, int E, int i0, int i1, int j0, int j1)
{
    unsigned int c = 0;
    for (int i = i0; i < i1; ++i)
        for (int j = j0; j < j1; ++j)
            for (int z = 0; z < E; ++z)
                if (X[((size_t)i * F + j) * E + z]) c++;
    return c;
}

/* ---- code_parts.pyx:33-48  _count_edges_mask --------------------------------------
 * NOTE: the reference indexes the mask with the ABSOLUTE window coordinates, M[i,j]
 * with i in [i_start,i_end): with boundscheck off this reads out of bounds of the
 * (ph,pw) mask for every window but the first.  That is undefined behaviour in the
 * reference; the intended semantics (mask relative to the window, as the callers in
 * Notebook/12/reestimate_parts.py:122 assume) are restated here and the divergence is
 * listed in DESIGN.md.  For i_start=j_start=0 both agree, and tests pin that case. */
static unsigned int count_edges_mask(const uint8_t *X, const uint8_t *M, int F, int E,
                                     int i0, int j0, int ph, int pw)
{
    unsigned int c = 0;
    for (int i = 0; i < ph; ++i)
        for (int j = 0; j < pw; ++j)
            if (M[i * pw + j])
                for (int z = 0; z < E; ++z)
                    if (X[((size_t)(i0 + i) * F + (j0 + j)) * E + z]) c++;
    return c;
}

/* accumulate the K part log-likelihoods of one window: code_parts.pyx:123-131 */
static void accumulate_window(const uint8_t *X, int F, int E, const double *lp, const double *lip,
                              int K, int ph, int pw, int i0, int j0, double *out /* K+1 */)
{
    for (int i = 0; i < ph; ++i)
        for (int j = 0; j < pw; ++j)
            for (int z = 0; z < E; ++z) {
                const size_t w = ((size_t)i * pw + j) * E + z;
                const size_t stride = (size_t)ph * pw * E;
                if (X[((size_t)(i0 + i) * F + (j0 + j)) * E + z]) {
                    for (int k = 0; k < K; ++k) out[k + 1] += lp[k * stride + w];
                } else {
                    for (int k = 0; k < K; ++k) out[k + 1] += lip[k * stride + w];
                }
            }
}

/* ---- code_parts.pyx:70-135  code_parts --------------------------------------------
 * out: (T-ph+1, F-pw+1, K+1) float64, zero-initialised here. */
void oracle_code_parts(const uint8_t *X, int T, int F, int E, const double *lp, const double *lip,
                       int K, int ph, int pw, int threshold, double *out)
{
    const int To = T - ph + 1, Fo = F - pw + 1;
    memset(out, 0, sizeof(double) * (size_t)To * Fo * (K + 1));
    for (int i0 = 0; i0 < To; ++i0)
        for (int j0 = 0; j0 < Fo; ++j0) {
            double *o = out + ((size_t)i0 * Fo + j0) * (K + 1);
            unsigned int count = count_edges(X, F, E, i0, i0 + ph, j0, j0 + pw);
            if (count >= (unsigned int)threshold) {
                o[0] = -INFINITY;
                accumulate_window(X, F, E, lp, lip, K, ph, pw, i0, j0, o);
            } else {
                o[0] = 0.0;
            }
        }
}

/* ---- code_parts.pyx:247-295  code_parts_fast -------------------------------------- */
void oracle_code_parts_fast(const uint8_t *X, int T, int F, int E, const double *lp, const double *lip,
                            int K, int ph, int pw, int threshold, double *out)
{
    const int To = T - ph + 1, Fo = F - pw + 1;
    memset(out, 0, sizeof(double) * (size_t)To * Fo * (K + 1));
    for (int i0 = 0; i0 < To; ++i0)
        for (int j0 = 0; j0 < Fo; ++j0) {
            double *o = out + ((size_t)i0 * Fo + j0) * (K + 1);
            unsigned int count = count_edges(X, F, E, i0, i0 + ph, j0, j0 + pw);
            if (count >= (unsigned int)threshold) {
                accumulate_window(X, F, E, lp, lip, K, ph, pw, i0, j0, o);
                for (int k = 0; k < K; ++k) o[0] += o[k + 1] - 1;   /* :290-291 */
            } else {
                o[0] = 1.;                                              /* :293 */
            }
        }
}

/* ---- code_parts.pyx:353-404  code_parts_mask -------------------------------------- */
void oracle_code_parts_mask(const uint8_t *X, int T, int F, int E, const uint8_t *M,
                            const double *lp, const double *lip, int K, int ph, int pw,
                            int mask_threshold, int absolute_threshold, double *out)
{
    const int To = T - ph + 1, Fo = F - pw + 1;
    memset(out, 0, sizeof(double) * (size_t)To * Fo * (K + 1));
    for (int i0 = 0; i0 < To; ++i0)
        for (int j0 = 0; j0 < Fo; ++j0) {
            double *o = out + ((size_t)i0 * Fo + j0) * (K + 1);
            unsigned int count = count_edges(X, F, E, i0, i0 + ph, j0, j0 + pw);
            unsigned int mcount = count_edges_mask(X, M, F, E, i0, j0, ph, pw);
            if (count >= (unsigned int)absolute_threshold && mcount > (unsigned int)mask_threshold) {
                accumulate_window(X, F, E, lp, lip, K, ph, pw, i0, j0, o);
                for (int k = 0; k < K; ++k) o[0] += o[k + 1] - 1;
            } else {
                o[0] = 1.;
            }
        }
}

/* ---- code_parts.pyx:138-242  code_parts_mag_aux -------------------------------------
 * Same loop nest and float64 accumulation order as the reference: per window row i the
 * auxiliary terms (:199-206), then per column j the band-magnitude flag (:209-210), the edge
 * terms (:211-217) and -- at the centre column only -- the band profile (:218-228), then the
 * row's magnitude term (:232-237).  Gate: X_thresh_mask at the window centre (:191-192).
 * All index arithmetic is unsigned-int C division, as in the Cython. */
void oracle_code_parts_mag_aux(const uint8_t *X, int T, int F, int E, const uint8_t *X_mags, int mag_dim,
                               const uint8_t *X_aux, int aux_dim, const double *lp, const double *lip,
                               const double *lmp, const double *lmip, const double *lap, const double *laip,
                               const uint8_t *tmask, int K, int ph, int pw, double *out)
{
    const int To = T - ph + 1, Fo = F - pw + 1;
    const size_t stride = (size_t)ph * pw * E;
    const int magw = ph + mag_dim, auxw = ph * aux_dim;
    memset(out, 0, sizeof(double) * (size_t)To * Fo * (K + 1));
    for (unsigned int i0 = 0; i0 < (unsigned int)To; ++i0)
        for (unsigned int j0 = 0; j0 < (unsigned int)Fo; ++j0) {
            double *o = out + ((size_t)i0 * Fo + j0) * (K + 1);
            if (!tmask[(size_t)(i0 + ph / 2) * F + (j0 + pw / 2)]) { o[0] = 0.0; continue; }
            o[0] = -INFINITY;
            for (unsigned int i = 0; i < (unsigned int)ph; ++i) {
                unsigned int tmp_freq_mag = 0;
                for (unsigned int j = 0; j < (unsigned int)aux_dim; ++j) {
                    const double *tab = X_aux[(size_t)(i0 + i) * aux_dim + j] ? lap : laip;
                    for (int k = 0; k < K; ++k) o[k + 1] += tab[(size_t)k * auxw + i * aux_dim + j];
                }
                for (unsigned int j = 0; j < (unsigned int)pw; ++j) {
                    if (X_mags[(size_t)(i0 + i) * mag_dim + (((j0 + j) * (unsigned int)mag_dim) / (unsigned int)F)]) tmp_freq_mag = 1;
                    for (int z = 0; z < E; ++z) {
                        const size_t w = ((size_t)i * pw + j) * E + z;
                        const double *tab = X[((size_t)(i0 + i) * F + (j0 + j)) * E + z] ? lp : lip;
                        for (int k = 0; k < K; ++k) o[k + 1] += tab[k * stride + w];
                    }
                    if (j == (unsigned int)pw / 2)
                        for (unsigned int q = 0; q < (unsigned int)mag_dim; ++q) {
                            const unsigned int c = ((j0 + j) * (unsigned int)mag_dim) / (unsigned int)F;
                            const double *tab = ((c - q) * (c - q) > 1) ? lmip : lmp;
                            for (int k = 0; k < K; ++k) o[k + 1] += tab[(size_t)k * magw + q + ph];
                        }
                }
                {
                    const double *tab = tmp_freq_mag ? lmp : lmip;
                    for (int k = 0; k < K; ++k) o[k + 1] += tab[(size_t)k * magw + i];
                }
            }
        }
}

/* ---- code_parts.pyx:406-457 spread_patches == spread_waliji_patches.pyx:15-59 -------
 * X int64 (D0,D1) codes (0 = background); out uint8 (D0,D1,K), zeroed here. */
void oracle_spread_patches(const int64_t *X, int D0, int D1, int r0, int r1, int K, uint8_t *out)
{
    memset(out, 0, (size_t)D0 * D1 * K);
    for (int i = 0; i < D0; ++i) {
        int lo0 = i - r0 > 0 ? i - r0 : 0;
        int hi0 = i + r0 + 1 < D0 ? i + r0 + 1 : D0;
        for (int j = 0; j < D1; ++j) {
            int lo1 = j - r1 > 0 ? j - r1 : 0;
            int hi1 = j + r1 + 1 < D1 ? j + r1 + 1 : D1;
            for (int x0 = lo0; x0 < hi0; ++x0)
                for (int x1 = lo1; x1 < hi1; ++x1) {
                    int64_t p = X[(size_t)x0 * D1 + x1];
                    if (p > 0) out[((size_t)i * D1 + j) * K + (p - 1)] = 1;
                }
        }
    }
}

/* ---- code_parts.pyx:459-503  spread_graph_patches (in place, sequential in p1) ----- */
void oracle_spread_graph_patches(uint8_t *X, int D0, int D1, int K, const uint8_t *graph)
{
    for (int i = 0; i < D0; ++i)
        for (int j = 0; j < D1; ++j) {
            uint8_t *x = X + ((size_t)i * D1 + j) * K;
            for (int p1 = 0; p1 < K; ++p1)
                if (x[p1] > 0)
                    for (int p2 = 0; p2 < K; ++p2)
                        if (graph[p1 * K + p2]) x[p2] = 1;
        }
}

/* ---- code_parts.pyx:297-348  get_parts_mask ----------------------------------------
 * Two-pass like the reference: pass 1 (patches==NULL) fills counts and returns n;
 * pass 2 fills patches (n,ph,pw,E) u8, S_patches (n,ph,pw) f64, locs (n,2) u32.
 * Mask is window-relative (see count_edges_mask note). */
int oracle_get_parts_mask(const uint8_t *X, int T, int F, int E, const uint8_t *M, int ph, int pw,
                          const double *S, int threshold, uint32_t *counts,
                          uint8_t *patches, double *S_patches, uint32_t *locs)
{
    const int To = T - ph + 1, Fo = F - pw + 1;
    int n = 0;
    for (int i = 0; i < To; ++i)
        for (int j = 0; j < Fo; ++j) {
            unsigned int c = count_edges_mask(X, M, F, E, i, j, ph, pw);
            counts[i * Fo + j] = c;
            if (c >= (unsigned int)threshold) {
                if (patches) {
                    locs[2 * n] = i; locs[2 * n + 1] = j;
                    for (int a = 0; a < ph; ++a)
                        for (int b = 0; b < pw; ++b) {
                            S_patches[((size_t)n * ph + a) * pw + b] = S[(size_t)(i + a) * F + (j + b)];
                            for (int z = 0; z < E; ++z)
                                patches[(((size_t)n * ph + a) * pw + b) * E + z] =
                                    X[((size_t)(i + a) * F + (j + b)) * E + z];
                        }
                }
                n++;
            }
        }
    return n;
}

/* ---- compute_likelihood_linear_filter.pyx:135-176  detect --------------------------
 * F uint8 (T,D1,D2), LF float32 (L,D1,D2); out float32 (T,), tail L-1 entries stay 0.
 * float32 sequential accumulation, zeros skipped. */
void oracle_detect(const uint8_t *Fm, int T, int D1, int D2, const float *LF, int L, float *out)
{
    const size_t row = (size_t)D1 * D2;
    memset(out, 0, sizeof(float) * (size_t)T);
    for (int t = 0; t < T - L + 1; ++t) {
        float acc = 0.f;
        for (int tau = 0; tau < L; ++tau) {
            const uint8_t *f = Fm + (size_t)(t + tau) * row;
            const float *w = LF + (size_t)tau * row;
            for (size_t d = 0; d < row; ++d)
                if (f[d]) acc += w[d];
        }
        out[t] = acc;
    }
}

/* ---- compute_likelihood_linear_filter.pyx:178-219 detect_float (and the 2-D variants
 *      detect_float2d :281-318, detect_spectral :322-359 with D2 = 1) ----------------- */
void oracle_detect_float(const float *Fm, int T, int D1, int D2, const float *LF, int L, float *out)
{
    const size_t row = (size_t)D1 * D2;
    memset(out, 0, sizeof(float) * (size_t)T);
    for (int t = 0; t < T - L + 1; ++t) {
        float acc = 0.f;
        for (int tau = 0; tau < L; ++tau) {
            const float *f = Fm + (size_t)(t + tau) * row;
            const float *w = LF + (size_t)tau * row;
            for (size_t d = 0; d < row; ++d) {
                volatile float prod = w[d] * f[d];   /* separate multiply, then add */
                acc += prod;
            }
        }
        out[t] = acc;
    }
}

/* ---- compute_likelihood_linear_filter.pyx:221-278 detect_float_max ------------------ */
void oracle_detect_float_max(const float *Fm, int T, int D1, int D2, const float *LF, int M, int L,
                             const float *cs, float *out, uint16_t *arg)
{
    const size_t row = (size_t)D1 * D2;
    memset(out, 0, sizeof(float) * (size_t)T);
    memset(arg, 0, sizeof(uint16_t) * (size_t)T);
    for (int t = 0; t < T - L + 1; ++t)
        for (int m = 0; m < M; ++m) {
            float acc = cs[m];
            for (int tau = 0; tau < L; ++tau) {
                const float *f = Fm + (size_t)(t + tau) * row;
                const float *w = LF + ((size_t)m * L + tau) * row;
                for (size_t d = 0; d < row; ++d) {
                    volatile float prod = w[d] * f[d];
                    acc += prod;
                }
            }
            if (m == 0 || acc > out[t]) { out[t] = acc; arg[t] = (uint16_t)m; }
        }
}

/* ---- compute_likelihood_linear_filter.pyx:79-131 detect_examples (intended semantics;
 *      the reference allocates float32 into a float64-typed buffer and raises) and
 *      :21-76 get_max_detect_examples.  F (n,T,D1,D2) u8, LF (L,D1,D2) f64. ----------- */
void oracle_detect_examples(const uint8_t *Fm, int n, int T, int D1, int D2, const double *LF, int L,
                            const int64_t *lengths, int max_length, double *out /* n*max_length */,
                            double *maxout /* n or NULL */)
{
    const size_t row = (size_t)D1 * D2;
    if (out) memset(out, 0, sizeof(double) * (size_t)n * max_length);
    for (int e = 0; e < n; ++e) {
        double best = 0.0;
        for (int t = 0; t < lengths[e]; ++t) {
            int max_time = T - t < L ? T - t : L;
            double acc = 0.0;
            for (int tau = 0; tau < max_time; ++tau) {
                const uint8_t *f = Fm + ((size_t)e * T + t + tau) * row;
                const double *w = LF + (size_t)tau * row;
                for (size_t d = 0; d < row; ++d)
                    if (f[d]) acc += w[d];
            }
            if (out) out[(size_t)e * max_length + t] = acc;
            if (t == 0 || acc > best) best = acc;
        }
        if (maxout) maxout[e] = best;
    }
}

/* ---- edge_signal_proc.py:1406-1417 _spread_edges on one orientation block ----------
 * Tm (H, W) 0/1 doubles; out[i,j] = max_{k=-s..s} Tm[i+k*d0, j+k*d1], zero outside.
 * (The reference zero-pads by 2s; the padded reads are exactly the out-of-range zeros.) */
void oracle_spread_edges(const double *Tm, int H, int W, int d0, int d1, int s, double *out)
{
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            double m = 0.0;
            for (int k = -s; k <= s; ++k) {
                int ii = i + k * d0, jj = j + k * d1;
                if (ii >= 0 && ii < H && jj >= 0 && jj < W) {
                    double v = Tm[(size_t)ii * W + jj];
                    if (v > m) m = v;
                }
            }
            out[(size_t)i * W + j] = m;
        }
}

/* ---- edge_signal_proc.py:1623-1632 'all' branch: (2s+1)^2 box max, constant 0 border */
void oracle_spread_edges_box(const double *Tm, int H, int W, int s, double *out)
{
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            double m = 0.0;
            for (int a = -s; a <= s; ++a)
                for (int b = -s; b <= s; ++b) {
                    int ii = i + a, jj = j + b;
                    if (ii >= 0 && ii < H && jj >= 0 && jj < W) {
                        double v = Tm[(size_t)ii * W + jj];
                        if (v > m) m = v;
                    }
                }
            out[(size_t)i * W + j] = m;
        }
}

/* ------------------------------------------------------------------ detection-time clustering
 * cluster_times.pyx:133-219 (cluster_times): partition[i] marks times[i] as the start of a cluster
 * (2 = final, 1 = may still need splitting while the loop runs), partition[n] = 2 closes the list.
 * Peaks at least C0 frames apart start a new cluster; a cluster whose extent exceeds C1 is cut at its
 * widest gap.  The reference keeps times and distances in uint16 and C0 / C1 in float32. */
void oracle_cluster_times(const uint16_t *times, int n, float C0, float C1, uint8_t *part)
{
    for (int i = 0; i <= n; ++i) part[i] = 0;
    part[0] = 1;
    part[n] = 2;
    for (int i = 1; i < n; ++i)
        if ((float)(times[i] - times[i - 1]) >= C0) part[i] = 1;
    int start = 0;
    int widest = 0;                                  /* the reference leaves this uninitialised; unused before it is set */
    while (start < n) {
        if (part[start + 1] > 0) {                   /* single-element cluster */
            part[start] = 2;
            ++start;
            while (part[start] != 1 && start < n) ++start;
            continue;
        }
        uint16_t max_gap = 0, extent = 0;
        int k = 1;
        while (part[start + k] == 0) {
            uint16_t gap = (uint16_t)(times[start + k] - times[start + k - 1]);
            if (gap > max_gap) { widest = start + k; max_gap = gap; }
            extent = (uint16_t)(extent + gap);
            ++k;
        }
        if ((float)extent > C1) {
            part[widest] = ((float)(times[start + k - 1] - times[widest]) <= C1) ? 2 : 1;
            if ((float)(times[widest - 1] - times[start]) <= C1) {
                part[start] = 2;
                if (part[widest] == 1) start = widest;
            } else {
                part[start] = 1;
            }
        } else {
            part[start] = 2;
            start += k;
            while (part[start] != 1 && start < n) ++start;
        }
    }
}
