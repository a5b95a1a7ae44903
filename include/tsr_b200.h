/*
 * tsr_b200.h -- C ABI of libtsr_b200.so, the B200 (sm_100a) implementation of the
 * template_speech_rec data-parallel scoring path.
 *
 * The reference has no FFI: its boundary is "import code_parts", "import
 * spread_waliji_patches", "import compute_likelihood_linear_filter" (Cython modules built in
 * place, template_speech_rec/setup.py:35-41) plus the pure-Python classes
 * bernoulli_mixture.BernoulliMixture / bernoulli_em.Bernoulli_Mixture.  Each entry point
 * below is what a binding for one of those reference functions calls; the reference
 * function it replaces is cited as file:line (paths relative to the reference root).
 * INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *   - every array pointer is a DEVICE pointer unless the parameter is named host_*;
 *   - every call is asynchronous on the caller's `stream` (a cudaStream_t passed as void*);
 *   - return value: TSR_OK (0) or a negative TSR_ERR_*; tsr_last_error() gives the text;
 *   - no exceptions, no allocation of caller-visible memory, no hidden synchronisation
 *     (the *_blob_bytes / *_rowwords helpers let the caller size every buffer);
 *   - there is NO CPU fallback: on a device that is not compute capability 10.x the compute
 *     entry points return TSR_ERR_UNSUPPORTED.
 *
 * Packed representations (resident in HBM between stages)
 *   edge bits   uint8  [B][T][F]            bit e of byte (t,f) = X[t,f,e], e < E <= 8
 *   codes       uint8  [B][To][Fo]          0 = background, p in 1..K = best part (K <= 255)
 *   feature bits uint32 [B][T][rowwords]    bit (d % 32) of word (d / 32) = feature d of the row;
 *                                           rowwords = tsr_rowwords(D) (even: rows padded to 64 features)
 */
#ifndef TSR_B200_H
#define TSR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSR_OK               0
#define TSR_ERR_INVALID     -1   /* bad argument (shape, range, null pointer) */
#define TSR_ERR_CUDA        -2   /* a CUDA runtime call or launch failed */
#define TSR_ERR_UNSUPPORTED -3   /* not an sm_100 device, or shape outside the kernel's limits */

#define TSR_ABI_VERSION 2

/* detection kernels selectable through `impl` arguments (per-stage choice) */
#define TSR_IMPL_TENSOR 0   /* tcgen05 / TMEM path (default) */
#define TSR_IMPL_SIMT   1   /* CUDA-core cross-check path, fp32 sequential accumulation */
/* tsr_code_parts only: tcgen05 kind::i8 path.  All parts share one 16-bit fixed-point scale (two balanced base-255
 * int8 digits that accumulate into the same int32 column, half the MMAs of the 16-bit hi+lo form), integer arg-max
 * with the runner-up tracked; every location whose runner-up lies within the rounding bound (set bits + 2 units) is
 * re-evaluated by a second kernel with the reference's own float64 sum (code_parts.pyx:123-131).  Codes are therefore
 * the reference's codes with no near-tie exception; scores_or_null, if requested, carry the 16-bit fixed point. */
#define TSR_IMPL_TENSOR_INT8 2

/* 16-bit float format of the hi+lo weight planes (and of the expanded {0,1} bits) on the tensor
 * path; both run at the same tcgen05 kind::f16 rate.  AUTO picks FP16 (22 mantissa bits over the
 * two planes) when max|w| <= 1024, else BF16 (16 bits, fp32 range).  Decided on the device. */
#define TSR_PLANES_AUTO 0
#define TSR_PLANES_BF16 1
#define TSR_PLANES_FP16 2
/* Detection only: fixed-point filters q = round(w / scale[m]), scale[m] = max|w_m| / qmax, as 2
 * (16-bit, qmax 32639) or 3 (24-bit, qmax 8355711) balanced base-256 int8 digit planes, uint8 {0,1}
 * bits, tcgen05 kind::i8 with exact int32 accumulation (twice the f16 rate).  INT8X2 runs on
 * detect_tn_kernel (frames as UMMA N, ping-pong accumulators; the scoring pipeline's default:
 * |err| <= 0.5 * scale[m] * set bits, 2.6e-6 of the score scale for log-odds filters), INT8X3 on
 * detect_tc_kernel (6e-8). */
#define TSR_PLANES_INT8X2 3
#define TSR_PLANES_INT8X3 4

int         tsr_abi_version(void);
const char* tsr_last_error(void);            /* thread-local, valid until the next failing call */
int         tsr_device_supported(void);      /* 1 if the current device is compute capability 10.x */
int         tsr_rowwords(int D);             /* uint32 words per packed feature row (multiple of 2) */

/* ---- K0: boundary packing (NumPy-layout uint8 <-> packed bits) ------------------------- */
/* X uint8 [B][T][F][E] (any nonzero = edge, as `if X[i,j,z]` code_parts.pyx:29) -> bits [B][T][F] */
int tsr_pack_edges(const uint8_t* X, uint8_t* bits, int B, int T, int F, int E, void* stream);
int tsr_unpack_edges(const uint8_t* bits, uint8_t* X, int B, int T, int F, int E, void* stream);
/* The same packing for edge tensors in HOST memory (both pointers are host pointers; E must be 8), on `nthreads`
 * host threads of a persistent pool: callers that stream from the host pack before the copy and move 1 byte per
 * (t,f) instead of 8 (ScoringPipeline.stream).  Pure data movement; no arithmetic of the path runs on the host. */
int tsr_host_pack_edges(const uint8_t* X_host, uint8_t* bits_host, long long n_cells, int E, int nthreads);
/* F uint8 [R][D] (nonzero test, compute_likelihood_linear_filter.pyx:168) -> bits [R][rowwords] */
int tsr_pack_features_u8(const uint8_t* Fm, uint32_t* bits, long long R, int D, int rowwords, void* stream);
int tsr_unpack_features_u8(const uint32_t* bits, uint8_t* Fm, long long R, int D, int rowwords, void* stream);

/* ---- K1a: edge spreading --------------------------------------------------------------
 * Replaces edge_signal_proc.spread_edgemap/_spread_edges (edge_signal_proc.py:1397-1417)
 * and the 'all' box branch (edge_signal_proc.py:1623-1632, 1664-1673).
 * out bit e at (t,f) = OR_{k=-radius..radius} in bit e at (t + k*dt[e], f + k*df[e]), zero
 * outside the utterance and outside the `block_len`-frame block of t (block_len = 0: no
 * blocks; the reference spreads per 40-frame block, edge_signal_proc.py:1635-1663).
 * tail_window != 0: a ragged last block (T % block_len frames) is spread as the tail of the window of the final
 *         block_len frames, as _edge_map_threshold_segments does (:1595-1633), instead of as a short block of its own.
 * mode 0: line along host_dirs[e] = {df (frequency step), dt (time step)} per orientation
 *         (int8 [8][2], HOST pointer, copied at launch);  mode 1: (2r+1)^2 box, dirs ignored. */
int tsr_spread_edges(const uint8_t* bits_in, uint8_t* bits_out, int B, int T, int F, int radius, int mode,
                     const int8_t* host_dirs, int block_len, int tail_window, void* stream);

/* ---- K2: part coding --------------------------------------------------------------------
 * Replaces code_parts.code_parts (code_parts.pyx:70-135), code_parts_fast (:247-295),
 * code_parts_mask (:353-404) and the np.argmax the callers apply (get_train_data.py:2621).
 * tsr_prepare_parts turns the reference's float64 log_parts / log_invparts [K][ph][pw][E]
 * (device pointers) into the kernel's weight blob: W = lp - lip split into 16-bit hi + lo
 * planes (TSR_PLANES_*) in UMMA core-matrix order, c[k] = sum(lip) in fp32, edge-count columns -- and, for
 * TSR_IMPL_TENSOR_INT8, the same model in one fixed-point scale (two base-255 int8 digits per weight, integer constants)
 * together with the float64 log tables its refinement sums in the reference's order. */
size_t tsr_parts_blob_bytes(int K, int ph, int pw);
int    tsr_prepare_parts(const double* log_parts, const double* log_invparts, int K, int ph, int pw, int E,
                         const uint8_t* mask_or_null /* [ph][pw], code_parts_mask's M */, int plane_format,
                         void* blob, void* stream);
/* codes[b][i][j] = 0 if gate fails else 1 + argmax_k score (lowest k wins ties);
 * gate: count >= threshold [and mask_count > mask_threshold when mask_threshold >= 0].
 * scores_or_null float [B][To][Fo][K] (score = sum_{a,b,z} X ? lp : lip, written for every
 * location); counts_or_null uint16 [B][To][Fo] edge counts.  To = T-ph+1, Fo = F-pw+1.
 * edge_spread_radius = 1 (tensor path): K1a fused into the kernel's producers -- the input is the UNSPREAD edge
 * map and every byte is spread on the fly exactly as tsr_spread_edges(radius 1, mode 0, the unit direction table of
 * edge_signal_proc.py:1751-1821, block_len = edge_block_len) would, with the reference's ragged-last-block rule
 * (the frames past the last whole block belong to the window of the final block_len frames, :1595-1633).
 * edge_spread_radius = 0: bits are coded as given.
 * impl: TSR_IMPL_TENSOR (hi+lo planes; codes may differ from the reference where two parts score within 1e-5 relative),
 * TSR_IMPL_TENSOR_INT8 (codes equal the reference's arg-max at every location; takes B*To*Fo*4 + B*To*Fo + 8 KB bytes of
 * stream-ordered scratch through cudaMallocAsync for the call; needs B*To*Fo < 2^32), TSR_IMPL_SIMT (cross-check). */
int tsr_code_parts(const uint8_t* bits, const void* parts_blob, int K, int ph, int pw, int threshold,
                   int mask_threshold, uint8_t* codes, float* scores_or_null, uint16_t* counts_or_null, int B, int T,
                   int F, int edge_spread_radius, int edge_block_len, int impl, void* stream);

/* Replaces the magnitude / auxiliary Bernoulli terms of code_parts.code_parts_mag_aux
 * (code_parts.pyx:138-242).  scores float [B][To][Fo][K] holds the edge log-likelihoods written by
 * tsr_code_parts (scores_or_null); on return every gated location (thresh_mask[b][i+ph/2][j+pw/2] != 0,
 * :191-192) additionally carries, per part k, the sum over the ph window rows of the auxiliary terms
 * (X_aux uint8 [B][T][aux_dim], log_aux_* double [K][ph*aux_dim], :199-206), the band-magnitude term
 * (X_mags uint8 [B][T][mag_dim], log_mag_* double [K][ph+mag_dim], :208-210, :232-237) and the
 * centre-band profile (:218-228); ungated locations are set to 0.  gate_or_null uint8 [B][To][Fo]. */
int tsr_code_parts_mag_aux_terms(float* scores, uint8_t* gate_or_null, const uint8_t* X_mags, const uint8_t* X_aux,
                                 const uint8_t* thresh_mask, const double* log_mag_parts, const double* log_mag_invparts,
                                 const double* log_aux_parts, const double* log_aux_invparts, int B, int T, int F, int K,
                                 int ph, int pw, int mag_dim, int aux_dim, void* stream);

/* ---- N1: patch extraction for part-dictionary EM ---------------------------------------------
 * Replaces code_parts.get_parts_mask (code_parts.pyx:297-348) and the edge-count selection of
 * extract_local_features_tied (extract_local_features.py:36-38).
 * tsr_patch_counts: counts[b][i][j] = edges of the ph x pw window at (i,j) under the window-relative
 *   0/1 mask (mask_or_null uint8 [ph][pw], device; null = all ones, _count_edges :18-31).
 * tsr_patch_select: slot[idx] = rank of location idx among the locations with lo <= count < hi in
 *   row-major (b,i,j) order -- the reference's patch_id (:331-346) -- or -1; *total = their number.
 *   workspace: tsr_patch_select_workspace_bytes(n_loc) bytes.
 * tsr_patch_gather: for every selected location writes patch_set uint8 [n][ph][pw][E] (0/1),
 *   S_patch double [n][ph][pw] (from S double [B][T][F]), locs uint32 [n][3] = (b,i,j) and/or
 *   patchbits uint32 [n][tsr_rowwords(ph*pw*E)] (feature (a*pw+c)*E+z: the EM kernels' input). */
int    tsr_patch_counts(const uint8_t* bits, const uint8_t* mask_or_null, int B, int T, int F, int ph, int pw,
                        uint32_t* counts, void* stream);
size_t tsr_patch_select_workspace_bytes(long long n_loc);
int    tsr_patch_select(const uint32_t* counts, long long n_loc, unsigned int lo, unsigned int hi, int32_t* slot,
                        uint32_t* total, void* workspace, void* stream);
int    tsr_patch_gather(const uint8_t* bits, const double* S_or_null, const int32_t* slot, int B, int T, int F, int E,
                        int ph, int pw, uint8_t* patch_set_or_null, double* S_patch_or_null, uint32_t* locs_or_null,
                        uint32_t* patchbits_or_null, void* stream);

/* ---- K1b: part-code spreading -----------------------------------------------------------
 * Replaces code_parts.spread_patches (code_parts.pyx:406-457) == spread_waliji_patches
 * (spread_waliji_patches.pyx:15-59).  featbits row (b,i) holds Fo*K features, feature
 * j*K + (p-1) set iff part p occurs within |di|<=r0, |dj|<=r1 of (i,j) (clipped). */
int tsr_spread_codes(const uint8_t* codes, uint32_t* featbits, int B, int To, int Fo, int K, int r0, int r1,
                     int rowwords, void* stream);
/* Replaces code_parts.spread_graph_patches (code_parts.pyx:459-503): in place, sequential in
 * p1 within each cell.  graph uint8 [K][K]. */
int tsr_spread_graph(uint32_t* featbits, const uint8_t* graph, long long R, int Fo, int K, int rowwords, void* stream);

/* ---- K3: sliding Bernoulli-template LLR detection ------------------------------------------
 * Replaces compute_likelihood_linear_filter.detect (compute_likelihood_linear_filter.pyx:
 * 135-176), the per-component loop + np.maximum of get_train_data._compute_detection_E
 * (get_train_data.py:964-992) and detect_float_max's fused max/argmax (:221-278).
 * tsr_prepare_filters: LF float32 [M][L][D] -> blob of 16-bit hi/lo planes (TSR_PLANES_*) in UMMA order. */
size_t tsr_filter_blob_bytes(int M, int L, int D, int plane_format);
int    tsr_prepare_filters(const float* LF, int M, int L, int D, int plane_format, void* blob, void* stream);
/* featbits [B][T][rowwords].  scores float [B][T][M]: scores[b][t][m] = sum_{tau,d}
 * bit(b,t+tau,d) * LF[m][tau][d] for t <= T-L, 0 for the L-1 tail frames (detect's contract,
 * :157-158).  impl = TSR_IMPL_TENSOR | TSR_IMPL_SIMT (SIMT needs LF_or_null = the fp32 filters). */
int tsr_detect(const uint32_t* featbits, int B, int T, int D, int rowwords, const void* filter_blob,
               const float* LF_or_null, int M, int L, int plane_format /* as given to tsr_prepare_filters */,
               float* scores, int impl, void* stream);
/* best[b][t] = max_m (scores[b][t][m] + cs[m]), arg = first m attaining it (strict >, :275);
 * tail frames: tail_mode 0 -> best 0, arg 0 (detect_float_max); 1 -> max_m cs[m] with its arg
 * (what get_train_data.py:962-992 produces by adding c to the zero tail). */
int tsr_detect_reduce(const float* scores, const float* cs, int B, int T, int M, int L, int tail_mode, float* best,
                      uint16_t* arg, void* stream);
/* tsr_detect (tensor path) with the mixture max fused into its epilogue: best / arg as tsr_detect_reduce with
 * tail_mode 1 computes them (bit-identical), without materialising the [B][T][M] score matrix -- the running
 * np.maximum of get_train_data.py:964-992 / detect_float_max :268-277.  scores_or_null: also write the matrix.
 * keys_workspace: B*T*8 bytes of device scratch (64-bit (value, template) keys combined with atomicMax across the
 * CTAs that hold different template tiles). */
int tsr_detect_best(const uint32_t* featbits, int B, int T, int D, int rowwords, const void* filter_blob, int M, int L,
                    int plane_format, const float* cs_or_null, float* scores_or_null, void* keys_workspace, float* best,
                    uint16_t* arg, void* stream);
/* Real-valued features: detect_float (:178-219), detect_float2d (:281-318), detect_spectral
 * (:322-359), detect_float_max (:221-278).  fp32, separate multiply and add in the reference's
 * (tau,f,k) order; each accumulation starts at init_or_null[m] (detect_float_max's cs, :258) or 0. */
int tsr_detect_float(const float* Fm, int B, int T, int D, const float* LF, int M, int L, const float* init_or_null,
                     float* scores, void* stream);

/* ---- N2: peak picking on detection tracks ------------------------------------------------------
 * Replaces roc_functions.get_max_above_threshold (roc_functions.py:177-186): flags[b][t] = 1 iff
 * 1 <= t <= l-2, track[b][t] >= both neighbours and >= threshold, l = length[b] - C0 + 1
 * (lengths_or_null int32 [B]; null = T).  The flagged frames are compacted in order with
 * tsr_patch_select(flags, B*T, 1, 0xFFFFFFFF, ...) and written out by tsr_peaks_gather as
 * peak_bt int32 [n][2] = (b, t) and peak_val float [n]: detections instead of T-length tracks. */
int tsr_detect_peaks(const float* track, const int* lengths_or_null, int B, int T, int C0, float threshold, uint32_t* flags,
                     void* stream);
int tsr_peaks_gather(const float* track, const int32_t* slot, int B, int T, int32_t* peak_bt, float* peak_val_or_null,
                     void* stream);

/* N2, second half: clustering of the picked peaks, one utterance per thread.  Replaces cluster_times.cluster_times
 * (cluster_times.pyx:133-219) and roc_functions.get_cluster_starts_ends (roc_functions.py:187-212).
 * peak_bt int32 [n][2] = (b, t) sorted by b then t (tsr_peaks_gather's output), t < 65536.
 * offsets int32 [B+1]: utterance b's peaks are rows offsets[b] .. offsets[b+1]-1;
 * partition uint8 [n+B]: utterance b's n_b+1 codes start at offsets[b]+b (2 = cluster start, 0 = continuation,
 *   closing 2), exactly the reference's cluster_partition;
 * clusters int32 [n][2]: (first peak, last peak + 1) of utterance b's clusters from row offsets[b]; ncluster int32 [B]. */
int tsr_cluster_peaks(const int32_t* peak_bt, int n, int B, float C0, float C1, int32_t* offsets, uint8_t* partition,
                      int32_t* clusters, int32_t* ncluster, void* stream);

/* ---- N4 / N3: edge front end (feeds K1a) ------------------------------------------------------
 * tsr_edge_map replaces edge_signal_proc._edge_map_no_threshold (edge_signal_proc.py:1697-1836):
 *   S double [B][F][T] (frequency rows x time) -> mags double [B][8][F][T], the reference's (8H, time)
 *   array per utterance; orientation order and the zero borders as the reference writes them.
 * tsr_threshold_edges replaces threshold_edge_block / threshold_edgemap (:1184-1307) applied per
 *   `block_len`-frame block as _edge_map_threshold_segments does (:1537-1675), without critical bands:
 *   per (block, orientation) tau = sorted(positive magnitudes)[int(quantile * n)], an edge is set where
 *   mag > 0 and mag >= max(tau, host_abst_threshold[e]) (double [8], HOST pointer).  Output: packed edge
 *   bytes bits [B][T][F] (the input of tsr_spread_edges).  A ragged last block is thresholded on the
 *   window of the final block_len frames (:1595-1633); that whole window goes to
 *   tail_bits_or_null [B][min(T,block_len)][F] so the caller can spread it as the reference does.
 *   taus_or_null double [B][ceil(T/block_len)][8] receives tau (report_level; 0 for an empty block).
 * tsr_edges_from_spectrogram fuses both: spectrogram in, packed edge bytes out, magnitudes never stored. */
int tsr_edge_map(const double* S, double* mags, int B, int F, int T, void* stream);
int tsr_threshold_edges(const double* mags, uint8_t* bits, uint8_t* tail_bits_or_null, double* taus_or_null, int B, int F,
                        int T, int block_len, double quantile, const double* host_abst_threshold, void* stream);
int tsr_edges_from_spectrogram(const double* S, uint8_t* bits, uint8_t* tail_bits_or_null, double* taus_or_null, int B, int F,
                               int T, int block_len, double quantile, const double* host_abst_threshold, void* stream);

/* ---- K4: Bernoulli-mixture EM -----------------------------------------------------------------
 * Replaces BernoulliMixture.get_template_loglikelihoods/_compute_loglikelihoods
 * (bernoulli_mixture.py:263-291), M_step/threshold_templates (:177-193), the run_EM loop's stopping rule
 * (:139-156) and the same steps of bernoulli_em.Bernoulli_Mixture (bernoulli_em.py:22-49, 103-127).
 * Any number of components 1 <= K <= 4095, fewer than 2^24 rows per rank (the reference trains 100-200-part dictionaries with this class,
 * Notebook/12/reestimate_parts.py:79-83).
 *
 * Sufficient statistics are ONE float64 vector  stats[K*D + K + 1]:
 *   stats[k*D+d]   = sum_n aff[n][k] x[n][d]                     (numerators)
 *   stats[K*D+k]   = sum_n aff[n][k]                             (responsibility sums)
 *   stats[K*D+K]   = sum_n logsumexp_k(S[n][k] + logw[k])        (log-likelihood)
 * Inside a rank every sum is an exact integer sum (posteriors quantised per component to 24 bits below a
 * power-of-two scale s_k >= max_n aff[n][k], so components holding a vanishing share of the mass keep full
 * relative precision; row log-likelihoods in 2^-20 fixed point), converted to float64 once: a rank's statistics
 * are deterministic and independent of the summation order.  The caller's single allreduce(sum, float64) over the
 * ranks (NCCL) combines them; tsr_em_finalize then runs identically on every rank.
 *
 * E-step over this rank's rows: S[n][k] = sum_d x ? logT : logInvT; aff = row-softmax(S + logw) (or the one-hot
 * first arg-max of S when hard != 0).  With update != 0 it writes aff_or_null [N][K], the posterior digit planes
 * the next tsr_em_mstep uses, and the last K+1 entries of stats; with update == 0 only scores_or_null [N][K] is
 * written and no EM state is touched.  logT_or_null == NULL: use the weight planes the last
 * tsr_em_finalize(workspace, N_local) call laid out in this workspace.
 * M-step accumulation: stats[0 .. K*D) = numerators from the posteriors of the last E-step (aff_or_null == NULL)
 * or from the given aff [N][K] (then stats[K*D .. K*D+K) is set from it too).
 * ctl_or_null: device control block (tsr_em_control_*); once it has stopped, every call that carries it is a no-op
 * on the device, so the host may queue iterations without reading the log-likelihood back. */
size_t    tsr_em_workspace_bytes(long long N, int D, int K);
int       tsr_em_workspace_init(void* workspace, long long N, int D, int K, void* stream);   /* once, before first use */
long long tsr_em_npad(long long N);                                                           /* N rounded up to 256 */
/* XT uint32 [D][N_pad/32] = the bit matrix transposed once per data set (tensor-path M-step), N_pad = tsr_em_npad(N) */
int tsr_transpose_bits(const uint32_t* Xbits, long long N, int D, int rowwords, uint32_t* XTbits, long long N_pad,
                       void* stream);
/* Tile order of a bit matrix [R][rw]: [128-row tile][16-byte chunk][128 rows][4 words], zero padded to a multiple of 256
 * rows -- what the tensor-core contraction reads for one chunk of one row tile is then one contiguous 2 KB block
 * (sequential HBM reads; 16 bytes per row at a stride of the row pitch ran at 18 % of the HBM peak).  Built once per
 * data set for X (E-step) and for the transposed X (M-step). */
size_t tsr_em_tiled_words(long long R, int rw);
int    tsr_em_tile_bits(const uint32_t* bits, long long R, int rw, uint32_t* tiled, void* stream);
int tsr_em_estep(const uint32_t* Xbits, const uint32_t* Xtiled_or_null, long long N, int D, int rowwords, int K, const float* logT_or_null,
                 const float* logInvT_or_null, const float* logw, float* aff_or_null, float* scores_or_null,
                 double* stats_or_null, int hard, int update, void* workspace, const void* ctl_or_null, int impl,
                 void* stream);
int tsr_em_mstep(const uint32_t* Xbits, const uint32_t* XTbits_or_null, const uint32_t* XTtiled_or_null /* one of the two for TSR_IMPL_TENSOR */,
                 long long N, int D, int rowwords, int K, const float* aff_or_null, double* stats, void* workspace,
                 const void* ctl_or_null, int impl, void* stream);
/* weights = respsum / N_total; T = numer / N_total / weights clipped to [min_prob, 1-min_prob]; logT, logInvT,
 * logw refreshed (all float32, as bernoulli_mixture.py:178-184).  workspace_or_null != NULL: also lays out the
 * int8 weight planes of the next E-step over N_local rows in that workspace. */
int tsr_em_finalize(const double* stats, long long N_total, int D, int K, float min_prob, float* templates,
                    float* logT, float* logInvT, float* weights, float* logw, void* workspace_or_null,
                    long long N_local, const void* ctl_or_null, void* stream);
/* Sum of the ranks' statistics vectors over NVLink peer memory in ONE kernel launch (replaces the NCCL allreduce of
 * em_driver.py, i.e. the sum the reference would form over all rows in bernoulli_mixture.py:177-181 when rows are sharded).
 * Every rank owns one symmetric allocation mapped by all peers (torch.distributed._symmetric_memory); peer_bufs_dev is a
 * DEVICE array of the world base pointers in rank order.  Inside each allocation: the source vector (n_doubles float64) at
 * src_offset_bytes and 2 x world x 64 uint32 flags (tsr_em_peer_flag_bytes, zeroed once) at flag_offset_bytes.  out (local,
 * 16-byte aligned) receives sum_r src_r, added in rank order on every rank (bit-identical everywhere).  epoch_counters:
 * tsr_em_peer_epoch_bytes() bytes of local device memory, zeroed once.  The caller must alternate between TWO source
 * offsets from launch to launch (that is what makes the kernel safe without a closing barrier) and every rank must issue
 * the same sequence of launches.  Honours ctl_or_null like the other EM entries.
 * out_offset_bytes >= 0 selects the two-phase form (for more than two ranks): rank r sums only its 1/world share and
 * stores it into the result slot at out_offset_bytes of EVERY rank's allocation, a second flag round closes the launch;
 * `out` is then ignored and the result is read from the caller's own allocation. */
size_t tsr_em_peer_flag_bytes(int world);
size_t tsr_em_peer_epoch_bytes(void);
int    tsr_em_allreduce_peer(const void* const* peer_bufs_dev, int rank, int world, long long src_offset_bytes,
                             long long flag_offset_bytes, long long n_doubles, double* out, long long out_offset_bytes,
                             void* epoch_counters, const void* ctl_or_null, void* stream);
/* Device-side stopping rule.  ctl = tsr_em_control_bytes(max_trace) bytes of device memory:
 *   int32 done, iterations, ntrace, mode, min_iter, max_iter, cap, pad; double tol, ll_old; double trace[max_trace].
 * tsr_em_control_step (after the allreduce of stats) appends the log-likelihood to trace and applies
 *   mode 0: continue while (abs((old - new) / old) > tol and iterations < max_iter) or iterations < min_iter
 *   mode 1: continue while new - old > tol (and iterations < max_iter)
 * counting an iteration when it continues and setting done otherwise. */
size_t tsr_em_control_bytes(int max_trace);
int    tsr_em_control_init(void* ctl, int max_trace, double tol, int mode, int min_iter, int max_iter, void* stream);
int    tsr_em_control_step(void* ctl, const double* stats, int D, int K, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TSR_B200_H */
