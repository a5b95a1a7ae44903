"""The oracle (oracle/tsr_oracle.c + oracle/oracle.py) pinned against the reference:
golden vectors generated from the compiled reference Cython, the compiled reference itself
when present in this container, the docstring known-answer of BernoulliMixture, and
independent NumPy/SciPy formulations.  CPU only."""
import os

import numpy as np
import pytest

from tsr_b200 import synth


# ------------------------------------------------------------------ golden fixtures
def test_code_parts_golden(oracle, golden):
    g = golden
    out = oracle.code_parts(g["cp_X"], g["cp_lp"], g["cp_lip"], int(g["cp_thr"]))
    assert out.dtype == np.float64 and out.shape == g["cp_out"].shape
    np.testing.assert_array_equal(out, g["cp_out"])          # same order, same f64 arithmetic: bit-exact
    np.testing.assert_array_equal(oracle.code_parts_fast(g["cp_X"], g["cp_lp"], g["cp_lip"], 12), g["cp_out_fast"])
    np.testing.assert_array_equal(oracle.code_parts(g["cp_X"], g["cp2_lp"], g["cp2_lip"], 5), g["cp2_out"])
    # both sides of the threshold occur in the fixture
    gate = g["cp_out"][..., 0] != 0
    assert gate.any() and (~gate).any()


def test_code_parts_mag_aux_and_get_parts_mask_golden(oracle, golden):
    g = golden
    out = oracle.code_parts_mag_aux(g["cp_X"], g["ma_Xm"], g["ma_Xa"], g["cp_lp"], g["cp_lip"], g["ma_lmp"], g["ma_lmip"],
                                    g["ma_lap"], g["ma_laip"], g["ma_tm"])
    np.testing.assert_array_equal(out, g["ma_out"])          # same loop nest, same f64 arithmetic: bit-exact
    gate = g["ma_out"][..., 0] != 0
    assert gate.any() and (~gate).any()
    ps, Sp, pc, pl = oracle.get_parts_mask(g["cp_X"][8:13, 3:8], g["gp_M"], g["gp_S"], 3)
    for a, b in ((ps, g["gp_patches"]), (Sp, g["gp_Sp"]), (pc, g["gp_counts"]), (pl, g["gp_locs"])):
        assert a.dtype == b.dtype
        np.testing.assert_array_equal(a, b)


def test_spread_patches_golden(oracle, golden):
    g = golden
    np.testing.assert_array_equal(oracle.spread_patches(g["cp_codes"], 2, 2, 7), g["sp_22"])
    np.testing.assert_array_equal(oracle.spread_waliji_patches(g["cp_codes"], 1, 3, 7), g["sp_13"])
    np.testing.assert_array_equal(oracle.spread_patches(g["cp_codes"], 0, 0, 7), g["sp_00"])
    S = g["sp_00"].copy()
    oracle.spread_graph_patches(S, g["graph"])
    np.testing.assert_array_equal(S, g["sp_graph"])


def test_detect_golden(oracle, golden):
    g = golden
    np.testing.assert_array_equal(oracle.detect(g["dt_F"], g["dt_LF"]), g["dt_out"])      # f32 sequential: bit-exact
    np.testing.assert_array_equal(oracle.detect_float(g["df_F"], g["dt_LF"]), g["df_out"])
    mx, am = oracle.detect_float_max(g["df_F"], g["dm_LF"], g["dm_cs"])
    np.testing.assert_array_equal(mx, g["dm_max"])
    np.testing.assert_array_equal(am, g["dm_arg"])
    np.testing.assert_array_equal(oracle.detect_float2d(g["d2_F"], g["d2_LF"]), g["d2_out"])
    np.testing.assert_array_equal(oracle.detect_spectral(g["d2_F"], g["d2_LF"]), g["d2_spec"])
    np.testing.assert_array_equal(oracle.get_max_detect_examples(g["de_F"], g["de_LF"], g["de_len"]), g["de_max"])
    # detect's contract: length T, the L-1 tail entries are exactly 0
    L = g["dt_LF"].shape[0]
    assert g["dt_out"].shape == (g["dt_F"].shape[0],) and np.all(g["dt_out"][-(L - 1):] == 0)


# ------------------------------------------------- live differential vs compiled reference
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_against_compiled_reference(oracle, ref, seed):
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    cp, swp, clf = ref
    rs = np.random.RandomState(seed)
    T, F = rs.randint(8, 40), rs.randint(6, 20)
    K = rs.randint(1, 12)
    X = (rs.rand(T, F, 8) < rs.uniform(0.05, 0.3)).astype(np.uint8)
    P = np.clip(rs.rand(K, 5, 5, 8) ** 2, .01, .99)
    lp, lip = np.log(P), np.log(1 - P)
    thr = int(rs.randint(0, 30))
    a, b = oracle.code_parts(X, lp, lip, thr), cp.code_parts(X, lp, lip, thr)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(oracle.code_parts_fast(X, lp, lip, thr), cp.code_parts_fast(X, lp, lip, thr))
    codes = np.argmax(b, -1)
    r0, r1 = rs.randint(0, 4), rs.randint(0, 6)
    np.testing.assert_array_equal(oracle.spread_patches(codes, r0, r1, K), cp.spread_patches(codes, r0, r1, K))
    np.testing.assert_array_equal(oracle.spread_patches(codes, r0, r1, K), swp.spread_waliji_patches(codes, r0, r1, K))
    S = cp.spread_patches(codes, r0, r1, K)
    L = rs.randint(1, T - 4)
    LF = rs.randn(L, S.shape[1], K).astype(np.float32)
    np.testing.assert_array_equal(oracle.detect(S, LF), clf.detect(S, LF))
    # first-window mask count agrees with the reference (its later windows read M out of bounds)
    M = (rs.rand(5, 5) < 0.5).astype(np.uint8)
    o = oracle.code_parts_mask(X[:5, :5], M, lp, lip, 0, 0)
    r = cp.code_parts_mask(X[:5, :5].copy(), M, lp, lip, 0, 0)
    np.testing.assert_array_equal(o, r)


@pytest.mark.parametrize("seed", [0, 1])
def test_mag_aux_against_compiled_reference(oracle, ref, seed):
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rs = np.random.RandomState(100 + seed)
    T, F, K, md, ad = rs.randint(8, 30), rs.randint(6, 20), rs.randint(1, 9), rs.randint(1, 9), rs.randint(0, 4)
    X = (rs.rand(T, F, 8) < 0.2).astype(np.uint8)
    Xm, Xa, tm = ((rs.rand(*sh) < 0.5).astype(np.uint8) for sh in ((T, md), (T, ad), (T, F)))
    tabs = []
    for sh in ((K, 5, 5, 8), (K, 5 + md), (K, 5 * ad)):
        P = np.clip(rs.rand(*sh), .01, .99)
        tabs += [np.log(P), np.log(1 - P)]
    a = oracle.code_parts_mag_aux(X, Xm, Xa, *tabs, tm)
    b = ref[0].code_parts_mag_aux(X, Xm, Xa, *tabs, tm)
    np.testing.assert_array_equal(a, b)


def test_reference_rejects_float32_parts(ref):
    if ref is None:
        pytest.skip("oracle/_ref not built")
    X = np.zeros((6, 6, 8), np.uint8)
    lp = np.zeros((2, 5, 5, 8), np.float32)
    with pytest.raises(ValueError, match="Buffer dtype mismatch"):
        ref[0].code_parts(X, lp, lp, 1)


# ----------------------------------------------------- independent formulations
def test_code_parts_equals_correlation(oracle):
    """Development/092412/scripts/test_new_part_coding.py:9-37: scores are
    correlate(E, lp) + correlate(1-E, lip) over the valid region."""
    from scipy.signal import correlate
    rs = np.random.RandomState(5)
    X = (rs.rand(14, 11, 8) < 0.2).astype(np.uint8)
    P = np.clip(rs.rand(3, 5, 5, 8) ** 2, .01, .99)
    lp, lip = np.log(P), np.log(1 - P)
    out = oracle.code_parts(X, lp, lip, 0)
    for k in range(3):
        ref = correlate(X.astype(float), lp[k], mode="valid")[..., 0] + correlate(1. - X, lip[k], mode="valid")[..., 0]
        np.testing.assert_allclose(out[..., k + 1], ref, rtol=1e-12, atol=1e-10)


def test_int8_coding_bound_restatement(oracle):
    """The argument behind TSR_IMPL_TENSOR_INT8 (code_parts.cu, prepare_parts_i8_kernel), restated in NumPy: with one
    fixed-point scale for all parts (q = rint(w / scale), cq = rint(c / scale) - mid) the integer accumulator of a part is
    within (set bits + 1) / 2 units of its exact score, so wherever the runner-up is more than set bits + 2 units behind,
    the integer arg-max IS the float64 oracle's arg-max; the GPU kernel re-evaluates every other location in float64."""
    lp, lip = synth.parts(40)
    W = (lp - lip).reshape(40, -1)
    c = lip.reshape(40, -1).sum(1)
    scale = np.abs(W).max() / 32512.0
    q = np.rint(W / scale).astype(np.int64)
    assert np.abs(q).max() <= 32512
    d1 = np.rint(q / 255.0).astype(np.int64)                 # two balanced base-255 digits, both int8
    d0 = q - 255 * d1
    assert np.abs(d0).max() <= 127 and np.abs(d1).max() <= 127
    cq = np.rint(c / scale).astype(np.int64) - int(np.rint((c.max() + c.min()) / 2 / scale))
    X = synth.utterance_batch(1, T=60)[0]
    ref = oracle.code_parts(X, lp, lip, 0)[..., 1:]
    P = np.lib.stride_tricks.sliding_window_view(X, (5, 5), axis=(0, 1)).transpose(0, 1, 3, 4, 2).reshape(ref.shape[0], ref.shape[1], -1)
    cnt = P.sum(-1).astype(np.int64)
    acc = P.astype(np.int64) @ (d0 + 255 * d1).T + cq           # what the kind::i8 MMAs accumulate (exactly)
    mid = int(np.rint((c.max() + c.min()) / 2 / scale))
    assert np.all(np.abs(ref / scale - mid - acc) <= 0.5 * (cnt[..., None] + 1) + 1e-6)
    srt = np.sort(acc, -1)
    decided = srt[..., -1] - srt[..., -2] > cnt + 2
    assert decided.mean() > 0.95 and (~decided).any()        # both branches occur in this sample
    np.testing.assert_array_equal(np.argmax(acc, -1)[decided], np.argmax(ref, -1)[decided])
    # |acc| * 256 + 255 stays inside int32 (the epilogue packs (acc, 255 - k) into one key)
    assert (200 + (c.max() - c.min()) / 2 / np.abs(W).max() + 1) * 32512 < 2 ** 23


def test_planted_parts_are_recovered(oracle):
    """Development/092412/scripts/test_waliji_coding.py:72-120 pattern: every 5x5x8 block is the
    most likely configuration of part j, so the code at block origins must be j+1."""
    K = 12
    rs = np.random.RandomState(11)
    P = np.where(rs.rand(K, 5, 5, 8) < 0.3, 0.9, 0.05)
    lp, lip = np.log(P), np.log(1 - P)
    X = np.zeros((5 * K, 5, 8), np.uint8)
    for j in range(K):
        X[5 * j:5 * j + 5] = P[j] > 0.5
    codes = np.argmax(oracle.code_parts(X, lp, lip, 1), -1)
    np.testing.assert_array_equal(codes[::5, 0], np.arange(K) + 1)


def test_spread_edges_vs_ndimage(oracle):
    from scipy.ndimage import maximum_filter
    rs = np.random.RandomState(3)
    T = (rs.rand(9, 17) < 0.15).astype(np.float64)
    for s in (1, 2, 3):
        np.testing.assert_array_equal(oracle._spread_edges(T, (1, 0), s), maximum_filter(T, size=(2 * s + 1, 1), mode="constant"))
        np.testing.assert_array_equal(oracle._spread_edges(T, (0, -1), s), maximum_filter(T, size=(1, 2 * s + 1), mode="constant"))
        fp = np.eye(2 * s + 1)
        np.testing.assert_array_equal(oracle._spread_edges(T, (1, 1), s), maximum_filter(T, footprint=fp, mode="constant"))
        np.testing.assert_array_equal(oracle._spread_edges(T, (-1, 1), s), maximum_filter(T, footprint=fp[::-1], mode="constant"))
        E = np.vstack([T, T])
        oracle.spread_edgemap_all(E, [0, 9, 18], s)
        np.testing.assert_array_equal(E[:9], maximum_filter(T, size=(2 * s + 1, 2 * s + 1), mode="constant"))


def test_linear_filter_identities(oracle):
    """Development/071912/fast_like/test_fast_like.py:61-75: W, C against the closed forms and
    sum(E*W)+C == log P(E|T) - log P(E|bgd)."""
    rs = np.random.RandomState(7)
    T = rs.uniform(.05, .95, (6, 4, 3))
    bg = rs.uniform(.05, .4, (4, 3))
    W, c = oracle.construct_linear_filter(T, bg)
    assert W.dtype == np.float32
    np.testing.assert_allclose(W, np.log(T / (1 - T)) - np.log(bg / (1 - bg)), rtol=1e-6)
    np.testing.assert_allclose(c, np.log((1 - T) / (1 - bg)).sum(), rtol=1e-12)
    E = (rs.rand(6, 4, 3) < 0.3).astype(float)
    llr = (E * np.log(T) + (1 - E) * np.log(1 - T)).sum() - (E * np.log(bg) + (1 - E) * np.log(1 - bg)).sum()
    np.testing.assert_allclose((E * W).sum() + c, llr, rtol=1e-5)
    Wc, cc = oracle.construct_linear_filter(T, bg, all_cs=True)
    np.testing.assert_allclose(cc[-1], c, rtol=1e-12)
    # single-window score == detect at one offset (+ A13), test_template.py:25-41
    tmpl, bgd = T.reshape(6, 12).T, bg.reshape(12)
    s, C = oracle.score_template_background_section(tmpl, bgd, E.reshape(6, 12).T)
    np.testing.assert_allclose(s + C, llr, rtol=1e-9)
    d = oracle.detect(E.astype(np.uint8), W)
    np.testing.assert_allclose(d[0], s, rtol=1e-5)


# -------------------------------------------------------------- Bernoulli mixture KAT
def test_bernoulli_mixture_docstring_kat(oracle):
    """bernoulli_mixture.py:50-71, the only numeric known-answer on the hot path."""
    data = np.array([[1, 1, 0], [0, 0, 1], [1, 1, 1]])
    m = oracle.BernoulliMixture(2, data)
    m.run_EM(1e-3)
    np.testing.assert_allclose(m.templates, [[.95, .95, .50010438], [.05, .05, .95]], atol=1e-5)
    np.testing.assert_allclose(m.weights, [.66671347, .33328653], atol=1e-5)
    np.testing.assert_allclose(m.affinities, [[9.99861515e-01, 1.38484719e-04], [2.90861524e-03, 9.97091385e-01],
                                              [9.97376426e-01, 2.62357439e-03]], atol=1e-5)


def test_bernoulli_mixture_planted(oracle):
    """tests.py:491-506 recipe: planted templates are recovered and the log-likelihood is monotone."""
    X, z, T, w = synth.em_data(N=600, D=36, K=3, seed=4)
    m = oracle.BernoulliMixture(3, X)
    m.run_EM(1e-6, min_probability=.01)
    tr = np.array(m.loglikelihood_trace)
    assert np.all(np.diff(tr) > -1e-6 * np.abs(tr[:-1]))
    # each planted template is matched by some estimated one
    err = np.abs(m.templates[:, None, :] - T[None]).mean(-1)
    assert err.min(0).max() < 0.08
    m2 = oracle.Bernoulli_Mixture(3, X.reshape(600, 6, 6))
    m2.run_EM(1e-4)
    assert np.isfinite(m2.loglikelihood) and m2.templates.min() >= .05 and m2.templates.max() <= .95


def test_synth_generators_are_seeded():
    a, b = synth.utterance_edges(3), synth.utterance_edges(3)
    np.testing.assert_array_equal(a, b)
    assert a.shape == (300, 40, 8) and a.dtype == np.uint8 and 0.03 < a.mean() < 0.2
    assert synth.fast_edge_batch(4).shape == (4, 300, 40, 8)
    lp, lip = synth.parts(5)
    assert lp.dtype == np.float64 and np.allclose(np.exp(lp) + np.exp(lip), 1)


# ------------------------------------------------------------------ edge front end (N4, N3) and edge spreading
GOLD_EDGE_CASES = [(q, s, kind) for q in (0.7, 0.3) for s, kind in ((0, "unidirectional"), (1, "unidirectional"),
                                                                     (3, "unidirectional"), (2, "all"))]


@pytest.fixture(scope="module")
def golden_edges():
    return dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_edges.npz")))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_edge_front_end_golden(oracle, golden_edges, tag):
    """The restatement against outputs of the reference's OWN functions (tests/golden/make_golden_edges.py):
    edge extraction, block thresholding (full, exact and ragged/short last blocks), both spread types."""
    g = golden_edges
    E, breaks, ori = oracle._edge_map_no_threshold(g["S_" + tag])
    np.testing.assert_array_equal(E, g["E_" + tag])
    np.testing.assert_array_equal(breaks, g["breaks_" + tag])
    np.testing.assert_array_equal(ori, g["ori_" + tag])
    blk = int(g["blk_" + tag])
    for q, s, kind in GOLD_EDGE_CASES:
        Eb = g["E_" + tag].copy()
        oracle._edge_map_threshold_segments(Eb, blk, s, threshold=q, edge_orientations=ori,
                                            edge_feature_row_breaks=breaks, spread_type=kind)
        np.testing.assert_array_equal(Eb.astype(np.uint8), g["bin_%s_q%d_s%d_%s" % (tag, int(q * 10), s, kind[:3])])
    _, taus = oracle.threshold_edgemap(g["E_" + tag][:, :blk].copy(), 0.7, breaks)
    np.testing.assert_array_equal(taus, g["taus_" + tag])


def test_spread_edgemap_golden(oracle, golden_edges):
    g = golden_edges
    for s in (1, 2, 3):
        T = g["spread_in"].astype(np.float64)
        oracle.spread_edgemap(T, g["spread_breaks"], g["ori_a"], s)
        np.testing.assert_array_equal(T.astype(np.uint8), g["spread_s%d" % s])


# ------------------------------------------------------------------ model-side code executed from the reference
@pytest.fixture(scope="module")
def golden_models():
    return dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_models.npz")))


def _em_kwargs(g, tag):
    kw = eval(str(g["bm_%s_kw" % tag]), {"__builtins__": {}}, {"dict": dict, "True": True, "False": False})
    return kw.pop("tol"), kw


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_bernoulli_mixture_golden(oracle, golden_models, tag):
    """run_EM of the restatement against run_EM of the reference class itself (tests/golden/make_golden_models.py):
    soft and hard assignment, N not a multiple of K, 1-/2-/3-D data shapes, max_iter / min_iter stops."""
    g = golden_models
    tol, kw = _em_kwargs(g, tag)
    m = oracle.BernoulliMixture(int(g["bm_%s_K" % tag]), g["bm_%s_X" % tag])
    m.run_EM(tol, **kw)
    assert m.iterations == int(g["bm_%s_iterations" % tag])
    np.testing.assert_allclose(m.loglikelihood_trace, g["bm_%s_loglik" % tag], rtol=1e-12)
    np.testing.assert_allclose(m.templates, g["bm_%s_templates" % tag], rtol=1e-6)
    np.testing.assert_allclose(m.weights, g["bm_%s_weights" % tag], rtol=1e-6)
    np.testing.assert_allclose(m.affinities, g["bm_%s_affinities" % tag], atol=1e-7)


def test_bernoulli_mixture_preset_and_first_version_golden(oracle, golden_models):
    g = golden_models
    m = oracle.BernoulliMixture(3, g["bm_p_X"], init_type='preset', init_centers=g["bm_p_centers"])
    m.run_EM(1e-5, min_probability=.02)
    assert m.iterations == int(g["bm_p_iterations"])
    np.testing.assert_allclose(m.loglikelihood_trace, g["bm_p_loglik"], rtol=1e-12)
    np.testing.assert_allclose(m.templates, g["bm_p_templates"], rtol=1e-6)
    np.testing.assert_allclose(m.affinities, g["bm_p_affinities"], atol=1e-7)
    m = oracle.Bernoulli_Mixture(2, g["be_X"])
    m.run_EM(1e-4)
    np.testing.assert_allclose(m.templates, g["be_templates"], rtol=1e-12)
    np.testing.assert_allclose(m.weights, g["be_weights"], rtol=1e-12)
    np.testing.assert_allclose(m.affinities, g["be_affinities"], atol=1e-12)


def test_linear_filter_and_window_score_golden(oracle, golden_models):
    g = golden_models
    W, c = oracle.construct_linear_filter(g["lf_T"], g["lf_bgd"])
    assert W.dtype == np.float32
    np.testing.assert_array_equal(W, g["lf_W"])
    np.testing.assert_allclose(c, g["lf_c"], rtol=1e-14)
    W, c = oracle.construct_linear_filter(g["lf_T"], g["lf_bgd"], all_cs=True)
    np.testing.assert_array_equal(W, g["lf_W_allcs"])
    np.testing.assert_allclose(c, g["lf_c_allcs"], rtol=1e-14)
    fs = oracle.construct_linear_filters((g["lf_T"], g["lf_T"][::-1] ** 2), g["lf_bgd"])
    np.testing.assert_array_equal(fs[0][0], g["lfs_W0"])
    np.testing.assert_array_equal(fs[1][0], g["lfs_W1"])
    np.testing.assert_allclose([fs[0][1], fs[1][1]], [g["lfs_c0"], g["lfs_c1"]], rtol=1e-14)
    s, C = oracle.score_template_background_section(g["sc_template"], g["sc_bgd"], g["sc_E"])
    np.testing.assert_allclose([s, C], [g["sc_s"], g["sc_C"]], rtol=1e-14)


# ------------------------------------------------------------------ clustering of detection peaks
def test_cluster_times_golden_and_reference(oracle, golden):
    """cluster_times.pyx:133-219: the C restatement against the fixtures made by the compiled reference and, when
    oracle/_ref is built here, against the compiled reference on random peak trains."""
    for i in range(8):
        t, (C0, C1) = golden["ct_times_%d" % i], golden["ct_C_%d" % i]
        np.testing.assert_array_equal(oracle.cluster_times(t, float(C0), float(C1)), golden["ct_part_%d" % i])
    ct = oracle.load_ref_cluster_times()
    if ct is None:
        pytest.skip("compiled reference cluster_times not built here")
    rs = np.random.RandomState(3)
    for _ in range(400):
        t = np.cumsum(rs.choice([1, 2, 3, 5, 8, 13, 30, 60], size=rs.randint(0, 50))).astype(np.uint16)
        C0, C1 = float(rs.choice([2, 3, 5, 10, 33])), float(rs.choice([5, 10, 20, 33, 50, 100]))
        np.testing.assert_array_equal(oracle.cluster_times(t, C0, C1), ct.cluster_times(t, C0, C1))
    part = oracle.cluster_times(np.array([3, 4, 9, 30, 31], np.uint16), 5., 20.)
    assert oracle.get_cluster_starts_ends(np.array([3, 4, 9, 30, 31]), part) == ((3, 5), (9, 10), (30, 32))
