"""K4 parity on the GPU: Bernoulli-mixture EM against the NumPy restatement of
bernoulli_mixture.py / bernoulli_em.py and the docstring known-answer (bernoulli_mixture.py:50-71)."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
RTOL = 1e-5


def test_docstring_kat(cuda):
    from tsr_b200.bernoulli_mixture import BernoulliMixture
    data = np.array([[1, 1, 0], [0, 0, 1], [1, 1, 1]])
    m = BernoulliMixture(2, data)
    m.run_EM(1e-3)
    np.testing.assert_allclose(m.templates, [[.95, .95, .50010438], [.05, .05, .95]], atol=1e-5)
    np.testing.assert_allclose(m.weights, [.66671347, .33328653], atol=1e-5)
    np.testing.assert_allclose(m.affinities, [[9.99861515e-01, 1.38484719e-04], [2.90861524e-03, 9.97091385e-01],
                                              [9.97376426e-01, 2.62357439e-03]], atol=1e-5)
    assert m.templates.shape == (2, 3) and m.affinities.dtype == np.float32


QMAX = 8355711


def _params(K, D, seed=1, lo=.05):
    rs = np.random.RandomState(seed)
    Tm = np.clip(rs.rand(K, D), lo, 1 - lo).astype(np.float32)
    wts = rs.dirichlet(np.ones(K)).astype(np.float32)
    return Tm, np.log(Tm), np.log(1 - Tm), np.log(wts)


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tensor"])
@pytest.mark.parametrize("N,D,K", [(700, 200, 7), (300, 1000, 16), (129, 33, 2), (513, 260, 43), (1000, 200, 100),
                                   (600, 200, 200), (257, 4100, 3)])
def test_estep_matches_numpy(cuda, N, D, K, impl):
    """One E-step + M accumulation against float64 NumPy: scores, log-likelihood, posteriors, and the EXACT integer
    statistics.  K = 43 / 100 / 200 tile the components over several UMMA M tiles (the reference trains 100-200
    component part dictionaries, Notebook/12/reestimate_parts.py:79-83)."""
    from tsr_b200 import ops, synth
    X, _, T, w = synth.em_data(N=N, D=D, K=min(K, 16), seed=N)
    Tm, logT, logI, logw = _params(K, D)
    fb = ops.pack_features(torch.from_numpy(X[None]).cuda())[0]
    st = ops.EmState(fb, D, K)
    scores = torch.empty((N, K), dtype=torch.float32, device="cuda")
    st.estep(torch.from_numpy(logT).cuda(), torch.from_numpy(logI).cuda(), torch.from_numpy(logw).cuda(),
             impl=impl, scores_out=scores)
    st.mstep_accumulate(impl=impl)
    torch.cuda.synchronize()
    Xd = X.astype(np.float64)
    S = Xd @ logT.astype(np.float64).T + (1 - Xd) @ logI.astype(np.float64).T
    # tensor path: 24-bit fixed-point weights, exact integer sums, one float32 rounding; SIMT: sequential float32 adds
    np.testing.assert_allclose(scores.cpu().numpy(), S, rtol=2e-6 if impl == 0 else 1e-5)
    ll = S + logw[None].astype(np.float64)
    mx = ll.max(1, keepdims=True)
    p = np.exp(ll - mx)
    aff = p / p.sum(1, keepdims=True)
    np.testing.assert_allclose(st.loglik(), (np.log(p.sum(1)) + mx[:, 0]).sum(), rtol=1e-6)
    got_aff = st.aff.cpu().numpy()
    np.testing.assert_allclose(got_aff, aff, atol=1e-5 if impl == 0 else 2e-4)
    np.testing.assert_allclose(got_aff.sum(1), 1.0, atol=1e-5)
    # statistics are exact integer sums of the posteriors quantised per component: q = rint(aff32 / s_k * QMAX) with
    # s_k the smallest power of two >= max_n aff32[n][k]; one float64 scaling at the end
    sk = np.exp2(np.ceil(np.log2(np.maximum(got_aff.max(0).astype(np.float64), 1e-300)))).astype(np.float32)
    q = np.rint((got_aff * (np.float32(1) / sk)[None]) * np.float32(QMAX)).astype(np.int64)
    stats = st.stats.cpu().numpy()
    np.testing.assert_array_equal(stats[K * D:K * D + K], q.sum(0).astype(np.float64) / QMAX * sk.astype(np.float64))
    qs = q.T @ X.astype(np.int64)                      # em_pack_kernel: integer quotient + remainder term (exact for multiples of QMAX)
    np.testing.assert_array_equal(stats[:K * D].reshape(K, D),
                                  ((qs // QMAX).astype(np.float64) + (qs % QMAX).astype(np.float64) * (1.0 / QMAX)) * sk.astype(np.float64)[:, None])
    np.testing.assert_allclose(st.numerators().cpu().numpy(), aff.T @ Xd, rtol=1e-5, atol=1e-4 if impl == 0 else 2e-2)
    np.testing.assert_allclose(st.resp_sums().cpu().numpy(), aff.sum(0), rtol=1e-5 if impl == 0 else 1e-3)
    # the accumulators were cleared behind the reads: a second pass gives the same bits
    before = st.stats.clone()
    st.estep(torch.from_numpy(logT).cuda(), torch.from_numpy(logI).cuda(), torch.from_numpy(logw).cuda(), impl=impl)
    st.mstep_accumulate(impl=impl)
    assert torch.equal(before, st.stats)


@pytest.mark.parametrize("K,D", [(5, 200), (100, 200), (16, 1300)])
def test_mstep_from_identical_affinities(cuda, oracle, K, D):
    """M-step templates / weights from the SAME posteriors as the NumPy restatement's M_step (:177-193): 1e-5."""
    from tsr_b200 import ops, synth
    N = 2000
    X, _, _, _ = synth.em_data(N=N, D=D, K=min(K, 16), seed=4)
    rs = np.random.RandomState(2)
    aff = rs.dirichlet(np.ones(K) * .3, size=N).astype(np.float32)
    a = oracle.BernoulliMixture(K, X)
    a.affinities = aff.copy()
    a.min_probability = np.float32(.02)
    a.M_step()
    fb = ops.pack_features(torch.from_numpy(X[None]).cuda())[0]
    st = ops.EmState(fb, D, K)
    for impl in (0, 1):
        st.mstep_accumulate(impl=impl, aff=torch.from_numpy(aff).cuda())
        z = lambda *sh: torch.zeros(sh, dtype=torch.float32, device="cuda")  # noqa: E731
        par = dict(templates=z(K, D), logT=z(K, D), logInvT=z(K, D), weights=z(K), logw=z(K))
        ops.em_finalize(st.stats, N, D, K, .02, par["templates"], par["logT"], par["logInvT"], par["weights"], par["logw"])
        np.testing.assert_allclose(par["templates"].cpu().numpy(), a.work_templates, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(par["weights"].cpu().numpy(), a.weights, rtol=1e-5)
        np.testing.assert_allclose(par["logT"].cpu().numpy(), a.log_templates, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(par["logInvT"].cpu().numpy(), a.log_invtemplates, rtol=1e-5, atol=1e-6)


def test_vanishing_component_keeps_relative_precision(cuda, oracle):
    """A component whose total responsibility is 1e-20 of the mass still defines its template through the RATIO of
    its sums (bernoulli_mixture.py:179-181).  Posteriors are quantised below a per-component power-of-two scale, so
    such a component keeps 24 significant bits instead of rounding to zero (which would kill it for good)."""
    from tsr_b200 import ops, synth
    N, D, K = 1500, 96, 4
    X, _, _, _ = synth.em_data(N=N, D=D, K=3, seed=6)
    rs = np.random.RandomState(3)
    aff = rs.dirichlet(np.ones(K), size=N).astype(np.float32)
    aff[:, 2] *= np.float32(1e-20)                       # vanishing column
    aff[:, 3] = 0
    aff[17, 3] = np.float32(3e-31)                       # a single row with a tiny weight owns component 3
    a = oracle.BernoulliMixture(K, X)
    a.affinities = aff.copy()
    a.min_probability = np.float32(.05)
    a.M_step()
    st = ops.EmState(ops.pack_features(torch.from_numpy(X[None]).cuda())[0], D, K)
    for impl in (0, 1):
        st.mstep_accumulate(impl=impl, aff=torch.from_numpy(aff).cuda())
        z = lambda *sh: torch.zeros(sh, dtype=torch.float32, device="cuda")  # noqa: E731
        par = dict(templates=z(K, D), logT=z(K, D), logInvT=z(K, D), weights=z(K), logw=z(K))
        ops.em_finalize(st.stats, N, D, K, .05, par["templates"], par["logT"], par["logInvT"], par["weights"], par["logw"])
        np.testing.assert_allclose(par["templates"].cpu().numpy(), a.work_templates, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(par["weights"].cpu().numpy(), a.weights, rtol=1e-5)
        np.testing.assert_array_equal(par["templates"].cpu().numpy()[3], np.where(X[17] != 0, np.float32(.95), np.float32(.05)))


def test_template_loglikelihoods_is_a_pure_accessor(cuda, oracle):
    """get_template_loglikelihoods() (:287-291) must not disturb the EM state: calling it between the E- and the
    M-step leaves M_step() equal to the restatement's."""
    from tsr_b200 import synth
    from tsr_b200.bernoulli_mixture import BernoulliMixture
    from tsr_b200.bernoulli_em import Bernoulli_Mixture
    X, _, _, _ = synth.em_data(N=900, D=120, K=4, seed=8)
    a = oracle.BernoulliMixture(4, X); a.hard_assignment = False
    b = BernoulliMixture(4, X)
    la, lb = a._compute_loglikelihoods(), b._compute_loglikelihoods()
    np.testing.assert_allclose(lb, la, rtol=RTOL)
    sa, sb = a.get_template_loglikelihoods(), b.get_template_loglikelihoods()
    np.testing.assert_allclose(sb, sa, rtol=1e-5)
    a.M_step(); b.M_step()
    np.testing.assert_allclose(b.work_templates, a.work_templates, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(b.weights, a.weights, rtol=1e-4)
    np.testing.assert_allclose(b._compute_loglikelihoods(), a._compute_loglikelihoods(), rtol=RTOL)
    # first-version class: scoring other data than the training set
    Y = X[:50].reshape(50, 10, 12)
    e = Bernoulli_Mixture(3, X.reshape(900, 10, 12))
    s_train = e.get_template_loglikelihoods()
    s_other = e.get_template_loglikelihoods(Y)
    np.testing.assert_allclose(s_other, s_train[:50], rtol=1e-6)
    ll0 = e.compute_loglikelihoods()
    e.get_template_loglikelihoods(Y)
    e.M_step()
    assert np.isfinite(e.compute_loglikelihoods()) and e.compute_loglikelihoods() >= ll0 - 1e-6 * abs(ll0)


def test_em_matches_restatement(cuda, oracle):
    """Whole run_EM against the NumPy restatement: same iteration count, log-likelihood trace within
    1e-5 relative, templates and weights close."""
    from tsr_b200 import synth
    from tsr_b200.bernoulli_mixture import BernoulliMixture
    X, z, T, w = synth.em_data(N=3000, D=200, K=5, seed=21)
    a = oracle.BernoulliMixture(5, X)
    a.run_EM(1e-5, min_probability=.01)
    b = BernoulliMixture(5, X)
    b.run_EM(1e-5, min_probability=.01)
    assert abs(a.iterations - b.iterations) <= 1
    n = min(len(a.loglikelihood_trace), len(b.loglikelihood_trace))
    np.testing.assert_allclose(b.loglikelihood_trace[:n], a.loglikelihood_trace[:n], rtol=RTOL)
    assert len(b.loglikelihood_trace) == b.iterations + 1
    if a.iterations == b.iterations:
        np.testing.assert_allclose(b.templates, a.templates, atol=2e-4)
        np.testing.assert_allclose(b.weights, a.weights, atol=1e-4)
        np.testing.assert_allclose(b.affinities, a.affinities, atol=2e-3)
    assert b.templates.shape == (5, 200) and b.work_templates.dtype == np.float32
    # hard assignment variant (:271-275)
    c = oracle.BernoulliMixture(5, X); c.run_EM(1e-4, hard_assignment=True, max_iter=15)
    d = BernoulliMixture(5, X); d.run_EM(1e-4, hard_assignment=True, max_iter=15)
    np.testing.assert_allclose(d.loglikelihood_trace[0], c.loglikelihood_trace[0], rtol=RTOL)
    assert set(np.unique(d.affinities)) <= {0.0, 1.0}


@pytest.mark.parametrize("K", [100, 200])
def test_part_dictionary_em_many_components(cuda, oracle, K):
    """The reference's part-dictionary training shape -- 5x5x8 patches (D = 200), 100 / 200 components
    (Notebook/12/reestimate_parts.py:79-83, Development/090712/scripts/test_coding.py:104-107) -- on a 6 000-patch
    sample: log-likelihood trace within 1e-5 relative of the NumPy restatement."""
    from tsr_b200 import synth
    from tsr_b200.bernoulli_mixture import BernoulliMixture
    X, _, _, _ = synth.em_data(N=6000, D=200, K=40, seed=K)
    a = oracle.BernoulliMixture(K, X)
    a.run_EM(1e-6, min_probability=.01, max_iter=12)
    b = BernoulliMixture(K, X)
    b.run_EM(1e-6, min_probability=.01, max_iter=12)
    n = min(len(a.loglikelihood_trace), len(b.loglikelihood_trace))
    assert n >= 11
    np.testing.assert_allclose(b.loglikelihood_trace[:n], a.loglikelihood_trace[:n], rtol=RTOL)
    assert b.templates.shape == (K, 200) and b.affinities.shape == (6000, K)
    np.testing.assert_allclose(b.weights.sum(), 1.0, atol=1e-5)


def test_first_version_em(cuda, oracle):
    from tsr_b200 import synth
    from tsr_b200.bernoulli_em import Bernoulli_Mixture
    X, _, _, _ = synth.em_data(N=400, D=36, K=3, seed=5)
    a = oracle.Bernoulli_Mixture(3, X.reshape(400, 6, 6)); a.run_EM(1e-3)
    b = Bernoulli_Mixture(3, X.reshape(400, 6, 6)); b.run_EM(1e-3)
    np.testing.assert_allclose(b.loglikelihood, a.loglikelihood, rtol=1e-4)
    assert b.templates.shape == (3, 6, 6) and b.templates.dtype == np.float64
    assert b.templates.min() >= .05 - 1e-7 and b.templates.max() <= .95 + 1e-7


def test_config4_shape_against_restatement(cuda, oracle):
    """BASELINE config 4 shape (D = 12 800 = 40 frames x 40 bands x 8 edges, K = 16) on a 2 000-row
    sample: log-likelihood trace of the GPU EM against the NumPy restatement, 1e-5 relative."""
    from tsr_b200 import synth
    from tsr_b200.bernoulli_mixture import BernoulliMixture
    X, z, T, w = synth.em_data(N=2000, D=12800, K=16, seed=3)
    a = oracle.BernoulliMixture(16, X)
    a.run_EM(1e-6, min_probability=.05, max_iter=12)
    b = BernoulliMixture(16, X)
    b.run_EM(1e-6, min_probability=.05, max_iter=12)
    n = min(len(a.loglikelihood_trace), len(b.loglikelihood_trace))
    assert n >= 11
    np.testing.assert_allclose(b.loglikelihood_trace[:n], a.loglikelihood_trace[:n], rtol=RTOL)
    tr = np.array(b.loglikelihood_trace)
    assert np.all(np.diff(tr) > -1e-5 * np.abs(tr[:-1]))          # EM never decreases the likelihood
    np.testing.assert_allclose(b.weights, a.weights, atol=2e-3)


# ------------------------------------------------------------------ against the reference's own classes
def _golden_models():
    import os
    return dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_models.npz")))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_bernoulli_mixture_reference_golden(cuda, tag):
    """GPU run_EM against run_EM of the reference class executed from its source
    (tests/golden/make_golden_models.py): log-likelihood trace within 1e-5 relative, same stopping iteration."""
    from tsr_b200.bernoulli_mixture import BernoulliMixture
    g = _golden_models()
    kw = eval(str(g["bm_%s_kw" % tag]), {"__builtins__": {}}, {"dict": dict, "True": True, "False": False})
    tol = kw.pop("tol")
    m = BernoulliMixture(int(g["bm_%s_K" % tag]), g["bm_%s_X" % tag])
    m.run_EM(tol, **kw)
    ref = g["bm_%s_loglik" % tag]
    n = min(len(ref), len(m.loglikelihood_trace))
    np.testing.assert_allclose(m.loglikelihood_trace[:n], ref[:n], rtol=RTOL)
    assert abs(m.iterations - int(g["bm_%s_iterations" % tag])) <= 1
    assert m.templates.shape == g["bm_%s_templates" % tag].shape
    if m.iterations == int(g["bm_%s_iterations" % tag]):
        np.testing.assert_allclose(m.templates, g["bm_%s_templates" % tag], atol=2e-4)
        np.testing.assert_allclose(m.weights, g["bm_%s_weights" % tag], atol=1e-4)
        np.testing.assert_allclose(m.affinities, g["bm_%s_affinities" % tag], atol=2e-3)


def test_preset_first_version_filters_reference_golden(cuda):
    from tsr_b200 import estimate_template as et, test_template as tt
    from tsr_b200.bernoulli_em import Bernoulli_Mixture
    from tsr_b200.bernoulli_mixture import BernoulliMixture
    g = _golden_models()
    m = BernoulliMixture(3, g["bm_p_X"], init_type='preset', init_centers=g["bm_p_centers"])
    m.run_EM(1e-5, min_probability=.02)
    n = min(len(g["bm_p_loglik"]), len(m.loglikelihood_trace))
    np.testing.assert_allclose(m.loglikelihood_trace[:n], g["bm_p_loglik"][:n], rtol=RTOL)
    if m.iterations == int(g["bm_p_iterations"]):
        np.testing.assert_allclose(m.templates, g["bm_p_templates"], atol=2e-4)
    b = Bernoulli_Mixture(2, g["be_X"])
    b.run_EM(1e-4)
    np.testing.assert_allclose(b.loglikelihood, g["be_loglik"][-1], rtol=1e-4)
    np.testing.assert_allclose(b.templates, g["be_templates"], atol=1e-3)
    W, c = et.construct_linear_filter(g["lf_T"], g["lf_bgd"])
    np.testing.assert_array_equal(W, g["lf_W"])
    np.testing.assert_allclose(c, g["lf_c"], rtol=1e-12)
    W, c = et.construct_linear_filter(g["lf_T"], g["lf_bgd"], all_cs=True)
    np.testing.assert_allclose(c, g["lf_c_allcs"], rtol=1e-12)
    fs = et.construct_linear_filters((g["lf_T"], g["lf_T"][::-1] ** 2), g["lf_bgd"])
    np.testing.assert_array_equal(fs[1][0], g["lfs_W1"])
    s, C = tt.score_template_background_section(g["sc_template"], g["sc_bgd"], g["sc_E"])
    np.testing.assert_allclose([s, C], [g["sc_s"], g["sc_C"]], rtol=1e-5)


def test_peer_memory_allreduce_single_rank_group(cuda, oracle):
    """The library's own sum of the statistics over NVLink peer memory (tsr_em_allreduce_peer, ops.PeerStats) in a
    one-rank NCCL group on one GPU: the flag handshake, the two alternating source slots and the two-iteration CUDA
    graph all run (the only peer is the rank itself), and the result must equal the run without a group bit for bit.
    (Two ranks: tools/em_dist_check.py under torchrun; host logic over gloo: tests/test_em_driver_gloo.py.)"""
    import os
    import torch.distributed as dist
    from tsr_b200 import synth
    from tsr_b200.bernoulli_mixture import BernoulliMixture
    if dist.is_initialized():
        pytest.skip("a process group is already initialised in this process")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29577")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", torch.cuda.current_device()))
    try:
        X, _, _, _ = synth.em_data(N=3000, D=640, K=8, seed=5)
        a = BernoulliMixture(8, X)
        a.run_EM(1e-6, min_probability=.01, max_iter=14)
        b = BernoulliMixture(8, X, row_range=(0, 3000), group=dist.group.WORLD)
        if b._drv.peer is None:
            pytest.skip("symmetric memory unavailable here: %s" % b._drv.peer_error)
        b.run_EM(1e-6, min_probability=.01, max_iter=14)
        assert b._drv.peer.launches > 0
        assert b.iterations == a.iterations
        np.testing.assert_array_equal(np.array(b.loglikelihood_trace), np.array(a.loglikelihood_trace))
        np.testing.assert_array_equal(b.templates, a.templates)
        np.testing.assert_array_equal(b.affinities, a.affinities)
    finally:
        dist.destroy_process_group()
