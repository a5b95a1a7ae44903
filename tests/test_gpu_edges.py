"""N4 / N3 parity on the GPU through the C ABI: 8-orientation edge extraction and per-block quantile
thresholding, bit-exact (float64 differences, comparisons and an exact order statistic) against the golden
fixtures made by executing the reference's own functions (tests/golden/make_golden_edges.py) and against
the oracle on seeded inputs, including the reference-signature drop-ins."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

CASES = [(q, s, kind) for q in (0.7, 0.3) for s, kind in ((0, "unidirectional"), (1, "unidirectional"),
                                                           (3, "unidirectional"), (2, "all"))]


@pytest.fixture(scope="module")
def golden_edges():
    return dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_edges.npz")))


def _planes(bits, F):
    """packed bytes (T,F) -> the reference's (8F, T) 0/1 array"""
    b = bits.cpu().numpy()
    return np.concatenate([((b >> e) & 1).T for e in range(8)], 0).astype(np.uint8)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_golden_drop_in(cuda, golden_edges, tag):
    from tsr_b200 import edge_signal_proc as esp
    g = golden_edges
    E, breaks, ori = esp._edge_map_no_threshold(g["S_" + tag])
    np.testing.assert_array_equal(E, g["E_" + tag])
    np.testing.assert_array_equal(breaks, g["breaks_" + tag])
    np.testing.assert_array_equal(ori, g["ori_" + tag])
    blk = int(g["blk_" + tag])
    for q, s, kind in CASES:
        Eb = g["E_" + tag].copy()
        esp._edge_map_threshold_segments(Eb, blk, s, threshold=q, edge_orientations=ori, edge_feature_row_breaks=breaks,
                                         spread_type=kind)
        np.testing.assert_array_equal(Eb.astype(np.uint8), g["bin_%s_q%d_s%d_%s" % (tag, int(q * 10), s, kind[:3])])
    Eb = g["E_" + tag][:, :blk].copy()
    _, taus = esp.threshold_edgemap(Eb, 0.7, breaks, report_level=True)
    np.testing.assert_array_equal(taus, g["taus_" + tag])
    # one block through threshold_edge_block, keeping magnitudes
    H = g["S_" + tag].shape[0]
    blk0 = g["E_" + tag][2 * H:3 * H, :blk].copy()
    kept, tau = esp.threshold_edge_block(blk0.copy(), 0.7, True, 0.015, convert_to_binary=False)
    assert tau == g["taus_" + tag][2]
    np.testing.assert_array_equal(kept != 0, Eb[2 * H:3 * H] != 0)
    np.testing.assert_array_equal(kept[kept != 0], blk0[kept != 0])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_golden_fused(cuda, golden_edges, tag):
    """spectrogram -> packed edge bytes in one kernel (+ block spreading) equals the reference chain"""
    from tsr_b200 import ops
    g = golden_edges
    S, blk = g["S_" + tag], int(g["blk_" + tag])
    F = S.shape[0]
    Sd = torch.from_numpy(S[None]).cuda()
    for q, s, kind in CASES:
        bits, tail, taus = ops.edges_from_spectrogram(Sd, blk, q, want_taus=True)
        out = ops.spread_edge_blocks(bits, tail, s, blk, None, "line" if kind == "unidirectional" else "box")
        np.testing.assert_array_equal(_planes(out[0], F), g["bin_%s_q%d_s%d_%s" % (tag, int(q * 10), s, kind[:3])])
        if q == 0.7:
            np.testing.assert_array_equal(taus[0, 0].cpu().numpy(), g["taus_" + tag])


@pytest.mark.parametrize("F,T,blk", [(40, 300, 40), (40, 317, 40), (7, 23, 5), (64, 100, 100), (4, 4, 40), (33, 129, 64)])
def test_batched_against_oracle(cuda, oracle, F, T, blk):
    from tsr_b200 import ops
    rs = np.random.RandomState(F * 1000 + T)
    B = 6
    S = rs.randn(B, F, T).cumsum(2) * 0.03 + rs.randn(B, F, T) * 0.02
    S[1] = np.round(S[1] * 8) / 8            # many exact ties
    S[2, :, : T // 2] = 0.5                  # half the utterance flat: empty blocks
    if B > 3:
        S[3] = 0.0                           # no edge anywhere
    S[4] = 0.5 * np.arange(T)[None, :] + 0.25 * np.arange(F)[:, None]   # ramp: every cell is a candidate (list overflow path)
    Sd = torch.from_numpy(S).cuda()
    mags = ops.edge_map(Sd)
    bits_f, tail_f, taus_f = ops.edges_from_spectrogram(Sd, blk, 0.7, want_taus=True)
    bits_m, tail_m, taus_m = ops.threshold_edges(mags, blk, 0.7, want_taus=True)
    assert torch.equal(bits_f, bits_m) and torch.equal(taus_f, taus_m)
    assert (tail_f is None) == (T % blk == 0) and (tail_f is None or torch.equal(tail_f, tail_m))
    for r, mode, kind in ((1, "line", "unidirectional"), (2, "box", "all")):
        out = ops.spread_edge_blocks(bits_f, tail_f, r, blk, None, mode)
        for b in range(B):
            E, breaks, ori = oracle._edge_map_no_threshold(S[b])
            np.testing.assert_array_equal(mags[b].cpu().numpy().reshape(8 * F, T), E)
            oracle._edge_map_threshold_segments(E, blk, r, threshold=0.7, edge_orientations=ori,
                                                edge_feature_row_breaks=breaks, spread_type=kind)
            np.testing.assert_array_equal(_planes(out[b], F), E.astype(np.uint8))


def test_quantile_sweep_and_errors(cuda, oracle):
    from tsr_b200 import ops
    rs = np.random.RandomState(5)
    S = rs.randn(1, 12, 80) * 0.05
    Sd = torch.from_numpy(S).cuda()
    E0, breaks, _ = oracle._edge_map_no_threshold(S[0])
    for q in (0.0, 0.05, 0.5, 0.95, 0.999):
        bits, _, taus = ops.edges_from_spectrogram(Sd, 40, q, abst_threshold=np.zeros(8), want_taus=True)
        for blk in range(2):
            ref, tr = oracle.threshold_edgemap(E0[:, blk * 40:(blk + 1) * 40].copy(), q, breaks, np.zeros(8))
            np.testing.assert_array_equal(_planes(bits[0, blk * 40:(blk + 1) * 40], 12), ref.astype(np.uint8))
            np.testing.assert_array_equal(taus[0, blk].cpu().numpy(), tr)
    with pytest.raises(ValueError):
        ops.edges_from_spectrogram(Sd, 40, 1.0)
    with pytest.raises(ValueError):
        ops.edges_from_spectrogram(Sd.float(), 40, 0.7)
    with pytest.raises(ValueError):
        ops.edge_map(torch.zeros((1, 3, 10), dtype=torch.float64, device="cuda"))
    assert ops.edge_map(torch.zeros((0, 8, 10), dtype=torch.float64, device="cuda")).shape == (0, 8, 8, 10)


def test_front_end_feeds_coding(cuda, oracle):
    """spectrogram -> edges -> spread -> code_parts stays bit-exact end to end against the oracle chain"""
    from tsr_b200 import ops, synth
    rs = np.random.RandomState(11)
    F, T, K = 40, 120, 20
    S = rs.randn(2, F, T).cumsum(2) * 0.03 + rs.randn(2, F, T) * 0.02
    lp, lip = synth.parts(K)
    parts = ops.PartsModel(lp, lip)
    bits, tail, _ = ops.edges_from_spectrogram(torch.from_numpy(S).cuda(), 40, 0.7)
    spread = ops.spread_edge_blocks(bits, tail, 1, 40)
    codes = ops.code_parts(spread, parts, 12)[0].cpu().numpy()
    for b in range(2):
        E, breaks, ori = oracle._edge_map_no_threshold(S[b])
        oracle._edge_map_threshold_segments(E, 40, 1, edge_orientations=ori, edge_feature_row_breaks=breaks)
        X = np.ascontiguousarray(E.reshape(8, F, T).transpose(2, 1, 0)).astype(np.uint8)
        ref = np.argmax(oracle.code_parts(X, lp, lip, 12), -1)
        assert (codes[b] != ref).mean() < 1e-3          # near-ties only (tests/test_gpu_code_parts.py pins the margin)


def test_spread_edgemap_reference_golden(cuda, golden_edges):
    """edge spreading drop-in against the reference's own spread_edgemap (executed from source)"""
    from tsr_b200 import edge_signal_proc as esp
    g = golden_edges
    for s in (1, 2, 3):
        T = g["spread_in"].astype(np.float64)
        esp.spread_edgemap(T, g["spread_breaks"], g["ori_a"], s)
        np.testing.assert_array_equal(T.astype(np.uint8), g["spread_s%d" % s])
