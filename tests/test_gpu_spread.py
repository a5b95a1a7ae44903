"""K0/K1a/K1b parity on the GPU through the C ABI: bit-exact against the oracle and the golden
fixtures (integer / bit work)."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

ORI = np.array([[0, 1], [0, -1], [1, 0], [-1, 0], [1, 1], [-1, -1], [-1, 1], [1, -1]])


def _oracle_spread(oracle, X, radius, ori, block_len=0, box=False):
    """X (T,F,8) u8 -> spread (T,F,8) via the oracle's 2-D (rows,time) functions."""
    T, F, _ = X.shape
    out = np.zeros_like(X)
    blocks = [(0, T)] if block_len == 0 else [(s, min(T, s + block_len)) for s in range(0, T, block_len)]
    for e in range(8):
        for s, t in blocks:
            blk = X[s:t, :, e].T.astype(np.float64)          # (frequency rows, time)
            if box:
                E2 = blk.copy()
                oracle.spread_edgemap_all(E2, [0, F], radius)
                out[s:t, :, e] = E2.T
            else:
                out[s:t, :, e] = oracle._spread_edges(blk, ori[e], radius).T
    return out


def test_pack_unpack_roundtrip(cuda):
    from tsr_b200 import ops
    rs = np.random.RandomState(0)
    for E in (8, 5, 1):
        X = (rs.rand(3, 17, 9, E) < 0.3).astype(np.uint8) * rs.randint(1, 255, (3, 17, 9, E)).astype(np.uint8)
        bits = ops.pack_edges(torch.from_numpy(X).cuda())
        ref = np.zeros((3, 17, 9), np.uint8)
        for e in range(E):
            ref |= ((X[..., e] != 0).astype(np.uint8) << e)
        np.testing.assert_array_equal(bits.cpu().numpy(), ref)
        np.testing.assert_array_equal(ops.unpack_edges(bits, E).cpu().numpy(), (X != 0).astype(np.uint8))
    for D in (1, 31, 64, 65, 3600):
        Fm = (rs.rand(2, 5, D) < 0.2).astype(np.uint8) * 7
        fb = ops.pack_features(torch.from_numpy(Fm).cuda())
        assert fb.shape[2] == ops.rowwords(D) and fb.shape[2] % 2 == 0
        np.testing.assert_array_equal(ops.unpack_features(fb, D).cpu().numpy(), (Fm != 0).astype(np.uint8))
        words = fb.cpu().numpy().view(np.uint32)
        assert int(sum(bin(int(w)).count("1") for w in words.ravel())) == int((Fm != 0).sum())   # padding bits are zero


@pytest.mark.parametrize("radius,block_len,box", [(1, 0, False), (3, 0, False), (1, 40, False), (2, 7, False),
                                                   (1, 0, True), (3, 40, True), (0, 0, False)])
def test_spread_edges_bit_exact(cuda, oracle, radius, block_len, box):
    from tsr_b200 import ops, synth
    X = np.stack([synth.utterance_edges(u, T=95, F=40) for u in range(3)])
    bits = ops.pack_edges(torch.from_numpy(X).cuda())
    out = ops.spread_edges(bits, radius, ORI, "box" if box else "line", block_len)
    got = ops.unpack_edges(out, 8).cpu().numpy()
    for b in range(3):
        np.testing.assert_array_equal(got[b], _oracle_spread(oracle, X[b], radius, ORI, block_len, box))


@pytest.mark.parametrize("T,F,radius,block_len,box", [(95, 40, 1, 40, False), (300, 40, 1, 40, False), (95, 40, 3, 40, False),
                                                      (64, 37, 2, 25, False), (33, 40, 1, 40, False), (95, 40, 2, 40, True),
                                                      (120, 40, 1, 40, False), (51, 9, 3, 10, True)])
def test_spread_edges_ragged_tail_window(cuda, oracle, T, F, radius, block_len, box):
    """tail_window: the frames past the last whole block are the tail of the spread window of the final block_len frames
    (_edge_map_threshold_segments, edge_signal_proc.py:1595-1633) -- one launch, equal to the two-call composition
    ops.spread_edge_blocks and to the NumPy restatement."""
    from tsr_b200 import ops, synth
    X = np.stack([synth.utterance_edges(40 + u, T=T, F=F) for u in range(3)])
    bits = ops.pack_edges(torch.from_numpy(X).cuda())
    mode = "box" if box else "line"
    out = ops.spread_edges(bits, radius, ORI, mode, block_len, tail_window=True)
    tail = bits[:, max(0, T - block_len):].contiguous() if T % block_len else None
    two = ops.spread_edge_blocks(bits, tail, radius, block_len, ORI, mode)
    assert torch.equal(out, two)
    got = ops.unpack_edges(out, 8).cpu().numpy()
    for b in range(3):
        want = _oracle_spread(oracle, X[b], radius, ORI, block_len, box)
        last = T % block_len
        if last:
            want[-last:] = _oracle_spread(oracle, X[b, -block_len:], radius, ORI, 0, box)[-last:]
            if not box:
                np.testing.assert_array_equal(want, oracle.spread_edge_blocks(X[b], radius, block_len, ORI))
        np.testing.assert_array_equal(got[b], want)


def test_spread_edges_odd_shapes_and_other_orientations(cuda, oracle):
    from tsr_b200 import ops
    rs = np.random.RandomState(1)
    ori2 = np.array([[1, 0], [-1, 0], [0, 1], [0, -1], [1, 1], [-1, -1], [1, -1], [-1, 1]])   # edge_signal_proc.py:1541-1548
    for (T, F) in ((1, 1), (5, 3), (33, 45), (64, 7)):
        X = (rs.rand(2, T, F, 8) < 0.1).astype(np.uint8)
        out = ops.unpack_edges(ops.spread_edges(ops.pack_edges(torch.from_numpy(X).cuda()), 2, ori2), 8).cpu().numpy()
        for b in range(2):
            np.testing.assert_array_equal(out[b], _oracle_spread(oracle, X[b], 2, ori2))


def test_spread_edgemap_dropin_ragged_blocks(cuda, oracle):
    """edge_signal_proc.spread_edgemap on the reference's (rows, time) float layout, with the
    ragged row breaks compute_edge_feature_row_breaks produces (edge_signal_proc.py:1676-1691)."""
    from tsr_b200 import edge_signal_proc as esp
    rs = np.random.RandomState(2)
    ori = ORI.astype(float)
    heights = 10 + 3 * (ori[:, 0] == 0)
    breaks = np.concatenate([[0], np.cumsum(heights)]).astype(float)
    T = (rs.rand(int(breaks[-1]), 37) < 0.12).astype(np.float64)
    want = T.copy()
    oracle.spread_edgemap(want, breaks, ori.astype(int), 3)
    got = T.copy()
    assert esp.spread_edgemap(got, breaks, ori, spread_length=3) is None
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(esp._spread_edges(T[:13], (0, 1), 2), oracle._spread_edges(T[:13], (0, 1), 2))
    want = T.copy(); oracle.spread_edgemap_all(want, breaks, 2)
    got = T.copy(); esp.spread_edgemap_all(got, breaks, 2)
    np.testing.assert_array_equal(got, want)
    with pytest.raises(ValueError):
        esp.spread_edgemap(T * 0.5, breaks, ori)


def test_spread_codes_bit_exact(cuda, oracle, golden):
    from tsr_b200 import code_parts as cp, spread_waliji_patches as swp
    g = golden
    np.testing.assert_array_equal(cp.spread_patches(g["cp_codes"], 2, 2, 7), g["sp_22"])
    np.testing.assert_array_equal(swp.spread_waliji_patches(g["cp_codes"], 1, 3, 7), g["sp_13"])
    np.testing.assert_array_equal(cp.spread_patches(g["cp_codes"], 0, 0, 7), g["sp_00"])
    rs = np.random.RandomState(3)
    for (D0, D1, K, r0, r1) in ((296, 36, 100, 2, 2), (296, 36, 100, 5, 5), (50, 9, 200, 1, 0), (7, 3, 1, 9, 9),
                                (40, 36, 33, 0, 3), (12, 5, 255, 2, 2), (1, 1, 5, 2, 2)):
        codes = rs.randint(0, K + 1, (D0, D1)).astype(np.int64)
        codes[rs.rand(D0, D1) < 0.5] = 0
        got = cp.spread_patches(codes, r0, r1, K)
        assert got.dtype == np.uint8 and got.shape == (D0, D1, K)
        np.testing.assert_array_equal(got, oracle.spread_patches(codes, r0, r1, K))
    assert cp.spread_patches(np.zeros((0, 4), np.int64), 1, 1, 3).shape == (0, 4, 3)


def test_spread_graph_bit_exact(cuda, oracle, golden):
    from tsr_b200 import code_parts as cp
    S = golden["sp_00"].copy()
    assert cp.spread_graph_patches(S, golden["graph"]) is None
    np.testing.assert_array_equal(S, golden["sp_graph"])
    rs = np.random.RandomState(4)
    for K in (5, 32, 100, 200):
        X = (rs.rand(9, 11, K) < 0.05).astype(np.uint8)
        G = (rs.rand(K, K) < 2.0 / K).astype(np.uint8)
        want = X.copy(); oracle.spread_graph_patches(want, G)
        got = X.copy(); cp.spread_graph_patches(got, G)
        np.testing.assert_array_equal(got, want)
