"""End-to-end parity of the batched pipeline (spread -> code_parts -> spread_patches -> mixture
LLR detection) against the oracle chained exactly as get_train_data.py:2576-2631 and :927-1011 do."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def test_pipeline_matches_oracle_chain(cuda, oracle):
    from tsr_b200 import ops, synth
    from tsr_b200.pipeline import ScoringPipeline
    B, M = 4, 6
    lp, lip = synth.parts(100)
    T, bgd = synth.templates(M=M, L=40, Fd=36, Kd=100)
    W, c = synth.linear_filters(T, bgd)
    X = synth.utterance_batch(B)
    pipe = ScoringPipeline(lp, lip, W, c, edge_threshold=12, part_spread=(2, 2), edge_spread_radius=1, edge_block_len=40)
    bits = pipe.from_numpy(X)
    best, arg, codes = pipe(bits)
    torch.cuda.synchronize()
    best, arg, codes = best.cpu().numpy(), arg.cpu().numpy().view(np.uint16), codes.cpu().numpy()
    for b in range(B):
        # K1a: per 40-frame block, radius 1; the ragged last 20 frames are the tail of the window of the final 40
        # (edge_signal_proc.py:1595-1663)
        Xs = oracle.spread_edge_blocks(X[b], 1, 40)
        out = oracle.code_parts(Xs, lp, lip, 12)
        ref_codes = np.argmax(out, -1)
        mism = ref_codes != codes[b]
        assert mism.sum() <= 2
        S = oracle.spread_patches(codes[b].astype(np.int64), 2, 2, 100)       # continue from the GPU's codes
        v = np.stack([oracle.detect(S, W[m]).astype(np.float64) + c[m] for m in range(M)], -1)
        want = v.max(-1)
        np.testing.assert_allclose(best[b], want, atol=1e-5 * np.abs(want).max())
        top2 = np.sort(v, -1)
        clear = (top2[:, -1] - top2[:, -2]) > 1e-4 * np.abs(top2[:, -1])
        np.testing.assert_array_equal(arg[b][clear], v.argmax(-1)[clear])


def test_pipeline_with_a_large_dictionary_takes_the_int8_coding_form(cuda, oracle):
    """More than 126 parts: ScoringPipeline codes with TSR_IMPL_TENSOR_INT8 (its stream-ordered work list is allocated
    and freed on the pipeline's stream, also inside stream()'s multi-stream schedule); the codes equal the oracle's
    arg-max at every location and stream() returns what direct calls return."""
    from tsr_b200 import ops, synth
    from tsr_b200.pipeline import ScoringPipeline
    K, M = 140, 3
    lp, lip = synth.parts(K)
    T, bgd = synth.templates(M=M, L=10, Fd=36, Kd=K)
    W, c = synth.linear_filters(T, bgd)
    pipe = ScoringPipeline(lp, lip, W, c, edge_threshold=12, part_spread=(2, 2), edge_spread_radius=1, edge_block_len=40)
    assert pipe.impl_code == ops.IMPL_TENSOR_INT8
    batches = [synth.fast_edge_batch(b, T=100, seed=10 + s) for s, b in enumerate([3, 3, 6, 2])]
    want = []
    for X in batches:
        best, arg, codes = pipe(pipe.from_numpy(X))
        torch.cuda.synchronize()
        want.append((best.cpu().numpy().copy(), arg.cpu().numpy().copy()))
        for b in range(X.shape[0]):
            ref = oracle.code_parts(oracle.spread_edge_blocks(X[b], 1, 40), lp, lip, 12)
            np.testing.assert_array_equal(codes[b].cpu().numpy(), np.argmax(ref, -1))
    n = 0
    for (best, arg), (wb, wa) in zip(pipe.stream(iter(batches)), want):
        np.testing.assert_array_equal(best.numpy(), wb)
        np.testing.assert_array_equal(arg.numpy(), wa)
        n += 1
    assert n == len(batches)


def test_stream_equals_direct_calls(cuda):
    """ScoringPipeline.stream (double-buffered H2D / D2H on side streams) returns, batch by batch, exactly what a
    direct call on the same batch returns -- pinned tensors and NumPy arrays in, changing batch sizes, ring reuse."""
    from tsr_b200 import synth
    from tsr_b200.pipeline import ScoringPipeline
    lp, lip = synth.parts(30)
    T, bgd = synth.templates(M=3, L=10, Fd=36, Kd=30)
    W, c = synth.linear_filters(T, bgd)
    pipe = ScoringPipeline(lp, lip, W, c)
    batches = [synth.fast_edge_batch(b, T=120, seed=s) for s, b in enumerate([5, 5, 3, 8, 8, 8, 1])]
    want = []
    for X in batches:
        best, arg, _ = pipe(pipe.from_numpy(X))
        want.append((best.cpu().numpy().copy(), arg.cpu().numpy().copy()))
    mixed = [torch.from_numpy(X).pin_memory() if i % 2 == 0 else X for i, X in enumerate(batches)]
    n = 0
    for (best, arg), (wb, wa) in zip(pipe.stream(iter(mixed)), want):
        np.testing.assert_array_equal(best.numpy(), wb)
        np.testing.assert_array_equal(arg.numpy(), wa)
        n += 1
    assert n == len(batches)
    assert list(pipe.stream(iter([]))) == []
    # pre-packed edge bytes in, and the unpacked tensor packed on the device: same results
    packed = [np.packbits(X != 0, axis=-1, bitorder="little")[..., 0] for X in batches]
    for kw, feed in ((dict(), [torch.from_numpy(P).pin_memory() if i % 2 else P for i, P in enumerate(packed)]),
                     (dict(pack="device"), mixed), (dict(pack_threads=1), batches)):
        n = 0
        for (best, arg), (wb, wa) in zip(pipe.stream(iter(feed), **kw), want):
            np.testing.assert_array_equal(best.numpy(), wb)
            np.testing.assert_array_equal(arg.numpy(), wa)
            n += 1
        assert n == len(batches)
    with pytest.raises(ValueError):
        list(pipe.stream(iter([np.zeros((2, 120, 40, 8), np.float32)])))
    # abandoning the generator early stops the feeder thread
    it = pipe.stream(iter(batches))
    next(it)
    it.close()


_FRESH = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
from tsr_b200 import synth
from tsr_b200.pipeline import ScoringPipeline
lp, lip = synth.parts(100)
T, bgd = synth.templates(M=8, L=40, Fd=36, Kd=100)
W, c = synth.linear_filters(T, bgd)
pipe = ScoringPipeline(lp, lip, W, c)
batches = [synth.fast_edge_batch(b, seed=s) for s, b in enumerate([640, 640, 512, 640])]
feed = [torch.from_numpy(b).pin_memory() for b in batches] if sys.argv[1] == "split" else batches     # split needs pinned tensors
got = [(b.numpy().copy(), a.numpy().copy()) for b, a in pipe.stream(iter(feed), pack=sys.argv[1])]   # FIRST work of the process
torch.cuda.synchronize()
for X, (gb, ga) in zip(batches, got):
    best, arg, _ = pipe(pipe.from_numpy(X))
    assert np.array_equal(best.cpu().numpy(), gb) and np.array_equal(arg.cpu().numpy(), ga)
print("fresh-process stream ok")
"""


@pytest.mark.parametrize("pack", ["host", "device", "split"])
def test_stream_large_batches_in_a_fresh_process(cuda, pack):
    """Cross-stream hazard check (device input buffers are handed out by the caching allocator while kernels of the
    previous batch may still be running on the main stream): bench-sized batches, allocator cache cold, stream() is
    the first thing the process does; results must equal direct calls."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _FRESH % root, pack], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "fresh-process stream ok" in r.stdout, r.stdout + r.stderr


def test_stream_split_packing_equals_direct_calls(cuda):
    """pack="split": the host packs the first rows of every batch, the GPU the rest, and the split point moves from batch
    to batch with the measured rates -- whatever it is, every batch must score exactly like a direct call."""
    from tsr_b200 import synth
    from tsr_b200.pipeline import ScoringPipeline
    lp, lip = synth.parts(30)
    T, bgd = synth.templates(M=3, L=10, Fd=36, Kd=30)
    W, c = synth.linear_filters(T, bgd)
    pipe = ScoringPipeline(lp, lip, W, c)
    batches = [synth.fast_edge_batch(b, T=120, seed=s) for s, b in enumerate([64, 64, 64, 40, 40, 64, 64, 8])]
    want = []
    for X in batches:
        best, arg, _ = pipe(pipe.from_numpy(X))
        want.append((best.cpu().numpy().copy(), arg.cpu().numpy().copy()))
    pinned = [torch.from_numpy(X).pin_memory() for X in batches]
    n = 0
    for (best, arg), (wb, wa) in zip(pipe.stream(iter(pinned), pack="split", pack_threads=2), want):
        np.testing.assert_array_equal(best.numpy(), wb)
        np.testing.assert_array_equal(arg.numpy(), wa)
        n += 1
    assert n == len(batches)
    assert pipe.stream_stats.get("split_rows") is not None
