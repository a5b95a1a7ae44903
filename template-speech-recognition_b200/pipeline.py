"""Batched, device-resident scoring pipeline: the form the benchmark and large jobs use.

  packed edges (B,T,F) --K1a spread_edges--> --K2 code_parts--> codes (B,To,Fo)
      --K1b spread_codes--> part feature bits (B,To,rowwords) --K3 detect--> scores (B,To,M)
      --reduce--> best / argbest (B,To)

It is get_train_data.get_part_features (:2576-2631) followed by _compute_detection_E
(:927-1011) for a whole batch of utterances per launch; the defaults are the reference's
(edge spread radius 1 per 40-frame block, main_multiprocessing.py:50-52; edge-count threshold
12 and part-spread radius (2,2), :71-73).
"""
import numpy as np
import torch

from . import ops


class ScoringPipeline:
    def __init__(self, log_parts, log_invparts, filters, cs, edge_threshold=12, part_spread=(2, 2),
                 edge_spread_radius=1, edge_block_len=40, edge_orientations=None, device="cuda",
                 detect_planes=ops.PLANES_INT8X3):
        self.parts = ops.PartsModel(log_parts, log_invparts, device=device)
        self.bank = ops.FilterBank(filters, cs=cs, device=device, planes=detect_planes)
        self.edge_threshold = int(edge_threshold)
        self.part_spread = (int(part_spread[0]), int(part_spread[1]))
        self.edge_spread_radius = int(edge_spread_radius)
        self.edge_block_len = int(edge_block_len)
        self.edge_orientations = edge_orientations
        self.impl_code = ops.IMPL_TENSOR
        self.impl_detect = ops.IMPL_TENSOR

    def code(self, bits):
        """K1a + K2 + K1b: packed edges -> (codes, part feature bits)."""
        if self.edge_spread_radius > 0:
            bits = ops.spread_edges(bits, self.edge_spread_radius, self.edge_orientations, "line", self.edge_block_len)
        codes, _, _ = ops.code_parts(bits, self.parts, self.edge_threshold, impl=self.impl_code)
        fb = ops.spread_codes(codes, self.parts.K, *self.part_spread)
        return codes, fb

    def score(self, fb):
        """K3 + reduction: part feature bits -> (best, argbest, per-template scores)."""
        Fo_K = fb.shape[2]
        if self.bank.D > Fo_K * 32:
            raise ValueError("filter width %d does not match the coded feature rows" % self.bank.D)
        scores = ops.detect(fb, self.bank, impl=self.impl_detect)
        best, arg = ops.detect_reduce(scores, self.bank.cs, self.bank.L, tail_mode=1)
        return best, arg, scores

    def __call__(self, bits):
        codes, fb = self.code(bits)
        best, arg, scores = self.score(fb)
        return best, arg, codes

    def from_numpy(self, X):
        """X uint8 (B,T,F,E) host array -> packed edge bytes on the device."""
        return ops.pack_edges(torch.from_numpy(np.ascontiguousarray(X)).to(self.parts.blob.device))

    def stream(self, host_batches, ring=3):
        """Score a sequence of host batches with the copies hidden behind the kernels.

        host_batches: iterable of uint8 (B,T,F,E) edge tensors in host memory (pinned torch tensors copy
        asynchronously; NumPy arrays are accepted and staged through a pinned buffer).  Yields, per batch and in
        order, ``(best float32 (B,To), argbest int16 (B,To))`` as pinned host tensors that stay valid until
        ``ring - 1`` further batches have been requested.

        Three streams: the H2D copy of batch i+1 (into the other of two device buffers) and the D2H copy of
        batch i-1's results run while batch i is packed, coded and scored; every batch still moves its own
        inputs in and its own results out.
        """
        dev = self.parts.blob.device
        main = torch.cuda.current_stream(dev)
        h2d, d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        xd, staged, free_ev = [None, None], [None, None], [None, None]
        outs, out_ev = [None] * ring, [None] * ring
        pending = None
        for i, xb in enumerate(host_batches):
            k = i & 1
            if not isinstance(xb, torch.Tensor):
                xb = np.ascontiguousarray(xb)
                if staged[k] is None or staged[k].shape != xb.shape:
                    staged[k] = torch.empty(xb.shape, dtype=torch.uint8).pin_memory()
                if free_ev[k] is not None:
                    free_ev[k].synchronize()              # the previous copy out of this staging buffer has been consumed
                staged[k].numpy()[...] = xb
                xb = staged[k]
            if xb.dtype != torch.uint8 or xb.dim() != 4:
                raise ValueError("edge batches must be uint8 (B,T,F,E)")
            if xd[k] is None or xd[k].shape != xb.shape:
                xd[k] = torch.empty(xb.shape, dtype=torch.uint8, device=dev)
            with torch.cuda.stream(h2d):
                if free_ev[k] is not None:
                    h2d.wait_event(free_ev[k])            # pack_edges of batch i-2 has finished reading this buffer
                xd[k].copy_(xb, non_blocking=True)
                arrived = torch.cuda.Event()
                arrived.record(h2d)
            main.wait_event(arrived)
            bits = ops.pack_edges(xd[k])
            free_ev[k] = torch.cuda.Event()
            free_ev[k].record(main)
            best, arg, _ = self(bits)
            scored = torch.cuda.Event()
            scored.record(main)
            r = i % ring
            if outs[r] is None or outs[r][0].shape != best.shape:
                outs[r] = (torch.empty(best.shape, dtype=torch.float32).pin_memory(),
                           torch.empty(arg.shape, dtype=arg.dtype).pin_memory())
            with torch.cuda.stream(d2h):
                d2h.wait_event(scored)
                outs[r][0].copy_(best, non_blocking=True)
                outs[r][1].copy_(arg, non_blocking=True)
                best.record_stream(d2h)
                arg.record_stream(d2h)
                out_ev[r] = torch.cuda.Event()
                out_ev[r].record(d2h)
            if pending is not None:
                out_ev[pending].synchronize()
                yield outs[pending]
            pending = r
        if pending is not None:
            out_ev[pending].synchronize()
            yield outs[pending]
