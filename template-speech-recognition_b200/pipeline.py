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
