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
                 detect_planes=ops.PLANES_INT8X2):
        self.parts = ops.PartsModel(log_parts, log_invparts, device=device)
        self.bank = ops.FilterBank(filters, cs=cs, device=device, planes=detect_planes)
        self.edge_threshold = int(edge_threshold)
        self.part_spread = (int(part_spread[0]), int(part_spread[1]))
        self.edge_spread_radius = int(edge_spread_radius)
        self.edge_block_len = int(edge_block_len)
        self.edge_orientations = edge_orientations
        # K2: the 16-bit hi+lo planes up to 126 parts (one 128-column accumulator per tile: 0.63 ms at config 3 against
        # 0.75 ms), the int8 form above (200 parts: 1.48 against 1.91 ms).  The int8 form's codes equal the reference's
        # float64 arg-max everywhere (set impl_code = ops.IMPL_TENSOR_INT8 to have that at any size); the hi+lo form
        # may differ at near-ties below 1e-5 relative.
        self.impl_code = ops.IMPL_TENSOR_INT8 if self.parts.K > 126 else ops.IMPL_TENSOR
        self.impl_detect = ops.IMPL_TENSOR
        self.fused_max = True
        # K1a inside K2's producers (tsr_code_parts edge_spread_radius = 1) is bit-identical but measured SLOWER than the
        # two launches (1.01 ms against 0.69 + 0.06 ms at config 3: the four producer warps become the bottleneck of an
        # MMA-bound kernel), so it is off by default; see DESIGN.md section 3.
        self.fuse_edge_spread = False

    def spread(self, bits):
        """K1a as its own launch (what code() uses when the spreading cannot ride inside K2): per edge_block_len-frame
        block, the ragged last block as the tail of the window of the final block_len frames (edge_signal_proc.py:
        1595-1633)."""
        if self.edge_spread_radius <= 0:
            return bits
        return ops.spread_edges(bits, self.edge_spread_radius, self.edge_orientations, "line", max(self.edge_block_len, 0),
                                tail_window=True)

    def code(self, bits):
        """K1a + K2 + K1b: packed edges -> (codes, part feature bits).  With fuse_edge_spread set (and the reference's
        direction table, radius 1, main_multiprocessing.py:50-52) the edge spreading runs inside K2's producers."""
        fused = (self.fuse_edge_spread and self.edge_spread_radius == 1 and self.edge_orientations is None
                 and self.impl_code in (ops.IMPL_TENSOR, ops.IMPL_TENSOR_INT8))
        if not fused:
            bits = self.spread(bits)
        codes, _, _ = ops.code_parts(bits, self.parts, self.edge_threshold, impl=self.impl_code,
                                     edge_spread_radius=1 if fused else 0, edge_block_len=self.edge_block_len if fused else 0)
        fb = ops.spread_codes(codes, self.parts.K, *self.part_spread)
        return codes, fb

    def score(self, fb, want_scores=False):
        """K3 with the mixture max fused into its epilogue: part feature bits -> (best, argbest, per-template
        scores or None).  The (B,To,M) score matrix is only materialised on request (or on the SIMT cross-check
        path, which reduces it with a second kernel)."""
        Fo_K = fb.shape[2]
        if self.bank.D > Fo_K * 32:
            raise ValueError("filter width %d does not match the coded feature rows" % self.bank.D)
        if self.impl_detect == ops.IMPL_TENSOR and self.fused_max:
            return ops.detect_best(fb, self.bank, want_scores=want_scores)
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

    def stream(self, host_batches, ring=3, pack="host", pack_threads=None):
        """Score a sequence of host batches with packing, copies and kernels overlapped.

        host_batches: iterable of edge batches in host memory -- uint8 (B,T,F,E) edge tensors (the reference's
        dtype; E = 8) or pre-packed uint8 (B,T,F) edge bytes (bit e = orientation e, what ``ops.pack_edges``
        produces).  Pinned torch tensors copy asynchronously; NumPy arrays are accepted.  Yields, per batch and in
        order, ``(best float32 (B,To), argbest int16 (B,To))`` as pinned host tensors that stay valid until
        ``ring - 1`` further batches have been requested.

        pack="host" (default): (B,T,F,E) batches are packed to one byte per (t,f) by a pool of host threads
        (``tsr_host_pack_edges``, pure data movement) BEFORE the copy, so 1/8 of the bytes cross PCIe; a feeder thread
        packs batch i+1 while batch i is copied and scored.  pack="device": the unpacked tensor is copied and packed
        by the GPU (8x the H2D bytes; for hosts with few cores and a fast link).  pack="split": both at once -- the first
        rows of every batch are packed by the host threads, the rest cross PCIe unpacked and are packed by the GPU; the
        split point follows the measured rates of the two paths from batch to batch (for boxes where neither the cores
        per GPU nor the link alone keep up: 8 GPUs behind 32 cores).  Needs pinned (B,T,F,8) torch tensors.

        Streams: H2D of batch i+1 (into the other of two device buffers) and D2H of batch i-1's results run while
        batch i is coded and scored; every batch still moves its own inputs in and its own results out.
        """
        import os
        import queue
        import threading
        from . import _lib
        if pack not in ("host", "device", "split"):
            raise ValueError("pack must be 'host', 'device' or 'split'")
        dev = self.parts.blob.device
        nthreads = int(pack_threads) if pack_threads else max(1, min(32, (os.cpu_count() or 2) - 1))
        q = queue.Queue(maxsize=2)
        stop = threading.Event()
        # Four staging buffers for a queue of two: when the feeder reuses slot i % 4 for batch i + 4 it has queued
        # batches i + 2 and i + 3, so the consumer has taken batch i + 1 and therefore finished issuing batch i
        # (whose copy-done event is then in `consumed`).
        NSTAGE = 4
        # pinned buffers the feeder packs / stages into, and the pinned result ring: kept on the pipeline between calls
        # (pinning host memory costs milliseconds per buffer)
        cache = self.__dict__.setdefault("_stream_cache", {"staging": [None] * NSTAGE, "outs": {}})
        if cache.get("busy"):
            cache = {"staging": [None] * NSTAGE, "outs": {}}     # a second stream() running concurrently gets its own
        cache["busy"] = True
        staging = cache["staging"]
        split = cache.setdefault("split", {"rows": None})     # split mode: rows of a batch the host packs (adapted per batch, kept between calls)
        consumed = [None] * NSTAGE               # CUDA event: the H2D copy out of staging[j] has finished

        import time as _time
        stats = self.__dict__.setdefault("stream_stats", {})
        stats.update(pack_s=0.0, pack_calls=0, feeder_wait_s=0.0, consumer_wait_s=0.0, h2d_issue_s=0.0, compute_issue_s=0.0,
                     d2h_issue_s=0.0, result_wait_s=0.0)

        def feeder():
            try:
                for i, xb in enumerate(host_batches):
                    if stop.is_set():
                        return
                    j = i % NSTAGE
                    is_t = isinstance(xb, torch.Tensor)
                    if not is_t:
                        xb = np.ascontiguousarray(xb)
                    if xb.dtype not in (torch.uint8, np.uint8) or len(xb.shape) not in (3, 4):
                        raise ValueError("edge batches must be uint8 (B,T,F,E) or packed uint8 (B,T,F)")
                    if len(xb.shape) == 4 and xb.shape[3] != 8 and pack == "host":
                        raise ValueError("host packing needs E == 8 orientation planes")
                    direct = is_t and xb.is_pinned() and xb.is_contiguous() and (len(xb.shape) == 3 or pack == "device")
                    if direct:                   # already what the copy engine needs
                        q.put((xb, None))
                        continue
                    if pack == "split" and is_t and len(xb.shape) == 4 and xb.shape[3] == 8 and xb.shape[0] >= 16 and xb.is_pinned() \
                            and xb.is_contiguous():
                        Bn = xb.shape[0]
                        nh = int(min(max(split["rows"] if split["rows"] is not None else Bn // 2, 1), Bn - 1)) if Bn > 1 else Bn
                        shape = tuple(xb.shape[:3])
                        if staging[j] is None or tuple(staging[j].shape) != shape:
                            staging[j] = torch.empty(shape, dtype=torch.uint8).pin_memory()
                        elif consumed[j] is not None:
                            t0 = _time.perf_counter()
                            consumed[j].synchronize()
                            stats["feeder_wait_s"] += _time.perf_counter() - t0
                        t0 = _time.perf_counter()
                        _lib.call("tsr_host_pack_edges", xb.data_ptr(), staging[j].data_ptr(), nh * shape[1] * shape[2], 8, nthreads)
                        dt = _time.perf_counter() - t0
                        stats["pack_s"] += dt
                        stats["pack_calls"] += 1
                        q.put(((staging[j][:nh], xb[nh:], dt), j))
                        continue
                    shape = tuple(xb.shape[:3]) if (len(xb.shape) == 4 and pack == "host") else tuple(xb.shape)
                    if staging[j] is None or tuple(staging[j].shape) != shape:
                        staging[j] = torch.empty(shape, dtype=torch.uint8).pin_memory()
                    elif consumed[j] is not None:
                        t0 = _time.perf_counter()
                        consumed[j].synchronize()
                        stats["feeder_wait_s"] += _time.perf_counter() - t0
                    src = xb.contiguous() if is_t else xb
                    if len(xb.shape) == 4 and pack == "host":
                        sp = src.data_ptr() if is_t else src.ctypes.data
                        t0 = _time.perf_counter()
                        _lib.call("tsr_host_pack_edges", sp, staging[j].data_ptr(), int(np.prod(shape)), 8, nthreads)
                        stats["pack_s"] += _time.perf_counter() - t0
                        stats["pack_calls"] += 1
                    elif is_t:
                        staging[j].copy_(src)
                    else:
                        staging[j].numpy()[...] = src
                    q.put((staging[j], j))
                q.put(None)
            except BaseException as e:  # noqa: BLE001 -- re-raised in the consumer
                q.put(e)

        th = threading.Thread(target=feeder, daemon=True, name="tsr-stream-feeder")
        th.start()
        main = torch.cuda.current_stream(dev)
        h2d, d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        xd, free_ev = [None, None], [None, None]
        xu = [None, None]                        # split mode: the unpacked rows of a batch on the device
        split_ev = [None, None]                  # split mode: (start, end) events around the unpacked copy, host pack seconds, rows
        outs, out_ev = cache["outs"].setdefault(ring, [None] * ring), [None] * ring
        pending = None
        i = 0
        try:
            while True:
                t0 = _time.perf_counter()
                item = q.get()
                stats["consumer_wait_s"] += _time.perf_counter() - t0
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                xb, slot = item
                k = i & 1
                t1 = _time.perf_counter()
                if isinstance(xb, tuple):        # split mode: (host-packed head, unpacked tail, host pack seconds)
                    hp, tail, pack_dt = xb
                    nh, nt_rows = hp.shape[0], tail.shape[0]
                    full = (nh + nt_rows,) + tuple(hp.shape[1:])
                    with torch.cuda.stream(h2d):
                        if xd[k] is None or tuple(xd[k].shape) != full or xu[k] is None or tuple(xu[k].shape[1:]) != tuple(tail.shape[1:]) \
                                or xu[k].shape[0] < full[0]:
                            h2d.wait_stream(main)
                            xd[k] = torch.empty(full, dtype=torch.uint8, device=dev)
                            xu[k] = torch.empty((full[0],) + tuple(tail.shape[1:]), dtype=torch.uint8, device=dev)
                            xd[k].record_stream(main)
                            xu[k].record_stream(main)
                        elif free_ev[k] is not None:
                            h2d.wait_event(free_ev[k])
                        xd[k][:nh].copy_(hp, non_blocking=True)
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(h2d)
                        xu[k][:nt_rows].copy_(tail, non_blocking=True)
                        e1.record(h2d)
                        arrived = torch.cuda.Event()
                        arrived.record(h2d)
                        consumed[slot] = arrived
                    main.wait_event(arrived)
                    split_ev[k] = (e0, e1, pack_dt, nh, nt_rows)
                    t2 = _time.perf_counter()
                    stats["h2d_issue_s"] += t2 - t1
                    ops.pack_edges(xu[k][:nt_rows], out=xd[k][nh:])          # the GPU packs its share into place
                    bits = xd[k]
                else:
                    with torch.cuda.stream(h2d):
                        if xd[k] is None or xd[k].shape != xb.shape:
                            # A fresh block may be one that kernels still in flight on `main` were using (the caching
                            # allocator reuses blocks per stream): order the first copy after everything queued there.
                            h2d.wait_stream(main)
                            xd[k] = torch.empty(xb.shape, dtype=torch.uint8, device=dev)
                            xd[k].record_stream(main)
                        elif free_ev[k] is not None:
                            h2d.wait_event(free_ev[k])        # batch i-2 has finished reading this buffer
                        xd[k].copy_(xb, non_blocking=True)
                        arrived = torch.cuda.Event()
                        arrived.record(h2d)
                        if slot is not None:
                            consumed[slot] = arrived
                    main.wait_event(arrived)
                    t2 = _time.perf_counter()
                    stats["h2d_issue_s"] += t2 - t1
                    bits = ops.pack_edges(xd[k]) if xd[k].dim() == 4 else xd[k]
                best, arg, _ = self(bits)
                free_ev[k] = torch.cuda.Event()
                free_ev[k].record(main)                   # the whole step is queued: xd[k] (== bits when packed) is free after it
                t3 = _time.perf_counter()
                stats["compute_issue_s"] += t3 - t2
                r = i % ring
                if outs[r] is None or outs[r][0].shape != best.shape:
                    outs[r] = (torch.empty(best.shape, dtype=torch.float32).pin_memory(),
                               torch.empty(arg.shape, dtype=arg.dtype).pin_memory())
                with torch.cuda.stream(d2h):
                    d2h.wait_event(free_ev[k])
                    outs[r][0].copy_(best, non_blocking=True)
                    outs[r][1].copy_(arg, non_blocking=True)
                    best.record_stream(d2h)
                    arg.record_stream(d2h)
                    out_ev[r] = torch.cuda.Event()
                    out_ev[r].record(d2h)
                t4 = _time.perf_counter()
                stats["d2h_issue_s"] += t4 - t3
                if pending is not None:
                    out_ev[pending].synchronize()
                    stats["result_wait_s"] += _time.perf_counter() - t4
                    pk = (i - 1) & 1
                    if split_ev[pk] is not None:              # batch i-1 is complete: its copy time and pack time set the next split
                        e0, e1, pack_dt, nh, nt_rows = split_ev[pk]
                        split_ev[pk] = None
                        c_h = pack_dt / max(nh, 1)                        # host seconds per packed row
                        c_d = e0.elapsed_time(e1) * 1e-3 / max(nt_rows, 1)   # link seconds per unpacked row
                        tot = nh + nt_rows
                        want = tot * c_d / max(c_h + c_d, 1e-12)
                        cur = split["rows"] if split["rows"] is not None else nh
                        new_rows = int(round((0.5 * cur + 0.5 * want) / 8.0)) * 8
                        split["rows"] = min(max(new_rows, 8), max(tot - 8, 8))
                        stats["split_rows"] = split["rows"]
                    yield outs[pending]
                pending = r
                i += 1
            if pending is not None:
                out_ev[pending].synchronize()
                yield outs[pending]
        finally:
            stop.set()
            cache["busy"] = False
            while th.is_alive():                          # unblock a feeder waiting on the full queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    th.join(timeout=0.05)
