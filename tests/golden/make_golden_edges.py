"""Generate tests/golden/ref_edges.npz by EXECUTING the reference's own edge front end
(template_speech_rec/edge_signal_proc.py) on seeded inputs.  Run in the build container only (needs
/root/reference); the .npz is committed so the GPU box can check against it.

    python tests/golden/make_golden_edges.py

edge_signal_proc.py is Python 2 and cannot be imported under Python 3 (print statements elsewhere in the
file), so the functions on the path are cut out of the file by name and executed unchanged except for these
documented Python-2 / old-NumPy idioms, none of which touches the arithmetic:

  xrange                      -> range              (supplied through the namespace, no text change)
  print view_threshold        -> print(view_threshold)
  length/block_length         -> length//block_length       (Python-2 integer division)
  -sig_idx  (bool array)      -> ~sig_idx           (unary minus on bool arrays WAS logical not before NumPy 1.13)

and integer (not float) orientation / row-break arrays are passed in, because modern NumPy refuses the
float slice indices that _spread_edges builds from the float orientation table.
"""
import os
import re
import sys

import numpy as np
from scipy.ndimage import maximum_filter

REF = "/root/reference/template_speech_rec/edge_signal_proc.py"
HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ("spread_edgemap", "_spread_edges", "threshold_edgemap", "threshold_edge_block", "_edge_map_threshold_segments",
         "compute_edge_feature_row_breaks", "_edge_map_no_threshold")
SUBS = (("print view_threshold", "print(view_threshold)"), ("length/block_length", "length//block_length"),
        ("-sig_idx", "~sig_idx"))


def load_reference_functions(path=REF):
    src = open(path).read().split("\n")
    ns = {"np": np, "xrange": range, "maximum_filter": maximum_filter}
    for name in NAMES:
        start = next(i for i, l in enumerate(src) if re.match(r"def %s\(" % re.escape(name), l))
        end = next(i for i in range(start + 1, len(src)) if re.match(r"(def|class) ", src[i]))
        text = "\n".join(src[start:end])
        for a, b in SUBS:
            text = text.replace(a, b)
        exec(compile(text, "%s:%d" % (path, start + 1), "exec"), ns)
    return ns


def synth_spectrogram(rs, F, T):
    """log-spectrogram-like surface: smooth ridges + noise, with exact ties (quantised patch) and a flat region"""
    t, f = np.arange(T)[None, :], np.arange(F)[:, None]
    S = np.sin(0.31 * t + 0.5 * f) + 0.6 * np.cos(0.17 * t - 0.9 * f) + 0.35 * rs.randn(F, T)
    S[:, T // 2:T // 2 + 6] = np.round(S[:, T // 2:T // 2 + 6] * 2) / 2      # ties between neighbouring differences
    S[F // 2:, -9:-4] = 0.25                                                 # flat: zero differences
    return S * 0.05


def main():
    ref = load_reference_functions()
    rs = np.random.RandomState(20261019)
    out = {}
    for tag, (F, T, blk) in {"a": (12, 53, 20), "b": (9, 40, 40), "c": (16, 31, 40)}.items():
        S = synth_spectrogram(rs, F, T)
        E, breaks, ori = ref["_edge_map_no_threshold"](S)
        breaks_i, ori_i = breaks.astype(np.int64), ori.astype(np.int64)
        out["S_" + tag], out["E_" + tag] = S, E
        out["breaks_" + tag], out["ori_" + tag], out["blk_" + tag] = breaks_i, ori_i, np.int64(blk)
        for q in (0.7, 0.3):
            for s, kind in ((0, "unidirectional"), (1, "unidirectional"), (3, "unidirectional"), (2, "all")):
                Eb = E.copy()
                ref["_edge_map_threshold_segments"](Eb, blk, s, threshold=q, edge_orientations=ori_i,
                                                    edge_feature_row_breaks=breaks_i, spread_type=kind)
                out["bin_%s_q%d_s%d_%s" % (tag, int(q * 10), s, kind[:3])] = Eb.astype(np.uint8)
        # report_level: the per-orientation quantile values of the first block
        _, taus = ref["threshold_edgemap"](E[:, :min(blk, T)].copy(), 0.7, breaks_i, report_level=True)
        out["taus_" + tag] = taus
    # spread_edgemap alone on a random binary map, radius 1..3
    Tm = (rs.rand(8 * 10, 37) < 0.08).astype(np.float64)
    br = (np.arange(9) * 10).astype(np.int64)
    out["spread_in"], out["spread_breaks"] = Tm.astype(np.uint8), br
    for s in (1, 2, 3):
        Ts = Tm.copy()
        ref["spread_edgemap"](Ts, br, out["ori_a"], spread_length=s)
        out["spread_s%d" % s] = Ts.astype(np.uint8)
    path = os.path.join(HERE, "ref_edges.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%d arrays, %d bytes" % (len(out), os.path.getsize(path)))


if __name__ == "__main__":
    sys.exit(main())
