"""Drop-in for ``test_template.score_template_background_section``
(template_speech_rec/test_template.py:25-41), the single-window LLR that
classification.Classifier.score* (classification.py:44-56, 87-98) and
Experiment.get_detection_scores_slow (template_experiments.py:400) call.
"""
import numpy as np

from . import ops
from ._np import to_dev


def score_template_background_section(template, bgd, E, front_bgd_pad=0, back_bgd_pad=0):
    """template (H,L) probabilities, bgd (H,), E (H,L) window.  Returns (sum(E * W), C) with
    W = log((T/B)/((1-T)/(1-B))), C = sum log((1-T)/(1-B)); the contraction runs on the GPU."""
    template, bgd, E = np.asarray(template, np.float64), np.asarray(bgd, np.float64), np.asarray(E)
    assert E.shape[1] == template.shape[1]
    T = np.hstack((np.tile(bgd, (front_bgd_pad, 1)).transpose(), template,
                   np.tile(bgd, (back_bgd_pad, 1)).transpose()))
    L = T.shape[1]
    U_bgd = np.tile(bgd, (L, 1)).transpose()
    C_exp_inv_long = (1. - T) / (1. - U_bgd)
    C = np.log(C_exp_inv_long).sum()
    W = np.log((T / U_bgd) / C_exp_inv_long)
    if front_bgd_pad or back_bgd_pad:
        # the reference multiplies E (H,L0) with the padded W (H,L): shapes only agree without padding
        if E.shape != W.shape:
            raise ValueError("operands could not be broadcast together with shapes %s %s" % (E.shape, W.shape))
    H = T.shape[0]
    Fm = to_dev(np.ascontiguousarray(E.T, dtype=np.float32).reshape(1, L, H))
    LF = to_dev(np.ascontiguousarray(W.T, dtype=np.float32).reshape(1, L, H))
    return float(ops.detect_float(Fm, LF)[0, 0, 0].item()), C
