"""Generate tests/golden/ref_models.npz by EXECUTING the reference's own pure-Python model code on seeded
inputs: bernoulli_mixture.BernoulliMixture, bernoulli_em.Bernoulli_Mixture,
estimate_template.construct_linear_filter(s) and test_template.score_template_background_section.
Run in the build container only (needs /root/reference); the .npz is committed.

    python tests/golden/make_golden_models.py

Those modules are Python 2 and do not import under Python 3.  They are executed from their source text with
these documented Python-2 / old-NumPy idioms patched -- none touches the arithmetic:

  print "..." / print (...)           -> print(...)         (statement -> function; output discarded)
  xrange                              -> range              (supplied through the namespace)
  idx = range(n); random.shuffle(idx) -> idx = list(range(n))   (the shuffled idx is never used afterwards)
  self.num_data/self.num_mix          -> self.num_data//self.num_mix     (Python-2 integer division)
  np.array(map(...))                  -> np.array(list(map(...)))
  tuple( np.ones(k))  as np.tile reps -> tuple( np.ones(k,dtype=int))    (float reps are refused by NumPy >= 1.12)
"""
import contextlib
import io
import itertools
import os
import random
import re
import sys

import numpy as np

REF = "/root/reference/template_speech_rec"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def _py3(text):
    text = re.sub(r"^(\s*)print (.+)$", r"\1print(\2)", text, flags=re.M)
    text = text.replace("idx = range(", "idx = list(range(").replace("list(range(self.num_data)", "list(range(self.num_data))")
    text = text.replace("idx = list(range(data_mat.shape[0])", "idx = list(range(data_mat.shape[0]))")
    text = text.replace("self.num_data/self.num_mix", "self.num_data//self.num_mix")
    text = text.replace("np.array(map(lambda x:\\\n                                self.get_template_loglikelihood(x),\\\n                                data_mat))",
                        "np.array(list(map(lambda x:\\\n                                self.get_template_loglikelihood(x),\\\n                                data_mat)))")
    text = text.replace("tuple( np.ones(len(T.shape)-1))", "tuple( np.ones(len(T.shape)-1,dtype=int))")
    return text


def load_module(name):
    ns = {"xrange": range, "__name__": "ref_" + name}
    exec(compile(_py3(open(os.path.join(REF, name + ".py")).read()), os.path.join(REF, name + ".py"), "exec"), ns)
    return ns


def load_functions(name, funcs, extra):
    src = open(os.path.join(REF, name + ".py")).read().split("\n")
    ns = dict(extra, xrange=range)
    for f in funcs:
        start = next(i for i, l in enumerate(src) if re.match(r"def %s\(" % re.escape(f), l))
        end = next(i for i in range(start + 1, len(src)) if re.match(r"(def|class) ", src[i]))
        exec(compile(_py3("\n".join(src[start:end])), "%s.py:%d" % (name, start + 1), "exec"), ns)
    return ns


def traced(m, method):
    """record every value the E-step returns (run_EM's log-likelihood sequence)"""
    trace, orig = [], getattr(m, method)

    def wrapped():
        trace.append(orig())
        return trace[-1]
    setattr(m, method, wrapped)
    return trace


def planted(rs, N, D, K):
    T = np.clip(rs.rand(K, D) ** 2, .05, .95)
    z = rs.randint(0, K, N)
    return (rs.rand(N, D) < T[z]).astype(np.uint8)


def main():
    out = {}
    quiet = io.StringIO()
    bm = load_module("bernoulli_mixture")["BernoulliMixture"]
    be = load_module("bernoulli_em")["Bernoulli_Mixture"]
    rs = np.random.RandomState(20261020)
    # ---- BernoulliMixture: (N, K, data shape, run_EM kwargs)
    cases = {"a": (121, 3, (6, 5), dict(tol=1e-6, min_probability=.05)),
             "b": (200, 4, (30,), dict(tol=1e-5, min_probability=.01, hard_assignment=True)),
             "c": (90, 2, (3, 4, 2), dict(tol=1e-4, min_probability=.05, max_iter=6, min_iter=2))}
    with contextlib.redirect_stdout(quiet):
        for tag, (N, K, shape, kw) in cases.items():
            X = planted(rs, N, int(np.prod(shape)), K).reshape((N,) + shape)
            m = bm(K, X)
            trace = traced(m, "_compute_loglikelihoods")
            m.run_EM(**kw)
            out["bm_%s_X" % tag] = X
            out["bm_%s_K" % tag] = np.int64(K)
            out["bm_%s_kw" % tag] = np.array(repr(kw))
            out["bm_%s_templates" % tag] = np.asarray(m.templates)
            out["bm_%s_weights" % tag] = np.asarray(m.weights)
            out["bm_%s_affinities" % tag] = np.asarray(m.affinities)
            out["bm_%s_iterations" % tag] = np.int64(m.iterations)
            out["bm_%s_loglik" % tag] = np.array(trace, np.float64)
        # preset initial centres (:211-217)
        X = planted(rs, 150, 24, 3)
        centers = np.clip(rs.rand(3, 24), .1, .9)
        m = bm(3, X, init_type='preset', init_centers=centers)
        trace = traced(m, "_compute_loglikelihoods")
        m.run_EM(1e-5, min_probability=.02)
        out.update(bm_p_X=X, bm_p_centers=centers, bm_p_templates=np.asarray(m.templates), bm_p_weights=np.asarray(m.weights),
                   bm_p_affinities=np.asarray(m.affinities), bm_p_iterations=np.int64(m.iterations),
                   bm_p_loglik=np.array(trace, np.float64))
        # ---- first-version EM (bernoulli_em.py)
        X = planted(rs, 80, 20, 2).reshape(80, 4, 5)
        m = be(2, X)
        trace = traced(m, "compute_loglikelihoods")
        m.run_EM(1e-4)
        out.update(be_X=X, be_templates=m.templates, be_weights=m.weights, be_affinities=m.affinities,
                   be_loglik=np.array(trace, np.float64))
    # ---- linear filters (estimate_template.py:156-172, 230-251) and the single-window score (test_template.py:25-41)
    et = load_functions("estimate_template", ("construct_linear_filters", "construct_linear_filter"), {"np": np, "itertools": itertools})
    tt = load_functions("test_template", ("score_template_background_section",), {"np": np})
    T3 = rs.rand(7, 5, 4)          # (L, F, K) with values outside [.01,.99] after squaring
    T3[0, 0, 0], T3[1, 1, 1] = 0.0, 1.0
    bg3 = np.clip(0.3 * rs.rand(5, 4), .01, .4)
    W, c = et["construct_linear_filter"](T3, bg3)
    Wc, cc = et["construct_linear_filter"](T3, bg3, all_cs=True)
    out.update(lf_T=T3, lf_bgd=bg3, lf_W=W, lf_c=np.float64(c), lf_W_allcs=Wc, lf_c_allcs=cc)
    fs = et["construct_linear_filters"]((T3, T3[::-1] ** 2), bg3)
    out.update(lfs_W0=fs[0][0], lfs_c0=np.float64(fs[0][1]), lfs_W1=fs[1][0], lfs_c1=np.float64(fs[1][1]))
    tmpl = np.clip(rs.rand(12, 6), .05, .95)      # (features, time)
    bgd = np.clip(0.3 * rs.rand(12), .01, .4)
    E = (rs.rand(12, 6) < 0.3).astype(np.float64)
    s, C = tt["score_template_background_section"](tmpl, bgd, E)
    Ep = (rs.rand(12, 9) < 0.3).astype(np.float64)
    s2, C2 = tt["score_template_background_section"](np.hstack([tmpl, tmpl[:, :3]])[:, :9], bgd, Ep)
    out.update(sc_template=tmpl, sc_bgd=bgd, sc_E=E, sc_s=np.float64(s), sc_C=np.float64(C), sc_E2=Ep, sc_s2=np.float64(s2),
               sc_C2=np.float64(C2))
    path = os.path.join(HERE, "ref_models.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%d arrays, %d bytes" % (len(out), os.path.getsize(path)))


if __name__ == "__main__":
    random.seed(0)
    sys.exit(main())
