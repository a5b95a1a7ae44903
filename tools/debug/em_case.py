"""Debug: GPU EM vs the NumPy restatement, iteration by iteration, at the config-4 shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import oracle
import tsr_b200
from tsr_b200 import synth, ops
from tsr_b200.bernoulli_mixture import BernoulliMixture

N, D, K = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (2000, 12800, 16)))
X, _, _, _ = synth.em_data(N=N, D=D, K=K, seed=3)
a = oracle.BernoulliMixture(K, X); a.hard_assignment = False; a.min_probability = np.float32(.05)
b = BernoulliMixture(K, X); b.min_probability = np.float32(.05)
c = BernoulliMixture(K, X); c.min_probability = np.float32(.05)   # SIMT path
Xd = X.astype(np.float64)
for it in range(3):
    la = a._compute_loglikelihoods()
    lb = b._compute_loglikelihoods()
    c._drv._estep(False, 1, None, False); lc = c._engine.loglik(); c.affinities = c._engine.affinities().cpu().numpy()
    print("it %d  ll oracle %.3f  gpu %.3f (rel %.2e)  simt %.3f (rel %.2e)" % (it, la, lb, abs(la - lb) / abs(la), lc, abs(la - lc) / abs(la)))
    da = np.abs(a.affinities - b.affinities)
    print("   aff max diff tensor %.3e  simt %.3e   argmax mismatches %d / %d" % (da.max(), np.abs(a.affinities - c.affinities).max(),
          (a.affinities.argmax(1) != b.affinities.argmax(1)).sum(), (a.affinities.argmax(1) != c.affinities.argmax(1)).sum()))
    # statistics exactness of the tensor path
    st = b._engine.state
    q = np.rint(b.affinities * np.float32(8355711)).astype(np.int64)
    stats = st.stats.cpu().numpy()
    want = q.T @ X.astype(np.int64)
    got = stats[:K * D].reshape(K, D)
    bad = (got != want)
    print("   numerators exact: %s (%d wrong, rows with errors %s)  resp exact: %s" % (not bad.any(), bad.sum(), np.unique(np.nonzero(bad)[0])[:8], np.array_equal(stats[K * D:K * D + K], q.sum(0))))
    if bad.any():
        kk, dd = np.nonzero(bad)
        print("   first wrong: k=%d d=%d got %d want %d; d range %d..%d" % (kk[0], dd[0], got[kk[0], dd[0]], want[kk[0], dd[0]], dd.min(), dd.max()))
    a.M_step(); b.M_step(); c.M_step()
    print("   after M: templates max diff tensor %.3e simt %.3e; weights diff %.3e; min weight %.5f" % (np.abs(a.work_templates - b.work_templates).max(),
          np.abs(a.work_templates - c.work_templates).max(), np.abs(a.weights - b.weights).max(), a.weights.min()))
# the queued loop
d = BernoulliMixture(K, X); d.run_EM(1e-9, min_probability=.05, max_iter=3, min_iter=3)
print("run() trace", d.loglikelihood_trace)
