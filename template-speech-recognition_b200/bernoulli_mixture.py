"""Drop-in for ``bernoulli_mixture.BernoulliMixture`` (template_speech_rec/bernoulli_mixture.py:9-311)
with the E- and M-steps on the GPU.

Same constructor / run_EM signature and the same public attributes (templates,
work_templates, weights, affinities, log_templates, log_invtemplates, iterations, num_mix,
num_data, data_length, data_shape).  Differences, all documented in DESIGN.md:
  * posteriors are a true row soft-max (the reference's :277-279 is algebraically the same
    but can underflow to 0/0 with >= ~100 near-equal components);
  * template scores are exact integer sums of 24-bit fixed-point weights (three int8 digit planes on the
    tensor cores; the reference: float32/float64 BLAS); log-likelihoods agree to ~1e-6 relative;
  * any num_mix from 1 to 4095 (the reference trains 100-200-part dictionaries with this class);
  * prints are dropped; `init_type='unif_rand'` is the reference's deterministic round-robin.
Multi-GPU: pass ``row_range=(lo, hi)`` (see em_driver.shard_rows) on every rank of an
initialised torch.distributed group; the statistics are all-reduced once per iteration.
"""
import numpy as np
import torch

from .em_driver import EmDriver, GpuEmEngine


class BernoulliMixture:
    def __init__(self, num_mix, data_mat, init_type='unif_rand', init_seed=0, init_centers=None,
                 min_probability=.01, row_range=None, group=None):
        data_mat = np.asarray(data_mat)
        self.num_mix = int(num_mix)
        self.num_data = data_mat.shape[0]                      # global N (:77)
        self.data_shape = data_mat.shape[1:]
        self.data_length = int(np.prod(data_mat.shape[1:]))
        self.data_mat = data_mat.reshape(self.num_data, self.data_length)
        self.iterations = 0
        self.seed = init_seed
        self.min_probability = np.float32(min_probability)
        self.weights = 1. / self.num_mix * np.ones(self.num_mix, dtype=np.float32)
        self._range = (0, self.num_data) if row_range is None else tuple(row_range)
        self._group = group
        if init_type == 'preset' and init_centers is not None:
            self.num_mix = init_centers.shape[0]
        lo, hi = self._range
        self._engine = GpuEmEngine(self.data_mat[lo:hi], self.data_length, self.num_mix)
        self._drv = EmDriver(self._engine, self.num_data, group=group if group is not None else (None if row_range else False))
        self.init_affinities_templates(init_type, init_centers=init_centers)

    # ------------------------------------------------------------------ initialisation
    def init_affinities_templates(self, init_type, init_centers=None):     # :195-237
        K, N = self.num_mix, self.num_data
        lo, hi = self._range
        if init_type in ('unif_rand', 'specific'):
            aff = np.zeros((hi - lo, K), np.float32)
            if init_type == 'unif_rand':
                full = K * (N // K)
                rows = np.arange(lo, min(hi, full))
                aff[rows - lo, rows % K] = 1.
            else:  # 'specific': one datum per component, index K*1 + mix_id (:228-229)
                for mix_id in range(K):
                    r = K + mix_id
                    if lo <= r < hi:
                        aff[r - lo, mix_id] = 1.
            self._engine.set_affinities(aff)
            self._drv.init_from_affinities(self.min_probability)
        elif init_type == 'preset' and init_centers is not None:
            T = np.clip(init_centers.reshape(K, self.data_length), self.min_probability, 1 - self.min_probability)
            p = self._drv.params
            p["templates"].copy_(torch.as_tensor(T.astype(np.float32)))
            p["logT"].copy_(torch.as_tensor(np.log(T).astype(np.float32)))
            p["logInvT"].copy_(torch.as_tensor(np.log(1 - T).astype(np.float32)))
            p["weights"].fill_(1.0 / K)
            p["logw"].fill_(float(np.log(np.float32(1.0 / K))))
        else:
            raise ValueError("unknown init_type %r" % (init_type,))
        self._pull_params()

    # ------------------------------------------------------------------------- EM loop
    def run_EM(self, tol, min_probability=0.05, debug_plot=False, hard_assignment=False, rand_seed=None,
               use_templates=None, max_iter=500, min_iter=10):
        """:108-160, same stopping rule: iterate while the relative log-likelihood change
        exceeds `tol` (and iterations < max_iter) or iterations < min_iter.  The rule is evaluated on the
        device (em_driver.EmDriver.run), so no log-likelihood is read back between iterations."""
        self.hard_assignment = hard_assignment
        self.min_probability = np.float32(min_probability)
        self.iterations, self.loglikelihood_trace = self._drv.run(
            tol, mode="relative", min_iter=min_iter, max_iter=max_iter, hard=hard_assignment,
            min_prob=self.min_probability)
        self.loglikelihood = self.loglikelihood_trace[-1]
        self._pull_params()
        self.set_templates()                                               # :160

    def M_step(self):                                                      # :177-184
        self._drv.maximization(self.min_probability)
        self._pull_params()

    def _compute_loglikelihoods(self):                                     # :263-280
        ll = self._drv.expectation(hard=getattr(self, "hard_assignment", False))
        self.affinities = self._engine.affinities().cpu().numpy()
        return ll

    def get_template_loglikelihoods(self):                                 # :287-291
        """Pure accessor, as in the reference: scores of the (local) training rows under the current
        templates; posteriors, statistics and the log-likelihood are left untouched."""
        scores = torch.empty((self._engine.N, self.num_mix), dtype=torch.float32, device=self._engine.device)
        p = self._drv.params
        self._engine.state.estep(p["logT"], p["logInvT"], p["logw"], scores_out=scores, update=False)
        return scores.cpu().numpy()

    # ----------------------------------------------------------------------- accessors
    def _pull_params(self):
        p = self._drv.params
        self.work_templates = p["templates"].cpu().numpy()
        self.weights = p["weights"].cpu().numpy()
        self.log_templates = p["logT"].cpu().numpy()
        self.log_invtemplates = p["logInvT"].cpu().numpy()
        self.affinities = self._engine.affinities().cpu().numpy()

    def threshold_templates(self):                                         # :192-193
        self.work_templates = np.clip(self.work_templates, self.min_probability, 1 - self.min_probability)

    def set_templates(self):                                               # :245-248
        self.templates = self.work_templates.reshape((self.num_mix,) + self.data_shape)
        self.log_templates = np.log(self.templates)
        self.log_invtemplates = np.log(1 - self.templates)

    def set_weights(self, new_weights):                                    # :257-260
        np.testing.assert_approx_equal(np.sum(new_weights), 1.)
        assert new_weights.shape == (self.num_mix,)
        self.weights = new_weights
        w = torch.as_tensor(np.asarray(new_weights, np.float32))
        self._drv.params["weights"].copy_(w)
        self._drv.params["logw"].copy_(torch.log(w))

    def save(self, filename, save_affinities=False):                       # :296-311
        entries = dict(templates=self.templates, weights=self.weights)
        if save_affinities:
            entries['affinities'] = self.affinities
        np.savez(filename, **entries)
