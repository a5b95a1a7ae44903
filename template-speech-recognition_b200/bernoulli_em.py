"""Drop-in for the first-version EM class ``bernoulli_em.Bernoulli_Mixture``
(template_speech_rec/bernoulli_em.py:8-131) on the same GPU kernels as BernoulliMixture.

Semantics kept from the reference: data (N,H,W); templates clipped to [.05,.95] (:48-49);
run_EM(tol) iterates while the log-likelihood improves by more than `tol` in ABSOLUTE terms
(:29), with no minimum / maximum iteration count; attributes are float64 arrays.
"""
import numpy as np
import torch

from .em_driver import EmDriver, GpuEmEngine


class Bernoulli_Mixture:
    def __init__(self, num_mix, data_mat, init_type='unif_rand', opt_type='expected'):
        data_mat = np.asarray(data_mat)
        self.num_mix = int(num_mix)
        self.height, self.width = data_mat.shape[1:]
        self.data_mat = data_mat
        self.num_data = data_mat.shape[0]
        self.weights = 1. / num_mix * np.ones(num_mix)
        self.opt_type = opt_type
        D = self.height * self.width
        # x > .5 is "on", x < .5 is "off" (get_template_loglikelihood, :116-120)
        self._engine = GpuEmEngine((data_mat.reshape(self.num_data, D) > .5).astype(np.uint8), D, self.num_mix)
        self._drv = EmDriver(self._engine, self.num_data, group=False)
        self.init_affinities_templates(init_type)

    def init_affinities_templates(self, init_type):                        # :51-79
        K, N = self.num_mix, self.num_data
        aff = np.zeros((N, K), np.float32)
        if init_type == 'unif_rand':
            rows = np.arange(K * (N // K))
            aff[rows, rows % K] = 1.
        elif init_type == 'specific':
            for mix_id in range(K):
                aff[K + mix_id, mix_id] = 1.
        else:
            raise ValueError("unknown init_type %r" % (init_type,))
        self._engine.set_affinities(aff)
        self._drv.init_from_affinities(.05)
        self._pull()

    def run_EM(self, tol, save_template=False):                            # :22-34
        """Iterate while the log-likelihood gains more than `tol` (absolute); rule evaluated on the device."""
        self.iterations, trace = self._drv.run(tol, mode="absolute", min_iter=0, max_iter=100000, min_prob=.05)
        self.loglikelihood_trace = trace
        self.loglikelihood = trace[-1]
        self._pull()

    def M_step(self):                                                      # :36-45
        self._drv.maximization(.05)

    def threshold_templates(self):                                         # :48-49
        self.templates = np.maximum(np.minimum(self.templates, .95), .05)

    def compute_loglikelihoods(self):                                      # :103-114
        return self._drv.expectation(hard=False)

    def get_template_loglikelihoods(self, data_mat=None):                  # :122-127
        """Scores (N,K) of `data_mat` (default: the training data) under the current templates; a pure accessor."""
        p = self._drv.params
        if data_mat is None or data_mat is self.data_mat:
            state = self._engine.state
        else:
            data_mat = np.asarray(data_mat)
            if data_mat.shape[1:] != (self.height, self.width):
                raise ValueError("data_mat must be (n, %d, %d)" % (self.height, self.width))
            D = self.height * self.width
            state = GpuEmEngine((data_mat.reshape(data_mat.shape[0], D) > .5).astype(np.uint8), D, self.num_mix).state
        scores = torch.empty((state.N, self.num_mix), dtype=torch.float32, device=self._engine.device)
        state.estep(p["logT"], p["logInvT"], p["logw"], scores_out=scores, update=False)
        return scores.cpu().numpy().astype(np.float64)

    def _pull(self):
        p = self._drv.params
        shape = (self.num_mix, self.height, self.width)
        self.templates = p["templates"].cpu().numpy().astype(np.float64).reshape(shape)
        self.log_templates = p["logT"].cpu().numpy().astype(np.float64).reshape(shape)
        self.log_invtemplates = p["logInvT"].cpu().numpy().astype(np.float64).reshape(shape)
        self.weights = p["weights"].cpu().numpy().astype(np.float64)
        self.affinities = self._engine.affinities().cpu().numpy().astype(np.float64)

    def get_templates(self):
        return self.templates

    def get_num_mix(self):
        return self.num_mix

    def get_weights(self):
        return self.weights

    def set_weights(self, new_weights):                                    # :97-100
        np.testing.assert_approx_equal(np.sum(new_weights), 1.)
        assert new_weights.shape == (self.num_mix,)
        self.weights = new_weights
        w = torch.as_tensor(np.asarray(new_weights, np.float32))
        self._drv.params["weights"].copy_(w)
        self._drv.params["logw"].copy_(torch.log(w))
