"""DGP_VI / DGP_IWVI with the reference's constructor and public methods (reference dgps_with_iwvi/models.py:9-150 and
the gpflow.models.GPModel conveniences its tests and scripts call: compute_log_likelihood, predict_f,
predict_f_full_cov, predict_f_multisample, predict_y_samples).  Objectives and their gradients run through
engine.Engine (C ABI -> CUDA); there is no CPU path."""
import numpy as np
import torch

from . import settings
from .engine import Engine, FlatParams
from .params import Parameterized, ParamList


class Minibatch:
    """gpflow.params.Minibatch(value, batch_size, seed): an endless shuffled stream of row indices.  The reference
    builds one per data array with the same seed so rows stay aligned (models.py:25-26); here ONE index stream
    serves X and Y.  TF's shuffle order is not reproducible without TF; a numpy Generator(seed) permutation per epoch
    replaces it."""

    def __init__(self, n, batch_size, seed=0):
        self.n, self.batch_size = int(n), int(batch_size)
        self.rng = np.random.default_rng(seed)
        self.perm = self.rng.permutation(self.n)
        self.pos = 0

    def next(self, count=None):
        count = count or self.batch_size
        out = []
        need = count
        while need > 0:
            take = min(need, self.n - self.pos)
            out.append(self.perm[self.pos:self.pos + take])
            self.pos += take
            need -= take
            if self.pos == self.n:
                self.perm = self.rng.permutation(self.n)
                self.pos = 0
        return np.concatenate(out)


class DGP_VI(Parameterized):
    """reference models.py:9-107."""
    _mode = 'vi'

    def __init__(self, X, Y, layers, likelihood, num_samples=1, minibatch_size=None, name=None):
        Parameterized.__init__(self, name=name)
        self.likelihood = likelihood
        X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        self.num_data = X.shape[0]
        self.num_samples = num_samples
        self.Dx, self.Dy = X.shape[1], Y.shape[1]
        dev = settings.device()
        self.X = torch.as_tensor(X, dtype=settings.float_type).to(dev)
        self.Y = torch.as_tensor(Y, dtype=settings.float_type).to(dev)
        self.minibatch_size = minibatch_size
        self.minibatch = None if minibatch_size is None else Minibatch(self.num_data, minibatch_size, seed=0)
        self.layers = ParamList(layers)
        self._engines = {}
        self.noise_seed = 0
        self._evals = 0

    # ---- layer loop (reference models.py:30-46) -------------------------------------------------
    def propagate(self, X, full_cov=False, inference_amorization_inputs=None, is_sampled_local_regularizer=False,
                  eps=None):
        """reference models.py:30-46: the sequential layer loop over the operator-level layers, returning
        (samples, means, covs, kls, kl_types) with one entry per layer.  X [..., Dx] (the reference passes the tiled
        [S*N, Dx] / [N, K, Dx] / [S, N, Dx] inputs) as a CUDA float64 tensor or array.  Every layer runs through its
        own `propagate` (C ABI -> CUDA, torch autograd for gradients); `eps` optionally injects the N(0,1) draws, one
        entry per layer (None = drawn with torch.randn).  Whole-model objectives do not come through here but through
        engine.Engine (same kernels, no autograd graph)."""
        X = torch.as_tensor(np.asarray(X, dtype=np.float64) if not torch.is_tensor(X) else X,
                            dtype=settings.float_type).to(self.X.device)
        samples, means, covs, kls, kl_types = [X], [], [], [], []
        for i, layer in enumerate(self.layers):
            sample, mean, cov, kl = layer.propagate(samples[-1], full_cov=full_cov,
                                                    inference_amorization_inputs=inference_amorization_inputs,
                                                    is_sampled_local_regularizer=is_sampled_local_regularizer,
                                                    eps=None if eps is None else eps[i])
            samples.append(sample)
            means.append(mean)
            covs.append(cov)
            kls.append(kl)
            kl_types.append(layer.regularizer_type)
        return samples[1:], means, covs, kls, kl_types

    # ---- plans ------------------------------------------------------------------------------
    MAX_ENGINES = 4       # plans own every buffer of a pass (c3: 1 GB of saved panels per training plan, c4: 27 GB)

    fast_reduce = False   # OPTIONAL reduced-precision fast path of the parameter contractions (engine.Engine); off = float64 parity path

    def engine(self, B, K, mode=None, world_size=1, rank=0):
        """The execution plan for (B, K, mode, world, rank, num_data); least-recently-used plans beyond MAX_ENGINES are
        dropped (their device buffers are freed) so that evaluating many different batch / test-set sizes cannot
        accumulate memory.  num_data is part of the key: the ELBO scale num_data / B is fixed inside a plan."""
        key = (int(B), int(K), mode or self._mode, int(world_size), int(rank), int(self.num_data), bool(self.fast_reduce))
        eng = self._engines.pop(key, None)
        if eng is None:
            eng = Engine(self, key[0], key[1], key[2], world_size, rank, fast_reduce=self.fast_reduce)
        self._engines[key] = eng                      # dicts keep insertion order: last = most recently used
        while len(self._engines) > self.MAX_ENGINES:
            self._engines.pop(next(iter(self._engines)))
        return eng

    def drop_engine(self, eng):
        for k in [k for k, v in self._engines.items() if v is eng]:
            del self._engines[k]

    def _next_batch(self):
        if self.minibatch is None:
            return self.X, self.Y
        idx = torch.as_tensor(self.minibatch.next(), device=self.X.device)
        return self.X[idx], self.Y[idx]

    # ---- objective --------------------------------------------------------------------------
    def _build_likelihood(self, X=None, Y=None, eps=None):
        """ELBO on the next minibatch (or the given one) as a 1-element device tensor; reference models.py:49-86
        (VI) / :112-150 (IW).  `eps`: optional injected noise, one entry per layer, point-major."""
        if X is None and self.minibatch is not None:
            idx = torch.as_tensor(self.minibatch.next(), device=self.X.device)
            eng = self.engine(len(idx), self.num_samples)
            eng.set_batch_indices(self.X, self.Y, idx)       # one gather launch (gpflow Minibatch, models.py:25-26)
        else:
            if X is None:
                X, Y = self.X, self.Y
            eng = self.engine(len(X), self.num_samples)
            eng.set_batch(X, Y)
        self._evals += 1
        eng.draw_noise(eps, seed=self.noise_seed, step=self._evals)
        self._last_engine = eng
        return eng.forward()

    def compute_log_likelihood(self, X=None, Y=None, eps=None):
        out = float(self._build_likelihood(X, Y, eps).item())
        self._last_engine.check_info()
        return out

    def compute_log_likelihood_and_grads(self, X=None, Y=None, eps=None):
        """(ELBO, {parameter name: d ELBO / d constrained value}) -- what tf.gradients gives the reference's
        optimisers (build_models.py:293-295), before the transform chain rule."""
        if X is None:
            X, Y = self._next_batch()
        eng = self.engine(len(X), self.num_samples)
        self._evals += 1
        loss = eng.elbo_and_grads(X, Y, eps, seed=self.noise_seed, step=self._evals)
        out = float(loss.item())
        eng.check_info()
        return out, {k: v.cpu().numpy() for k, v in FlatParams.of(self).grads_by_name().items()}

    # ---- prediction (reference models.py:89-107) --------------------------------------------
    def _predict_device(self, X, S, eps=None):
        """[S, N, Dy] mean and variance of the final layer as device tensors (views of the plan's buffers)."""
        X = np.asarray(X, dtype=np.float64)
        eng = self.engine(len(X), int(S), 'predict')
        eng.set_batch(X)
        self._evals += 1
        eng.draw_noise(eps, seed=self.noise_seed, step=self._evals)
        m, v = eng.forward()
        eng.check_info()
        return m.view(int(S), len(X), self.Dy), v.view(int(S), len(X), self.Dy)

    def predict_f_multisample(self, X, S, eps=None):
        m, v = self._predict_device(X, S, eps)
        return m.cpu().numpy(), v.cpu().numpy()

    def predict_y_samples(self, X, S, eps=None, eps_y=None):
        """reference models.py:99-107: y = f_mean + z sqrt(f_var + likelihood variance), z ~ N(0, 1) drawn on the device
        (iwvi_normal_fill, keyed by the evaluation count) unless injected as eps_y [S, N, Dy]."""
        from . import capi
        from .engine import layer_seed
        m, v = self._predict_device(X, S, eps)
        if eps_y is None:
            self._evals += 1
            z = torch.empty_like(m)
            capi.normal_fill(z, m.shape[0] * m.shape[1], self.Dy, 0, layer_seed(self.noise_seed, self._evals, 1 << 20))
        else:
            z = torch.as_tensor(np.asarray(eps_y, dtype=np.float64), dtype=settings.float_type).to(m.device).reshape(m.shape)
        lik = FlatParams.of(self).cview(self.likelihood.variance)
        return (m + z * torch.sqrt(v + lik)).cpu().numpy()

    def predict_f(self, X, eps=None):
        """GPModel.predict_f -> _build_predict(X, full_cov=False): one propagation of the 2-D inputs."""
        m, v = self.predict_f_multisample(X, 1, eps)
        return m[0], v[0]

    def predict_f_full_cov(self, X, eps=None):
        """GPModel.predict_f_full_cov -> _build_predict(X, full_cov=True).  Implemented for a single GPLayer (what
        reference tests/test_gp_layer.py:53-54 pins): mean [N, Dy], cov [Dy, N, N]."""
        from . import temp_workaround as tw
        if len(self.layers) != 1:
            raise NotImplementedError('full covariance over N is provided for single-layer models')
        layer = self.layers[0]
        Xt = torch.as_tensor(np.asarray(X, dtype=np.float64), dtype=settings.float_type).to(self.X.device)
        m, cov = tw.full_cov_conditional(layer, Xt)
        return m.cpu().numpy(), cov.cpu().numpy()


class DGP_IWVI(DGP_VI):
    """reference models.py:110-150: importance-weighted bound, points laid out data-major [N, K]."""
    _mode = 'iw'
