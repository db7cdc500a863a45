"""
TEST INFRASTRUCTURE ONLY -- CPU float64 oracle for the IW-ELBO hot path of DGPs_with_IWVI.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (dgps_with_iwvi_b200) never imports it and has no CPU fallback.

What it is: an op-for-op restatement, in plain torch float64 on the CPU, of

  * dgps_with_iwvi/temp_workaround.py:12-98    independent_multisample_sample_conditional
  * dgps_with_iwvi/temp_workaround.py:118-161  multisample_sample_conditional (+ SharedMixedMok :107-115)
  * dgps_with_iwvi/temp_workaround.py:167-188  gauss_kl wrapper
  * dgps_with_iwvi/layers.py:35-50             GPLayer.propagate
  * dgps_with_iwvi/layers.py:72-105            LatentVariableLayer.propagate
  * dgps_with_iwvi/layers.py:137-152           Encoder.__call__
  * dgps_with_iwvi/models.py:30-46             DGP_VI.propagate
  * dgps_with_iwvi/models.py:49-86             DGP_VI._build_likelihood
  * dgps_with_iwvi/models.py:112-150           DGP_IWVI._build_likelihood
  * dgps_with_iwvi/models.py:89-107            _build_predict / predict_f_multisample / predict_y_samples

Gradients come from torch.autograd (the reference uses tf.gradients).

Third-party arithmetic.  The reference delegates kernels, Kuu/Kuf, gauss_kl, the Gaussian likelihood,
mean functions and parameter transforms to GPflow 1.x (un-pinned, ~1.3) on TensorFlow 1.x (un-pinned,
~1.12).  Neither is vendored in /root/reference nor installable here (no wheels for Python 3.12, no
network), so their published formulas are restated below and each restated
function names the reference call site it serves.

PARITY UNPINNED for everything IW-specific: the reference ships no golden vectors, and its only live
numeric tests (tests/test_gp_layer.py:15-96) compare DGP_VI against GPflow's SVGP, which cannot be
imported.  What pins this oracle instead (tests/test_oracle.py):
  (1) an independent, unwhitened numpy/scipy SVGP closed form (oracle/svgp_closed_form.py) reproducing
      tests/test_gp_layer.py:15-54 (Matern52, Linear mean function, full covariance) and :57-96;
  (2) diag(full_cov) == diag path; whitened KL against the textbook KL formula;
  (3) the statistical properties asserted by the commented-out reference tests
      (tests/test_latent_var_layer.py:120-241): K=1 E[IW]=E[VI], sd(VI)<sd(IW), K>1 E[IW]>E[VI];
  (4) torch.autograd.gradcheck in float64.
Golden vectors under tests/golden/ are generated FROM this oracle by tests/golden/make_golden.py.

Noise is always injected explicitly (the reference draws it with unseeded tf.random_normal at
layers.py:86 and temp_workaround.py:89); index order is the reference's: eps_w [N,K,Lw], and per
Mok GP layer eps [N,K,L] in latent-GP space *before* mixing.
"""
import math

import torch

DTYPE = torch.float64
JITTER = 1e-6  # gpflow.settings.numerics.jitter_level default (temp_workaround.py:39)


# ----------------------------------------------------------------------------------------------
# GPflow 1.x numerics restated
# ----------------------------------------------------------------------------------------------

def positive_forward(x):
    """gpflow.transforms.positive (Log1pe): theta = softplus(x) + 1e-6. Used by kernel variance /
    lengthscales and likelihood variance (call sites build_models.py:198-199,212,238)."""
    return torch.nn.functional.softplus(x) + 1e-6


def positive_backward(theta):
    """Inverse of positive_forward: x = log(expm1(theta - 1e-6)), computed stably."""
    y = theta - 1e-6
    return y + torch.log(-torch.expm1(-y))


def square_dist(X, X2, lengthscales):
    """GPflow Stationary._scaled_square_dist: expanded form, NOT clamped at zero.
    Batches over leading axes of X when X2 is None (relied on at temp_workaround.py:45)."""
    X = X / lengthscales
    Xs = (X ** 2).sum(-1)
    if X2 is None:
        dist = -2.0 * X @ X.transpose(-1, -2)
        dist = dist + Xs[..., :, None] + Xs[..., None, :]
        return dist
    X2 = X2 / lengthscales
    X2s = (X2 ** 2).sum(-1)
    dist = -2.0 * X @ X2.transpose(-1, -2)
    dist = dist + Xs[..., :, None] + X2s[..., None, :]
    return dist


class Kern:
    """Stationary kernel descriptor: kind in {'RBF','Matern52','Matern32','Matern12'};
    variance scalar tensor, lengthscales scalar or [D] tensor (constrained values)."""

    def __init__(self, kind, variance, lengthscales, active_dims=None):
        self.kind = kind
        self.variance = variance
        self.lengthscales = lengthscales
        self.active_dims = active_dims      # gpflow Kern._slice: the kernel acts on these input columns only

    def K(self, X, X2=None):
        if self.active_dims is not None:
            X = X[..., self.active_dims]
            X2 = None if X2 is None else X2[..., self.active_dims]
        r2 = square_dist(X, X2, self.lengthscales)
        if self.kind == 'RBF':
            return self.variance * torch.exp(-r2 / 2.0)
        r = torch.sqrt(torch.clamp(r2, min=1e-40))  # GPflow: sqrt(max(r2, 1e-40)); zero grad where clamped
        if self.kind == 'Matern52':
            s5 = math.sqrt(5.0)
            return self.variance * (1.0 + s5 * r + 5.0 / 3.0 * r ** 2) * torch.exp(-s5 * r)
        if self.kind == 'Matern32':
            s3 = math.sqrt(3.0)
            return self.variance * (1.0 + s3 * r) * torch.exp(-s3 * r)
        if self.kind == 'Matern12':
            return self.variance * torch.exp(-r)
        raise ValueError(self.kind)

    def Kdiag(self, X):
        """fill(shape(X)[:-1], variance)"""
        return self.variance * torch.ones(X.shape[:-1], dtype=X.dtype)


def Kuu(Z, kern, jitter):
    """gpflow.features.Kuu(InducingPoints): kern.K(Z) + jitter * I  (temp_workaround.py:39)."""
    return kern.K(Z) + jitter * torch.eye(Z.shape[0], dtype=Z.dtype)


def Kuf(Z, kern, Xnew):
    """gpflow.features.Kuf(InducingPoints): kern.K(Z, Xnew) -> [M, N]  (temp_workaround.py:44)."""
    return kern.K(Z, Xnew)


def gauss_kl_white(q_mu, q_sqrt):
    """gpflow.kullback_leiblers.gauss_kl with K=None (whitened), q_mu [M,R], q_sqrt [R,M,M] or [M,R].
    0.5 * (sum q_mu^2 - M*R - sum log diag(Lq)^2 + sum Lq^2)   (call site temp_workaround.py:188)."""
    M, R = q_mu.shape
    if q_sqrt.dim() == 2:
        Lq_diag = q_sqrt
        trace = (q_sqrt ** 2).sum()
    else:
        Lq = torch.tril(q_sqrt)  # tf.matrix_band_part(q_sqrt, -1, 0)
        Lq_diag = torch.diagonal(Lq, dim1=-2, dim2=-1)
        trace = (Lq ** 2).sum()
    mahalanobis = (q_mu ** 2).sum()
    constant = -float(M * R)
    logdet_qcov = torch.log(Lq_diag ** 2).sum()
    return 0.5 * (mahalanobis + constant - logdet_qcov + trace)


def multivariate_normal(x, mu, L):
    """gpflow.logdensities.multivariate_normal: one log-density per column of x [M,R]
    (call site temp_workaround.py:184)."""
    d = x - mu
    alpha = torch.linalg.solve_triangular(L, d, upper=False)
    M = x.shape[0]
    return (-0.5 * (alpha ** 2).sum(0) - 0.5 * M * math.log(2 * math.pi)
            - torch.log(torch.diagonal(L)).sum())


def gauss_kl(q_mu, q_sqrt, K=None):
    """temp_workaround.py:167-188 wrapper: q_sqrt None -> negative log prob (SGHMC case)."""
    if q_sqrt is None:
        M = q_mu.shape[0]
        I = torch.eye(M, dtype=q_mu.dtype)
        L = I if K is None else torch.linalg.cholesky(K + I * JITTER)
        return -multivariate_normal(q_mu, torch.zeros_like(q_mu), L).sum()
    assert K is None, "only the whitened KL is on the reference's path (layers.py:44)"
    return gauss_kl_white(q_mu, q_sqrt)


def gaussian_variational_expectations(Fmu, Fvar, Y, lik_variance):
    """gpflow.likelihoods.Gaussian.variational_expectations (call sites models.py:66,134)."""
    return (-0.5 * math.log(2 * math.pi) - 0.5 * torch.log(lik_variance)
            - 0.5 * ((Y - Fmu) ** 2 + Fvar) / lik_variance)


def gaussian_predict_mean_and_var(Fmu, Fvar, lik_variance):
    """gpflow.likelihoods.Gaussian.predict_mean_and_var (call site models.py:105)."""
    return Fmu, Fvar + lik_variance


def normal_log_prob(x, mu, sigma):
    """tf.contrib.distributions.Normal(mu, sigma).log_prob(x) (layers.py:98-99)."""
    return -0.5 * ((x - mu) / sigma) ** 2 - 0.5 * math.log(2 * math.pi) - torch.log(sigma)


def normal_kl_to_standard(mu, sigma):
    """tf.contrib.distributions.kl_divergence(Normal(mu,sigma), Normal(0,1)) (layers.py:103)."""
    return 0.5 * mu ** 2 + 0.5 * (sigma ** 2 - 1.0 - torch.log(sigma ** 2))


class MeanFunction:
    """gpflow.mean_functions Linear / Identity / Zero (call site layers.py:46)."""

    def __init__(self, kind, A=None, b=None):
        self.kind, self.A, self.b = kind, A, b

    def __call__(self, F):
        if self.kind == 'Zero':
            return torch.zeros(F.shape[:-1] + (1,), dtype=F.dtype)  # broadcast add
        if self.kind == 'Identity':
            return F
        if self.kind == 'Linear':
            return F @ self.A + self.b
        raise ValueError(self.kind)


# ----------------------------------------------------------------------------------------------
# temp_workaround.py restated
# ----------------------------------------------------------------------------------------------

def independent_multisample_sample_conditional(Xnew, Z, kern, f, *, full_cov=False, full_output_cov=False,
                                               q_sqrt=None, white=False, eps=None, jitter=JITTER, eps_joint=None):
    """temp_workaround.py:12-98.  Xnew [S,N,D]; Z [M,D]; f [M,R]; q_sqrt [R,M,M] | [M,R] | None.
    eps [S,N,R] replaces tf.random_normal at :89.  Returns sample [S,N,R], fmean [S,N,R],
    fvar [S,N,R] (diag) or [S,R,N,N] (full_cov).  With full_cov=True the sample is None: the
    reference's full-cov sampler (:93-96) has a shape bug and is dead code in training.
    eps_joint [S,R,N] (the [S,R,N,1] draw of :94) asks for the joint draw those lines intend:
    sample[s,:,r] = fmean[s,:,r] + chol(fvar[s,r]) z[s,r]  (no jitter, :95)."""
    if full_output_cov:
        raise NotImplementedError
    Kmm = Kuu(Z, kern, jitter)                                        # :39
    S, N, D = Xnew.shape                                              # :41
    M = Kmm.shape[0]
    Kmn_M_SN = Kuf(Z, kern, Xnew.reshape(S * N, D))                   # :44
    Knn = kern.K(Xnew) if full_cov else kern.Kdiag(Xnew)              # :45
    num_func = f.shape[1]
    Lm = torch.linalg.cholesky(Kmm)                                   # :48
    A_M_SN = torch.linalg.solve_triangular(Lm, Kmn_M_SN, upper=False)  # :51
    A = A_M_SN.reshape(M, S, N).permute(1, 0, 2)                      # :52  S x M x N
    if full_cov:
        fvar = Knn - A.transpose(-1, -2) @ A                          # :56
        fvar = fvar[:, None, :, :].repeat(1, num_func, 1, 1)          # :57
    else:
        fvar = Knn - (A ** 2).sum(-2)                                 # :59
        fvar = fvar[:, None, :].repeat(1, num_func, 1)                # :60
    if not white:
        A_M_SN = torch.linalg.solve_triangular(Lm.t(), A_M_SN, upper=True)   # :64
        A = A_M_SN.reshape(M, S, N).permute(1, 0, 2)                  # :65
    fmean = A.transpose(-1, -2) @ f[None, :, :].repeat(S, 1, 1)        # :68
    if q_sqrt is not None:
        if q_sqrt.dim() == 2:
            LTA = A[:, None, :, :] * q_sqrt.t()[None, :, :, None]     # :73
        elif q_sqrt.dim() == 3:
            LTA = torch.einsum('rMm,sMn->srmn', torch.tril(q_sqrt), A)  # :78 (materialised on purpose)
        else:
            raise ValueError("Bad dimension for q_sqrt: %s" % str(q_sqrt.dim()))
        if full_cov:
            fvar = fvar + LTA.transpose(-1, -2) @ LTA                 # :83
        else:
            fvar = fvar + (LTA ** 2).sum(2)                           # :85
    if not full_cov:
        fvar = fvar.transpose(-1, -2)                                 # :90
        sample = None if eps is None else fmean + eps * fvar ** 0.5   # :91
    elif eps_joint is None:
        sample = None                                                 # :93-96 dead / buggy in the reference
    else:
        z = eps_joint.reshape(S, num_func, N, 1)                      # :93-94
        sample_SRN1 = fmean.transpose(1, 2)[..., None] + torch.linalg.cholesky(fvar) @ z   # :95 with the mean transposed
        sample = sample_SRN1[..., 0].transpose(1, 2)                  # :96
    return sample, fmean, fvar


def sample_conditional_2d(Xnew, Z, kern, f, *, full_cov=False, q_sqrt=None, white=False, eps=None,
                          jitter=JITTER):
    """gpflow.conditionals.sample_conditional on 2-D inputs (temp_workaround.py:134,157): the same
    algebra as above without the leading S axis (GPflow base_conditional).
    Xnew [N,D] -> sample/mean [N,R], var [N,R] or [R,N,N].  With full_cov GPflow's _sample_mvn draws
    jointly over N: mean + chol(cov + jitter*I) eps  (GPflow adds jitter there)."""
    _, m, v = independent_multisample_sample_conditional(
        Xnew[None], Z, kern, f, full_cov=full_cov, q_sqrt=q_sqrt, white=white, eps=None, jitter=jitter)
    m, v = m[0], v[0]
    if eps is None:
        return None, m, v
    if full_cov:
        N = Xnew.shape[0]
        chol = torch.linalg.cholesky(v + jitter * torch.eye(N, dtype=v.dtype))   # [R,N,N]
        s = m + (chol @ eps.t()[:, :, None])[:, :, 0].t()
    else:
        s = m + eps * v ** 0.5
    return s, m, v


class Mok:
    """SharedMixedMok (temp_workaround.py:107-115): shared kernel + mixing matrix W [P, L]."""

    def __init__(self, kernel, W):
        self.kernel, self.W = kernel, W


def multisample_sample_conditional(Xnew, Z, kern, f, *, full_cov=False, full_output_cov=False,
                                   q_sqrt=None, white=False, eps=None, jitter=JITTER):
    """temp_workaround.py:118-161."""
    if isinstance(kern, Mok):
        if Xnew.dim() == 3:
            sample, gmean, gvar = independent_multisample_sample_conditional(
                Xnew, Z, kern.kernel, f, white=white, q_sqrt=q_sqrt, full_output_cov=False,
                full_cov=False, eps=eps, jitter=jitter)                               # :125-129 (full_cov forced off)
        else:
            sample, gmean, gvar = sample_conditional_2d(
                Xnew, Z, kern.kernel, f, white=white, q_sqrt=q_sqrt, full_cov=False, eps=eps,
                jitter=jitter)                                                          # :134-138
        f_sample = None if sample is None else sample @ kern.W.t()                      # :143
        f_mu = gmean @ kern.W.t()                                                       # :144
        f_var = gvar @ (kern.W ** 2).t()                                                # :145
        return f_sample, f_mu, f_var
    if Xnew.dim() == 3:
        return independent_multisample_sample_conditional(
            Xnew, Z, kern, f, full_cov=full_cov, full_output_cov=full_output_cov, q_sqrt=q_sqrt,
            white=white, eps=eps, jitter=jitter)                                        # :151-155
    return sample_conditional_2d(Xnew, Z, kern, f, full_cov=full_cov, q_sqrt=q_sqrt, white=white,
                                 eps=eps, jitter=jitter)                                # :157-161


# ----------------------------------------------------------------------------------------------
# layers.py restated
# ----------------------------------------------------------------------------------------------

LOCAL, GLOBAL = 0, 1  # layers.py:9-11 RegularizerType


class GPLayer:
    """layers.py:14-50. Holds constrained values: q_mu [M,R], q_sqrt [R,M,M] (tril taken on use)."""
    regularizer_type = GLOBAL

    def __init__(self, kern, Z, q_mu, q_sqrt, mean_function=None, jitter=JITTER):
        self.kern, self.Z, self.q_mu, self.q_sqrt = kern, Z, q_mu, q_sqrt
        self.mean_function = mean_function or MeanFunction('Zero')
        self.jitter = jitter

    def num_noise(self):
        return self.q_mu.shape[1]

    def propagate(self, F, full_cov=False, eps=None, **kwargs):
        samples, mean, cov = multisample_sample_conditional(
            F, self.Z, self.kern, self.q_mu, full_cov=full_cov, q_sqrt=self.q_sqrt, white=True,
            eps=eps, jitter=self.jitter)                                                # :36-42
        kl = gauss_kl(self.q_mu, self.q_sqrt)                                           # :44
        mf = self.mean_function(F)                                                      # :46
        samples = None if samples is None else samples + mf                             # :47
        mean = mean + mf                                                                # :48
        return samples, mean, cov, kl


class Encoder:
    """layers.py:108-152. Ws[i] [d_i, d_{i+1}], bs[i] [d_{i+1}]; tanh; skip when dims match; softplus(.-3)."""

    ACTS = {'tanh': torch.tanh, 'relu': torch.relu, 'sigmoid': torch.sigmoid,
            'softplus': torch.nn.functional.softplus, 'elu': torch.nn.functional.elu, 'identity': lambda z: z}

    def __init__(self, Ws, bs, latent_dim, activation='tanh'):
        self.Ws, self.bs, self.latent_dim = Ws, bs, latent_dim
        self.activation = self.ACTS[activation]             # layers.py:122: activation_func or tf.nn.tanh

    def __call__(self, Z):
        n = len(self.bs)
        for i, (W, b) in enumerate(zip(self.Ws, self.bs)):
            dim_in, dim_out = W.shape
            Z0 = Z
            Z = Z @ W + b                                   # :141
            if i < n - 1:
                Z = self.activation(Z)                      # :143-144
            if dim_out == dim_in:
                Z = Z + Z0                                  # :146-147 skip AFTER the activation
        means, log_chol_diag = torch.split(Z, self.latent_dim, dim=-1)    # :149
        q_sqrt = torch.nn.functional.softplus(log_chol_diag - 3.0)         # :150
        return means, q_sqrt


class LatentVariableLayer:
    """layers.py:53-105."""
    regularizer_type = LOCAL

    def __init__(self, latent_dim, encoder):
        self.latent_dim, self.encoder = latent_dim, encoder

    def num_noise(self):
        return self.latent_dim

    def propagate(self, F, inference_amorization_inputs=None, is_sampled_local_regularizer=False,
                  eps=None, q_mu_feed=None, q_sqrt_feed=None, **kwargs):
        if inference_amorization_inputs is None:
            shape = F.shape[:-1] + (self.latent_dim,)                                   # :78
            ones = torch.ones(shape, dtype=F.dtype)
            q_mu = (torch.zeros(1, 1, dtype=F.dtype) if q_mu_feed is None else q_mu_feed) * ones      # :80
            q_sqrt = (torch.ones(1, 1, dtype=F.dtype) if q_sqrt_feed is None else q_sqrt_feed) * ones  # :81
        else:
            q_mu, q_sqrt = self.encoder(inference_amorization_inputs)                   # :83
        W = q_mu + eps * q_sqrt                                                         # :86-87
        samples = torch.cat([F, W], -1)                                                 # :89
        mean = torch.cat([F, q_mu], -1)                                                 # :90
        cov = torch.cat([torch.zeros_like(F), q_sqrt ** 2], -1)                          # :91
        if is_sampled_local_regularizer:
            zero = torch.zeros((), dtype=F.dtype)
            one = torch.ones((), dtype=F.dtype)
            kl = normal_log_prob(W, q_mu, q_sqrt) - normal_log_prob(W, zero, one)       # :100
        else:
            kl = normal_kl_to_standard(q_mu, q_sqrt)                                    # :103
        return samples, mean, cov, kl


# ----------------------------------------------------------------------------------------------
# models.py restated
# ----------------------------------------------------------------------------------------------

class DGP:
    """DGP_VI / DGP_IWVI (models.py:9-150) on explicit minibatches and explicit noise."""

    def __init__(self, layers, lik_variance, num_data, num_samples=1):
        self.layers, self.lik_variance = layers, lik_variance
        self.num_data, self.num_samples = num_data, num_samples

    def noise_shapes(self, lead_shape):
        """Shapes of the noise tensors propagate() consumes, one per layer (None where no draw is
        executed: a plain-kernel GP layer under full_cov=True, i.e. the IW final layer)."""
        return [tuple(lead_shape) + (l.num_noise(),) for l in self.layers]

    def propagate(self, X, eps, full_cov=False, inference_amorization_inputs=None,
                  is_sampled_local_regularizer=False):
        """models.py:30-46. eps: list (one entry per layer; entries may be None)."""
        samples, means, covs, kls, kl_types = [X], [], [], [], []
        for layer, e in zip(self.layers, eps):
            sample, mean, cov, kl = layer.propagate(
                samples[-1], full_cov=full_cov,
                inference_amorization_inputs=inference_amorization_inputs,
                is_sampled_local_regularizer=is_sampled_local_regularizer, eps=e)
            samples.append(sample); means.append(mean); covs.append(cov); kls.append(kl)
            kl_types.append(layer.regularizer_type)
        return samples[1:], means, covs, kls, kl_types

    def vi_likelihood(self, X, Y, eps):
        """DGP_VI._build_likelihood, models.py:49-86. eps[l] has shape [S*N, .] (sample-major)."""
        S = self.num_samples
        X_tiled = X.repeat(S, 1)                                                        # :50
        Y_tiled = Y.repeat(S, 1)                                                        # :51
        XY = torch.cat([X_tiled, Y_tiled], -1)                                          # :53
        samples, means, covs, kls, kl_types = self.propagate(
            X_tiled, eps, full_cov=False, inference_amorization_inputs=XY,
            is_sampled_local_regularizer=False)                                         # :58-61
        local_kls = [kl for kl, t in zip(kls, kl_types) if t == LOCAL]
        global_kls = [kl for kl, t in zip(kls, kl_types) if t == GLOBAL]
        var_exp = gaussian_variational_expectations(means[-1], covs[-1], Y_tiled, self.lik_variance)  # :66
        L_SN = var_exp.sum(-1)                                                          # :69
        L_S_N = L_SN.reshape(S, X.shape[0])                                             # :71-72
        if len(local_kls) > 0:
            local_kls_SN = torch.cat(local_kls, -1).sum(-1)                             # :75-76
            L_S_N = L_S_N - local_kls_SN.reshape(S, X.shape[0])                         # :77-78
        scale = float(self.num_data) / float(X.shape[0])                                # :80-81
        logp = L_S_N.mean(0)                                                            # :84
        return logp.sum() * scale - sum(global_kls)                                     # :86

    def iw_likelihood(self, X, Y, eps, reference_style=True, return_parts=False):
        """DGP_IWVI._build_likelihood, models.py:112-150. eps[l] has shape [N, K, .] (data-major).
        reference_style=True keeps the final layer's full KxK covariance and takes its diagonal
        (models.py:123,133); False takes the algebraically identical diag path."""
        K = self.num_samples
        X_tiled = X[:, None, :].repeat(1, K, 1)                                         # :113
        Y_tiled = Y[:, None, :].repeat(1, K, 1)                                         # :114
        XY = torch.cat([X_tiled, Y_tiled], -1)                                          # :116
        samples, means, covs, kls, kl_types = self.propagate(
            X_tiled, eps, full_cov=reference_style, inference_amorization_inputs=XY,
            is_sampled_local_regularizer=True)                                          # :122-125
        local_kls = [kl for kl, t in zip(kls, kl_types) if t == LOCAL]
        global_kls = [kl for kl, t in zip(kls, kl_types) if t == GLOBAL]
        if reference_style and covs[-1].dim() == 4:
            cov_diag = torch.diagonal(covs[-1], dim1=-2, dim2=-1).permute(0, 2, 1)      # :133
        else:
            cov_diag = covs[-1]
        var_exp = gaussian_variational_expectations(means[-1], cov_diag, Y_tiled, self.lik_variance)  # :134
        L_NK = var_exp.sum(2)                                                           # :138
        if len(local_kls) > 0:
            L_NK = L_NK - torch.cat(local_kls, -1).sum(2)                               # :140-142
        scale = float(self.num_data) / float(X.shape[0])                                # :144-145
        logp = torch.logsumexp(L_NK, 1) - math.log(K)                                   # :148
        elbo = logp.sum() * scale - sum(global_kls)                                     # :150
        if return_parts:
            return elbo, dict(L_NK=L_NK, logp=logp, samples=samples, means=means, covs=covs, kls=kls)
        return elbo

    def predict_f_multisample(self, X, S, eps):
        """models.py:93-98. eps[l] has shape [S, N, .]; LV layers sample from the prior."""
        X_tiled = X[None, :, :].repeat(S, 1, 1)
        _, means, covs, _, _ = self.propagate(X_tiled, eps)
        return means[-1], covs[-1]

    def predict_y_samples(self, X, S, eps, eps_y):
        """models.py:100-107."""
        m, v = self.predict_f_multisample(X, S, eps)
        m, v = gaussian_predict_mean_and_var(m, v, self.lik_variance)
        return m + eps_y * v ** 0.5

    def predict_f(self, X, eps, full_cov=False):
        """GPModel.predict_f / predict_f_full_cov -> _build_predict (models.py:89-91), 2-D inputs."""
        _, means, covs, _, _ = self.propagate(X, eps, full_cov=full_cov)
        return means[-1], covs[-1]


# ----------------------------------------------------------------------------------------------
# spec <-> oracle objects (plain dict of tensors; the product exports the same dict)
# ----------------------------------------------------------------------------------------------

def build_from_spec(spec, requires_grad=False):
    """spec = {'num_data', 'num_samples', 'lik_variance', 'layers': [ {...}, ... ]}
    GP layer:  {'type':'gp','kern':kind,'variance','lengthscales','Z','q_mu','q_sqrt',
                'W' (or None), 'mf': 'Zero'|'Identity'|'Linear', 'mf_A','mf_b', 'jitter'}
    LV layer:  {'type':'lv','latent_dim','Ws':[...],'bs':[...]}
    Returns (DGP, leaves) where leaves is a flat dict name -> tensor (grad-enabled if asked)."""
    leaves = {}

    def leaf(name, v):
        if v is None:
            return None
        t = torch.as_tensor(v, dtype=DTYPE).detach().clone()
        if requires_grad:
            t.requires_grad_(True)
        leaves[name] = t
        return t

    layers = []
    for i, ls in enumerate(spec['layers']):
        p = 'layers.%d.' % i
        if ls['type'] == 'lv':
            Ws = [leaf(p + 'encoder.Ws.%d' % j, w) for j, w in enumerate(ls['Ws'])]
            bs = [leaf(p + 'encoder.bs.%d' % j, b) for j, b in enumerate(ls['bs'])]
            layers.append(LatentVariableLayer(ls['latent_dim'], Encoder(Ws, bs, ls['latent_dim'],
                                                                        ls.get('activation', 'tanh'))))
        else:
            kern = Kern(ls['kern'], leaf(p + 'kern.variance', ls['variance']),
                        leaf(p + 'kern.lengthscales', ls['lengthscales']), ls.get('active_dims'))
            if ls.get('W') is not None:
                kern = Mok(kern, leaf(p + 'kern.W', ls['W']))
            mf = MeanFunction(ls['mf'], leaf(p + 'mf.A', ls.get('mf_A')), leaf(p + 'mf.b', ls.get('mf_b')))
            layers.append(GPLayer(kern, leaf(p + 'Z', ls['Z']), leaf(p + 'q_mu', ls['q_mu']),
                                  leaf(p + 'q_sqrt', ls['q_sqrt']), mf, jitter=ls.get('jitter', JITTER)))
    lik_var = leaf('likelihood.variance', spec['lik_variance'])
    return DGP(layers, lik_var, spec['num_data'], spec['num_samples']), leaves


def iw_elbo_and_grads(spec, X, Y, eps, reference_style=True):
    """ELBO and d ELBO / d (constrained parameter) for every leaf (zeros where a leaf is unused)."""
    model, leaves = build_from_spec(spec, requires_grad=True)
    X = torch.as_tensor(X, dtype=DTYPE); Y = torch.as_tensor(Y, dtype=DTYPE)
    eps = [None if e is None else torch.as_tensor(e, dtype=DTYPE) for e in eps]
    elbo = model.iw_likelihood(X, Y, eps, reference_style=reference_style)
    names = list(leaves)
    grads = torch.autograd.grad(elbo, [leaves[n] for n in names], allow_unused=True)
    out = {}
    for n, g in zip(names, grads):
        out[n] = torch.zeros_like(leaves[n]) if g is None else g
    return elbo.detach(), out


def vi_elbo_and_grads(spec, X, Y, eps):
    model, leaves = build_from_spec(spec, requires_grad=True)
    X = torch.as_tensor(X, dtype=DTYPE); Y = torch.as_tensor(Y, dtype=DTYPE)
    eps = [None if e is None else torch.as_tensor(e, dtype=DTYPE) for e in eps]
    elbo = model.vi_likelihood(X, Y, eps)
    names = list(leaves)
    grads = torch.autograd.grad(elbo, [leaves[n] for n in names], allow_unused=True)
    out = {}
    for n, g in zip(names, grads):
        out[n] = torch.zeros_like(leaves[n]) if g is None else g
    return elbo.detach(), out
