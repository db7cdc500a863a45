"""
TEST INFRASTRUCTURE ONLY -- a second, structurally independent restatement of the IW-specific half of the path
(reference dgps_with_iwvi/models.py:112-150, layers.py:72-105): a plain Python loop over every (n, k) pair, each
point pushed through the layer chain ON ITS OWN with the textbook UNWHITENED sparse-GP posterior (explicit Kuu^-1
through scipy cho_solve -- no Lm^-1 Kuf panel, no whitening), scipy.stats.norm.logpdf for every density and
scipy.special.logsumexp for the K-way reduction.  It shares no code and no intermediate quantity with
oracle/iwvi_oracle.py, so agreement pins the [N, K] tiling (models.py:113-116), the sampled local regulariser
log q(w) - log p(w) (layers.py:98-100), the Mok mixing (temp_workaround.py:142-145), the mean-function add on sample
AND mean (layers.py:46-48) and `reduce_logsumexp(L_NK, 1) - log K` (models.py:148) of the op-for-op oracle.

Only for small cases (pure-Python loops; N*K of a few dozen points).
"""
import numpy as np
from scipy.linalg import cho_factor, cho_solve
from scipy.special import logsumexp
from scipy.stats import norm

from .svgp_closed_form import kernel


def _encoder(Ws, bs, latent_dim, xy):
    """layers.py:137-152 for one row xy [Dx+Dy]: tanh MLP, skip after the activation where widths match,
    sigma = softplus(raw - 3)."""
    z = xy
    n = len(Ws)
    for i, (W, b) in enumerate(zip(Ws, bs)):
        z0 = z
        z = z @ W + b
        if i < n - 1:
            z = np.tanh(z)
        if W.shape[0] == W.shape[1]:
            z = z + z0
    mu, raw = z[:latent_dim], z[latent_dim:]
    return mu, np.logaddexp(0.0, raw - 3.0)


def _gp_point(ls, x):
    """Posterior of the R latent GPs at ONE input x [D]: mean [R], var [R], unwhitened:
        q(u_r) = N(m_r, S_r),  m_r = Lm q_mu_r,  S_r = Lm Lq_r Lq_r^T Lm^T          (the whitening of layers.py:42)
        mean_r = k^T Kuu^-1 m_r,   var_r = kxx - k^T Kuu^-1 k + k^T Kuu^-1 S_r Kuu^-1 k."""
    Z, q_mu, q_sqrt = ls['Z'], ls['q_mu'], ls['q_sqrt']
    M, R = q_mu.shape
    kind, var, lsc = ls['kern'], float(ls['variance']), np.asarray(ls['lengthscales'], dtype=np.float64)
    Kuu = kernel(kind, Z, Z, var, lsc) + ls.get('jitter', 1e-6) * np.eye(M)
    Lm = np.linalg.cholesky(Kuu)
    cf = cho_factor(Kuu, lower=True)
    k = kernel(kind, Z, x[None, :], var, lsc)[:, 0]
    kxx = kernel(kind, x[None, :], x[None, :], var, lsc)[0, 0]
    a = cho_solve(cf, k)                                  # Kuu^-1 k
    mean, v = np.zeros(R), np.zeros(R)
    for r in range(R):
        m_r = Lm @ q_mu[:, r]
        LS = Lm @ np.tril(q_sqrt[r])
        mean[r] = a @ m_r
        v[r] = kxx - k @ a + np.sum((LS.T @ a) ** 2)
    return mean, v


def iw_bound(spec, X, Y, eps):
    """IW-ELBO of models.py:112-150 by explicit loops.  eps[l]: [N, K, C] per layer (None for the final layer)."""
    N, K = X.shape[0], spec['num_samples']
    lik = float(spec['lik_variance'])
    L_NK = np.zeros((N, K))
    n_layers = len(spec['layers'])
    for n in range(N):
        for k in range(K):
            f = X[n].copy()                               # models.py:113: every k starts from the same row
            xy = np.concatenate([X[n], Y[n]])             # models.py:116
            local = 0.0
            for li, ls in enumerate(spec['layers']):
                if ls['type'] == 'lv':
                    mu, sigma = _encoder(ls['Ws'], ls['bs'], ls['latent_dim'], xy)          # layers.py:83
                    w = mu + eps[li][n, k] * sigma                                          # layers.py:86-87
                    local += np.sum(norm.logpdf(w, mu, sigma) - norm.logpdf(w, 0.0, 1.0))   # layers.py:98-100
                    f = np.concatenate([f, w])                                              # layers.py:89
                    continue
                gmean, gvar = _gp_point(ls, f)
                if ls['mf'] == 'Linear':
                    mf = f @ ls['mf_A'] + ls['mf_b']
                elif ls['mf'] == 'Identity':
                    mf = f
                else:
                    mf = 0.0
                if li == n_layers - 1:
                    W = ls.get('W')
                    m = (gmean if W is None else W @ gmean) + mf
                    v = gvar if W is None else (W ** 2) @ gvar
                    # E_{f ~ N(m, v)} log N(y | f, lik): models.py:134 -> Gaussian.variational_expectations
                    ve = norm.logpdf(Y[n], m, np.sqrt(lik)) - 0.5 * v / lik
                    L_NK[n, k] = np.sum(ve) - local                                          # models.py:138-142
                else:
                    s = gmean + eps[li][n, k] * np.sqrt(gvar)                                # temp_workaround.py:89-91
                    W = ls.get('W')
                    f = (s if W is None else W @ s) + mf                                     # :143, layers.py:47
    scale = float(spec['num_data']) / N                                                      # models.py:144-145
    logp = logsumexp(L_NK, axis=1) - np.log(K)                                               # models.py:148
    kl = 0.0
    for ls in spec['layers']:
        if ls['type'] != 'gp':
            continue
        # textbook KL[N(q_mu_r, Lq Lq^T) || N(0, I)] summed over outputs (whitened prior, layers.py:44)
        M, R = ls['q_mu'].shape
        for r in range(R):
            Lq = np.tril(ls['q_sqrt'][r])
            S = Lq @ Lq.T
            kl += 0.5 * (np.trace(S) + ls['q_mu'][:, r] @ ls['q_mu'][:, r] - M - np.linalg.slogdet(S)[1])
    return float(np.sum(logp) * scale - kl), L_NK                                           # models.py:150
