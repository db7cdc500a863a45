"""
TEST INFRASTRUCTURE ONLY -- independent numpy/scipy closed form used to PIN oracle/iwvi_oracle.py.

The reference's only live numeric tests (tests/test_gp_layer.py:15-54 and :57-96) assert that a
single-GPLayer DGP_VI equals GPflow's SVGP.  GPflow cannot be imported here, so this file restates
what SVGP computes from the textbook *unwhitened* formulas (Hensman et al. 2013), deliberately with
different algebra from the oracle (explicit q(u) = N(m_u, S_u) with m_u = Lm q_mu,
S_u = Lm Lq Lq^T Lm^T; cho_solve instead of the whitened triangular route; direct pairwise distances
instead of the expanded -2XZ^T + norms form), so agreement is a real check:

    mean = Kfu Kuu^-1 m_u + mf(X)
    cov  = Kff - Kfu Kuu^-1 Kuf + Kfu Kuu^-1 S_u Kuu^-1 Kuf
    KL   = 0.5 ( tr(Kuu^-1 S_u) + m_u^T Kuu^-1 m_u - M + logdet Kuu - logdet S_u )   per output
    ELBO = sum_n E_q[log N(y_n | f_n, s2)] - KL
"""
import numpy as np
from scipy.linalg import cho_factor, cho_solve, cholesky
from scipy.spatial.distance import cdist


def kernel(kind, X, X2, variance, lengthscales):
    ls = np.broadcast_to(np.asarray(lengthscales, dtype=np.float64), (X.shape[1],))
    r = cdist(X / ls, X2 / ls, metric='euclidean')
    if kind == 'RBF':
        return variance * np.exp(-0.5 * r ** 2)
    if kind == 'Matern52':
        return variance * (1 + np.sqrt(5) * r + 5.0 / 3.0 * r ** 2) * np.exp(-np.sqrt(5) * r)
    if kind == 'Matern32':
        return variance * (1 + np.sqrt(3) * r) * np.exp(-np.sqrt(3) * r)
    if kind == 'Matern12':
        return variance * np.exp(-r)
    raise ValueError(kind)


def svgp(kind, variance, lengthscales, Z, q_mu, q_sqrt, X, Y, Xs, lik_variance, mf_A=None, mf_b=None,
         jitter=1e-6):
    """Returns (elbo on (X,Y) full batch, predictive mean [Ns,R], predictive full cov [R,Ns,Ns])."""
    M, R = q_mu.shape
    Kuu = kernel(kind, Z, Z, variance, lengthscales) + jitter * np.eye(M)
    Lm = cholesky(Kuu, lower=True)
    cf = cho_factor(Kuu, lower=True)
    logdet_Kuu = 2 * np.log(np.diag(Lm)).sum()

    def mf(X_):
        if mf_A is None:
            return np.zeros((X_.shape[0], 1))
        return X_ @ mf_A + (0.0 if mf_b is None else mf_b)

    def predict(X_):
        Kuf = kernel(kind, Z, X_, variance, lengthscales)
        Kff = kernel(kind, X_, X_, variance, lengthscales)
        KiKuf = cho_solve(cf, Kuf)                       # Kuu^-1 Kuf
        means, covs = [], []
        for r in range(R):
            Lq = np.tril(q_sqrt[r])
            m_u = Lm @ q_mu[:, r]
            S_u = Lm @ Lq @ Lq.T @ Lm.T
            means.append(KiKuf.T @ m_u)
            covs.append(Kff - Kuf.T @ KiKuf + KiKuf.T @ S_u @ KiKuf)
        return np.stack(means, 1) + mf(X_), np.stack(covs, 0)

    kl = 0.0
    for r in range(R):
        Lq = np.tril(q_sqrt[r])
        m_u = Lm @ q_mu[:, r]
        S_u = Lm @ Lq @ Lq.T @ Lm.T
        logdet_S = 2 * np.log(np.abs(np.diag(Lm))).sum() + 2 * np.log(np.abs(np.diag(Lq))).sum()
        kl += 0.5 * (np.trace(cho_solve(cf, S_u)) + m_u @ cho_solve(cf, m_u) - M + logdet_Kuu - logdet_S)

    mean, cov = predict(X)
    var = np.stack([np.diag(cov[r]) for r in range(R)], 1)
    ve = -0.5 * np.log(2 * np.pi) - 0.5 * np.log(lik_variance) - 0.5 * ((Y - mean) ** 2 + var) / lik_variance
    elbo = ve.sum() - kl
    ms, cs = predict(Xs)
    return elbo, ms, cs
