"""Natural-gradient step on a GPLayer's (q_mu, q_sqrt) -- the NatGrad half of the reference's training iteration
(reference experiments/build_models.py:284-300: gpflow.training.NatGradOptimizer(gamma) on the last layer's variational
parameters, default XiNat parameterisation).

For every output r with mean mu [M], covariance S = L L^T (L = tril(q_sqrt_r)):
    natural parameters      theta1 = S^-1 mu,        theta2 = -1/2 S^-1
    expectation parameters  eta1 = mu,               eta2 = S + mu mu^T
    step                    theta <- theta - gamma * d(-ELBO)/d eta
    d/d eta from d/d(mu, L) by the chain rule through mu = eta1, S = eta2 - eta1 eta1^T, L = chol(S):
        Sbar   = sym( L^-T Phi(L^T Lbar) L^-1 )                     (TF CholeskyGrad; Phi = tril with halved diagonal)
        deta2  = Sbar,     deta1 = mubar - 2 Sbar mu
    back    S = (-2 theta2)^-1,  mu = S theta1,  L = chol(S)

This is M x M dense linear algebra ONCE per step on R (= 1 for the last layer) matrices -- not part of the per-point hot
path -- and uses torch.linalg on the GPU.  The gradients it consumes (d ELBO / d q_mu, d ELBO / d q_sqrt) are outputs of
the hand-written backward kernels."""
import torch


def _phi(A):
    """tril with halved diagonal."""
    return torch.tril(A) - 0.5 * torch.diag_embed(torch.diagonal(A, dim1=-2, dim2=-1))


def expectation_gradients(q_mu, q_sqrt, g_mu, g_sqrt):
    """(dF/d eta1 [R, M], dF/d eta2 [R, M, M]) from (dF/d q_mu [M, R], dF/d q_sqrt [R, M, M]) for any scalar F."""
    L = torch.tril(q_sqrt)
    Lbar = torch.tril(g_sqrt)
    P = _phi(L.transpose(-1, -2) @ Lbar)
    # Sbar = L^-T P L^-1
    X = torch.linalg.solve_triangular(L.transpose(-1, -2), P, upper=True)            # L^-T P
    Sbar = torch.linalg.solve_triangular(L.transpose(-1, -2), X.transpose(-1, -2), upper=True).transpose(-1, -2)
    Sbar = 0.5 * (Sbar + Sbar.transpose(-1, -2))
    mu = q_mu.t()                                                                     # [R, M]
    d1 = g_mu.t() - 2.0 * (Sbar @ mu[..., None])[..., 0]
    return d1, Sbar


def natgrad_step(q_mu, q_sqrt, g_elbo_mu, g_elbo_sqrt, gamma):
    """One XiNat natural-gradient step maximising the ELBO.  q_mu [M, R], q_sqrt [R, M, M] (constrained values),
    g_elbo_* = d ELBO / d (constrained value).  Returns (q_mu_new, q_sqrt_new)."""
    L = torch.tril(q_sqrt)
    M = L.shape[-1]
    eye = torch.eye(M, dtype=L.dtype, device=L.device).expand_as(L)
    Linv = torch.linalg.solve_triangular(L, eye, upper=False)
    Sinv = Linv.transpose(-1, -2) @ Linv
    mu = q_mu.t()
    theta1 = (Sinv @ mu[..., None])[..., 0]
    theta2 = -0.5 * Sinv
    d1, d2 = expectation_gradients(q_mu, q_sqrt, -g_elbo_mu, -g_elbo_sqrt)            # the optimiser minimises -ELBO
    theta1 = theta1 - gamma * d1
    theta2 = theta2 - gamma * d2
    prec = -2.0 * theta2
    prec = 0.5 * (prec + prec.transpose(-1, -2))
    C = torch.linalg.cholesky(prec)
    S = torch.cholesky_inverse(C)
    mu_new = (S @ theta1[..., None])[..., 0]
    L_new = torch.linalg.cholesky(0.5 * (S + S.transpose(-1, -2)))
    return mu_new.t().contiguous(), L_new.contiguous()
