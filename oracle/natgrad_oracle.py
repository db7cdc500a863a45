"""TEST INFRASTRUCTURE ONLY -- CPU restatement of one gpflow.training.NatGradOptimizer step (GPflow 1.x, XiNat) as the
reference uses it on the last layer's (q_mu, q_sqrt) (experiments/build_models.py:284-300).  GPflow is not vendored in
/root/reference nor installable here; the restatement follows its published algorithm: the gradient with respect to the
expectation parameters is obtained by AUTOGRAD through eta -> (mu, L) = (eta1, chol(eta2 - eta1 eta1^T)) -- structurally
different from the closed form in dgps_with_iwvi_b200/natgrad.py, which it checks.  PARITY UNPINNED against GPflow itself;
pinned instead by the conjugate-model property (gamma = 1 reaches the optimal q(u) in one step, tests/test_oracle.py)."""
import torch


def natgrad_step(q_mu, q_sqrt, g_elbo_mu, g_elbo_sqrt, gamma):
    q_mu = torch.as_tensor(q_mu, dtype=torch.float64)
    q_sqrt = torch.tril(torch.as_tensor(q_sqrt, dtype=torch.float64))
    g_mu = -torch.as_tensor(g_elbo_mu, dtype=torch.float64)          # objective = -ELBO
    g_sqrt = -torch.tril(torch.as_tensor(g_elbo_sqrt, dtype=torch.float64))
    R = q_sqrt.shape[0]
    mu_new, L_new = [], []
    for r in range(R):
        mu, L = q_mu[:, r], q_sqrt[r]
        S = L @ L.t()
        eta1 = mu.clone().requires_grad_(True)
        eta2 = (S + torch.outer(mu, mu)).clone().requires_grad_(True)
        var = eta2 - torch.outer(eta1, eta1)
        Lc = torch.linalg.cholesky(0.5 * (var + var.t()))
        d1, d2 = torch.autograd.grad([eta1, Lc], [eta1, eta2], grad_outputs=[g_mu[:, r], g_sqrt[r]])
        d2 = 0.5 * (d2 + d2.t())
        Sinv = torch.linalg.inv(S)
        th1 = Sinv @ mu - gamma * d1
        th2 = -0.5 * Sinv - gamma * d2
        S_new = torch.linalg.inv(-2.0 * th2)
        S_new = 0.5 * (S_new + S_new.t())
        mu_new.append(S_new @ th1)
        L_new.append(torch.linalg.cholesky(S_new))
    return torch.stack(mu_new, 1), torch.stack(L_new, 0)
