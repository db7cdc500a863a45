"""The stage decomposition + hand-derived adjoints the CUDA kernels implement (oracle/staged_np.py) against
torch.autograd on the op-for-op oracle.  CPU-only."""
import numpy as np
import pytest

from oracle import iwvi_oracle as O
from oracle import staged_np as ST
from oracle import synthetic as S


def _compare(spec, X, Y, eps, rtol=1e-9):
    e1, g1 = O.iw_elbo_and_grads(spec, X, Y, eps, reference_style=True)
    e2, g2 = ST.iw_elbo_and_grads(spec, X, Y, eps)
    np.testing.assert_allclose(e2, e1.item(), rtol=1e-12)
    assert set(g2) <= set(g1)
    for k in g1:
        a = g1[k].numpy()
        if k not in g2:
            assert np.all(a == 0), k
            continue
        scale = max(np.abs(a).max(), 1e-12)
        np.testing.assert_allclose(g2[k], a, rtol=rtol, atol=rtol * scale, err_msg=k)


@pytest.mark.parametrize("configuration,kern", [("L1", "RBF"), ("L1_G3", "RBF"), ("L1_G3_G2", "RBF"),
                                                ("L1_G3", "Matern52"), ("L2_G3", "Matern32"), ("G2", "Matern12"),
                                                ("G3_L1_G2", "RBF")])
def test_staged_matches_autograd(configuration, kern):
    N, D, M, K = 30, 3, 11, 4
    X, Y = S.make_data(N, D, seed=4)
    spec = S.make_spec(X, configuration, M, K, seed=4, perturb=0.3, inner_q_sqrt_scale=0.2, kern=kern)
    eps = S.make_noise(spec, (N, K), seed=5)
    # Matern12 is not differentiable at r=0: on the Kuu diagonal the expanded-form r2 is rounding noise
    # (~1e-16, not clamped), which autograd amplifies by 1/(2r); the staged adjoint zeroes that diagonal.
    _compare(spec, X, Y, eps, rtol=1e-6 if kern == 'Matern12' else 1e-9)


def test_staged_final_linear_mean_and_scalar_ls():
    N, D, M, K = 25, 2, 9, 3
    X, Y = S.make_data(N, D, seed=6)
    spec = S.make_spec(X, "L1_G2", M, K, seed=6, perturb=0.3, inner_q_sqrt_scale=0.5, final_mf='Linear')
    spec['layers'][-1]['lengthscales'] = np.array(1.3)
    eps = S.make_noise(spec, (N, K), seed=7)
    _compare(spec, X, Y, eps)
