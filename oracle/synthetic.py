"""
TEST INFRASTRUCTURE ONLY -- synthetic model specs and data of the BASELINE.json shapes.

Follows the shape logic of experiments/build_models.py:176-241 (the reference's model factory) without
its k-means / SVD dependencies on real data:
  * 'L<d>'  -> LatentVariableLayer(d, XY_dim=DX+1), D_in += d                    (:233-236)
  * 'G<d>'  -> GPLayer(SharedMixedMok(RBF(D_in, ls=sqrt(D_in), var=1, ARD), W[D_out,d]),
                       InducingPoints(ZZ[M,D_in]), d, Linear(A=[I;0]))            (:203-231)
  * final   -> GPLayer(RBF(D_in, ls=sqrt(D_in), var=1, ARD), InducingPoints(ZZ), DY=1)   (:238-241)
  * inner q_sqrt *= 1e-5                                                          (:275-278)
Z is a random subset of X (stand-in for kmeans2), W the top right-singular vectors of X (:186,216-217).
A seeded perturbation (perturb > 0) makes q_mu/q_sqrt/lengthscales/variance non-degenerate so that
every gradient is exercised.
"""
import numpy as np


def make_data(N, D, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D))
    w = rng.standard_normal((D, 1)) / np.sqrt(D)
    Y = np.sin(X @ w) + 0.1 * rng.standard_normal((N, 1))
    return X, Y


def demo_data(seed=0):
    """Shape-equivalent of experiments/demo.py:21-45 (N=200, Dx=1, bimodal); seeded with default_rng."""
    rng = np.random.default_rng(seed)
    x1 = rng.uniform(-3, -0.5, size=(100, 1))
    x3 = rng.uniform(1, 3, size=(100, 1))
    X = np.concatenate([x1, x3], 0)
    ind = rng.random(X.shape) < 0.6
    f1 = np.exp(-(X - 1) ** 2) + np.exp(-(X + 1) ** 2) + 0.1 * np.exp(rng.standard_normal(X.shape))
    f2 = np.exp(-(X - 1) ** 2) + rng.uniform(-0.1, 0.1, size=X.shape)
    Y = np.where(ind, f1, f2)
    return X, Y


def make_spec(X, configuration, M, K, lik_variance=0.01, seed=0, perturb=0.1, inner_q_sqrt_scale=1e-5,
              kern='RBF', final_mf='Zero', jitter=1e-6, encoder_dims=(20, 20)):
    """Returns a spec dict (numpy float64 arrays) accepted by oracle.iwvi_oracle.build_from_spec and by
    tests/helpers.model_from_spec."""
    rng = np.random.default_rng(seed + 1000)
    N, D = X.shape
    DX, DY = D, 1
    if N > M:
        Z = X[rng.choice(N, M, replace=False)].copy()
    else:
        Z = np.concatenate([X.copy(), rng.standard_normal((M - N, D))], 0)
    P = np.linalg.svd(X, full_matrices=False)[2]
    layers = []
    D_in, D_out = D, D

    def tril_rand(R, scale):
        q = np.tile(np.eye(M)[None], [R, 1, 1]) * scale
        if perturb > 0:
            q = q + scale * perturb * np.tril(rng.standard_normal((R, M, M))) / np.sqrt(M)
        return q

    tokens = [t for t in configuration.split('_') if t]
    for tok in tokens:
        c, d = tok[0], int(tok[1:])
        if c == 'G':
            A = np.zeros((D_in, D_out))
            D_min = min(D_in, D_out)
            A[:D_min, :D_min] = np.eye(D_min)
            PP = np.zeros((D_out, d))
            PP[:, :min(d, DX)] = P[:, :min(d, DX)]
            ZZ = rng.standard_normal((M, D_in))
            ZZ[:, :min(D_in, DX)] = Z[:, :min(D_in, DX)]
            ls = np.full((D_in,), float(D_in) ** 0.5)
            if perturb > 0:
                ls = ls * np.exp(perturb * rng.standard_normal(D_in))
            layers.append(dict(type='gp', kern=kern, variance=np.array(1.0), lengthscales=ls, Z=ZZ,
                               q_mu=perturb * rng.standard_normal((M, d)),
                               q_sqrt=tril_rand(d, inner_q_sqrt_scale),
                               W=PP, mf='Linear', mf_A=A, mf_b=np.zeros(D_out), jitter=jitter))
            D_in = D_out
        elif c == 'L':
            D_in += d
            dims = [DX + 1, *encoder_dims, 2 * d]
            Ws = [rng.standard_normal((a, b)) * (2.0 / (a + b)) ** 0.5 for a, b in zip(dims[:-1], dims[1:])]
            bs = [perturb * rng.standard_normal(b) for b in dims[1:]]
            layers.append(dict(type='lv', latent_dim=d, Ws=Ws, bs=bs))
        else:
            raise ValueError(tok)
    ZZ = rng.standard_normal((M, D_in))
    ZZ[:, :min(D_in, DX)] = Z[:, :min(D_in, DX)]
    ls = np.full((D_in,), float(D_in) ** 0.5)
    var = np.array(1.0)
    if perturb > 0:
        ls = ls * np.exp(perturb * rng.standard_normal(D_in))
        var = np.array(1.0 + perturb)
    final = dict(type='gp', kern=kern, variance=var, lengthscales=ls, Z=ZZ,
                 q_mu=perturb * rng.standard_normal((M, DY)), q_sqrt=tril_rand(DY, 1.0),
                 W=None, mf=final_mf, mf_A=None, mf_b=None, jitter=jitter)
    if final_mf == 'Linear':
        final['mf_A'] = rng.standard_normal((D_in, DY)) / np.sqrt(D_in)
        final['mf_b'] = perturb * rng.standard_normal(DY)
    layers.append(final)
    return dict(num_data=N, num_samples=K, lik_variance=np.array(lik_variance), layers=layers)


def make_noise(spec, lead_shape, seed=0, final_noise=False):
    """One standard-normal tensor per layer with the reference's index order [*lead, C]
    (layers.py:86, temp_workaround.py:89).  The final GP layer gets None unless final_noise (the IW
    objective never draws it)."""
    rng = np.random.default_rng(seed + 2000)
    eps = []
    n_layers = len(spec['layers'])
    for i, ls in enumerate(spec['layers']):
        if ls['type'] == 'lv':
            eps.append(rng.standard_normal(tuple(lead_shape) + (ls['latent_dim'],)))
        elif i == n_layers - 1 and not final_noise:
            eps.append(None)
        else:
            eps.append(rng.standard_normal(tuple(lead_shape) + (ls['q_mu'].shape[1],)))
    return eps


# BASELINE.json configs -> (configuration, N, D, M, K, B, lik_variance)
CONFIGS = {
    'c1': dict(configuration='L1', N=200, D=1, M=50, K=20, B=200, lik_variance=0.1),
    'c2': dict(configuration='L1_G5', N=10000, D=8, M=100, K=20, B=512, lik_variance=0.01),
    'c3': dict(configuration='L1_G5_G5', N=100000, D=16, M=256, K=50, B=512, lik_variance=0.01),
    'c4': dict(configuration='L1_G5', N=16384, D=8, M=512, K=256, B=4096, lik_variance=0.01),
    'c5': dict(configuration='L1_G5_G5', N=1000000, D=8, M=256, K=50, B=512, lik_variance=0.01),
}
