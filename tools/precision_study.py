"""What would a reduced-precision (tcgen05) fast path cost in accuracy?  CPU emulation on the oracle.

north_star allows an OPTIONAL fp32/TF32 fast path "reported separately with its stated tolerance".  tcgen05.mma has no f64
kind, so such a path would run the three contraction families of the per-point stage -- the gram -2 X Z^T, the triangular
solve A = Lm^-1 Kuf, the projections U_r = tril(q_sqrt_r)^T A -- on TF32 / BF16x3 / FP16x... operands with FP32
accumulation in TMEM.  This script restates the IW-ELBO forward (oracle formulas, reference temp_workaround.py:39-91,
models.py:112-150) with the OPERANDS of those products rounded to a given mantissa width and the products rounded to
fp32, differentiates it with a straight-through estimator, and reports ELBO / gradient errors against the float64 oracle:

    python tools/precision_study.py            # c2-shaped model, 64 rows

mantissa bits: 10 = TF32 / FP16, 7 = BF16, 23 = FP32 ("3xTF32" / BF16x3 splitting reaches about this), 52 = float64.
The outcome (profiles/precision_study_r02.json): the variance fvar = k(x,x) - |A|^2 + |U|^2 is a
difference of O(1) quantities that is itself 1e-3 .. 1e-10 at the reference's initialisation (inner q_sqrt = 1e-5 I), so
TF32 operands give NEGATIVE variances (NaN samples) and even FP32-grade products leave percent-level gradient errors; a
usable fast path needs error-free FP64 emulation (Ozaki slicing), which is a different project from a TF32 switch."""
import json
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import iwvi_oracle as O  # noqa: E402
from oracle import synthetic as S  # noqa: E402

DT = torch.float64


def rnd(x, bits):
    """Round to `bits` explicit mantissa bits (straight-through gradient)."""
    if bits >= 52:
        return x
    m, e = torch.frexp(x.detach())
    q = torch.ldexp(torch.round(m * 2.0 ** (bits + 1)) / 2.0 ** (bits + 1), e)
    return x + (q - x.detach())


def mm(a, b, bits):
    """Product of rounded operands, result rounded to fp32 (FP32 accumulation in TMEM)."""
    out = rnd(a, bits) @ rnd(b, bits)
    return out if bits >= 52 else rnd(out, 23)


def conditional(X, layer, eps, bits):
    """temp_workaround.py:39-91 (white, full_cov=False) with the three product families at `bits`."""
    kern = layer.kern.kernel if isinstance(layer.kern, O.Mok) else layer.kern
    Z, ls = layer.Z, kern.lengthscales
    Kmm = O.Kuu(Z, kern, layer.jitter)
    Lm = torch.linalg.cholesky(Kmm)                       # once per step: stays float64
    Xs, Zs = X / ls, Z / ls
    r2 = (Zs ** 2).sum(-1)[:, None] + (Xs ** 2).sum(-1)[None, :] - 2.0 * mm(Zs, Xs.t(), bits)
    if kern.kind != 'RBF':
        raise NotImplementedError
    Kmn = kern.variance * torch.exp(-0.5 * r2)
    # A = Lm^-1 Kmn as the blocked algorithm does it: products with (inverted) blocks of Lm -- emulated as one product
    # with the explicit inverse at the reduced precision
    Linv = torch.linalg.solve_triangular(Lm, torch.eye(Lm.shape[0], dtype=DT), upper=False)
    A = mm(Linv, Kmn, bits)
    fvar0 = kern.variance - (A ** 2).sum(0)
    gmean = mm(A.t(), layer.q_mu, bits)
    gvar = []
    for r in range(layer.q_mu.shape[1]):
        U = mm(torch.tril(layer.q_sqrt[r]).t(), A, bits)
        gvar.append(fvar0 + (U ** 2).sum(0))
    gvar = torch.stack(gvar, 1)
    smp = None if eps is None else gmean + eps * gvar ** 0.5
    if isinstance(layer.kern, O.Mok):
        W = layer.kern.W
        smp = None if smp is None else smp @ W.t()
        gmean, gvar = gmean @ W.t(), gvar @ (W ** 2).t()
    mf = layer.mean_function(X)
    return (None if smp is None else smp + mf), gmean + mf, gvar


def iw_elbo(model, X, Y, eps, bits):
    N, K = X.shape[0], model.num_samples
    F = X[:, None, :].repeat(1, K, 1).reshape(N * K, -1)
    XY = torch.cat([X, Y], 1)[:, None, :].repeat(1, K, 1).reshape(N * K, -1)
    local, kl = 0.0, 0.0
    for li, (layer, e) in enumerate(zip(model.layers, eps)):
        if isinstance(layer, O.LatentVariableLayer):
            s, _, _, k = layer.propagate(F, XY, True, eps=e.reshape(N * K, -1))
            F, local = s, local + k.sum(-1)
        else:
            last = li == len(model.layers) - 1
            s, m, v = conditional(F, layer, None if last else e.reshape(N * K, -1), bits)
            kl = kl + O.gauss_kl(layer.q_mu, layer.q_sqrt)
            F = s
    ve = O.gaussian_variational_expectations(m, v, Y[:, None, :].repeat(1, K, 1).reshape(N * K, -1), model.lik_variance)
    L = (ve.sum(-1) - local).reshape(N, K)
    return (torch.logsumexp(L, 1) - math.log(K)).sum() * (model.num_data / N) - kl, v


def run(qs, rows=64):
    c = S.CONFIGS['c2']
    X, Y = S.make_data(2000, c['D'], seed=0)
    spec = S.make_spec(X, c['configuration'], c['M'], c['K'], lik_variance=c['lik_variance'], seed=0, perturb=0.1,
                       inner_q_sqrt_scale=qs)
    eps = [None if e is None else torch.as_tensor(e) for e in S.make_noise(spec, (rows, c['K']), seed=1)]
    Xb, Yb = torch.as_tensor(X[:rows]), torch.as_tensor(Y[:rows])
    out = {}
    ref = None
    for name, bits in (('float64', 52), ('fp32-grade products (3xTF32 / BF16x3)', 23), ('TF32 / FP16 operands', 10),
                       ('BF16 operands', 7)):
        model, leaves = O.build_from_spec(spec, requires_grad=True)
        elbo, v = iw_elbo(model, Xb, Yb, eps, bits)
        names = list(leaves)
        ok = bool(torch.isfinite(elbo))
        grads = torch.autograd.grad(elbo, [leaves[n] for n in names], allow_unused=True) if ok else None
        g = None if grads is None else {n: (torch.zeros_like(leaves[n]) if gr is None else gr) for n, gr in zip(names, grads)}
        if bits == 52:
            ref = (elbo.item(), g)
            out[name] = {'elbo': elbo.item()}
            continue
        rec = {'elbo': elbo.item() if ok else None, 'elbo_rel_err': abs(elbo.item() - ref[0]) / abs(ref[0]) if ok else None,
               'negative_final_variances': int((v < 0).sum().item())}
        if g is not None and all(torch.isfinite(x).all() for x in g.values()):
            rec['worst_grad_normwise_err'] = max(((g[n] - ref[1][n]).abs().max() / ref[1][n].abs().max().clamp_min(1e-300)).item()
                                                 for n in names)
        else:
            rec['worst_grad_normwise_err'] = None
        out[name] = rec
    return out


if __name__ == '__main__':
    res = {'inner q_sqrt = 1e-5 I (reference initialisation, build_models.py:275-278)': run(1e-5),
           'inner q_sqrt = 0.3 I (well into training)': run(0.3)}
    print(json.dumps(res, indent=1))
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles',
                           'precision_study_r02.json'), 'w') as f:
        json.dump(res, f, indent=1)
