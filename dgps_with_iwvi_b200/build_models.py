"""Model construction from the reference's configuration strings and from the plain-dict spec the tests share with
the oracle.  `build_model` follows the shape logic of the reference factory (experiments/build_models.py:176-241,
264-287: 'L<d>' -> LatentVariableLayer, 'G<d>' -> Mok GPLayer with Linear mean function, final plain-RBF GPLayer, inner
q_sqrt scaled by 1e-5, frozen linear parts and inner kernel variances)."""
import numpy as np

from .features import InducingPoints, MixedKernelSharedMof
from .kernels import RBF, Matern12, Matern32, Matern52, SharedMixedMok
from .layers import Encoder, GPLayer, LatentVariableLayer
from .likelihoods import Gaussian
from .mean_functions import Identity, Linear, Zero
from .models import DGP_IWVI, DGP_VI

KERNELS = {'RBF': RBF, 'Matern12': Matern12, 'Matern32': Matern32, 'Matern52': Matern52}


def model_from_spec(spec, X, Y, mode='iw', minibatch_size=None):
    """spec: the dict format of oracle/synthetic.py (constrained numpy values).  Returns DGP_IWVI / DGP_VI."""
    layers = []
    for ls in spec['layers']:
        if ls['type'] == 'lv':
            dims = [w.shape[0] for w in ls['Ws']] + [ls['Ws'][-1].shape[1]]
            enc = Encoder(ls['latent_dim'], dims[0], dims[1:-1], activation_func=ls.get('activation'), seed=0)
            for p, w in zip(enc.Ws, ls['Ws']):
                p.assign(w)
            for p, b in zip(enc.bs, ls['bs']):
                p.assign(b)
            layers.append(LatentVariableLayer(ls['latent_dim'], encoder=enc))
        else:
            M, D = ls['Z'].shape
            R = ls['q_mu'].shape[1]
            lsc = np.asarray(ls['lengthscales'], dtype=np.float64)
            adims = ls.get('active_dims')
            kern = KERNELS[ls['kern']](D if adims is None else len(adims), variance=float(ls['variance']), lengthscales=lsc,
                                       ARD=lsc.ndim > 0 and lsc.size > 1, active_dims=adims)
            feat = InducingPoints(ls['Z'])
            if ls.get('W') is not None:
                kern = SharedMixedMok(kern, ls['W'])
                feat = MixedKernelSharedMof(feat)
            mf = {'Zero': lambda: Zero(), 'Identity': lambda: Identity(),
                  'Linear': lambda: Linear(ls['mf_A'], ls['mf_b'])}[ls['mf']]()
            layer = GPLayer(kern, feat, R, mean_function=mf)
            layer.q_mu = ls['q_mu']
            layer.q_sqrt = ls['q_sqrt']
            layer.jitter = ls.get('jitter', 1e-6)
            layers.append(layer)
    lik = Gaussian(variance=float(spec['lik_variance']))
    cls = DGP_IWVI if mode == 'iw' else DGP_VI
    model = cls(X, Y, layers, lik, num_samples=spec['num_samples'], minibatch_size=minibatch_size)
    model.num_data = spec['num_data']
    return model


def build_model(X, Y, configuration='L1_G5_G5', M=128, num_IW_samples=5, minibatch_size=512, likelihood_variance=1e-2,
                mode='IWAE', fix_linear=True, Z=None, seed=0, init='subset'):
    """experiments/build_models.py:146-287 for mode in {'IWAE', 'VI'}.  Z: [M, DX] initial inducing inputs (the
    reference runs scipy kmeans2 on X, :179-183).  Z=None: init='kmeans' does the same (scipy.cluster.vq.kmeans2 with
    minit='points', host side, seconds at N=100k), init='subset' takes a random subset of rows (benchmarks)."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    N, DX = X.shape
    DY = Y.shape[1]
    rng = np.random.default_rng(seed)
    if Z is None:
        if N > M and init == 'kmeans':
            from scipy.cluster.vq import kmeans2
            Z = kmeans2(X, M, minit='points', seed=seed)[0]
        elif N > M:
            Z = X[rng.choice(N, M, replace=False)]
        else:
            Z = np.concatenate([X, rng.standard_normal((M - N, DX))], 0)
    P = np.linalg.svd(X, full_matrices=False)[2]
    layers = []
    D_in = D_out = DX
    for tok in [t for t in configuration.split('_') if t]:
        c, d = tok[0], int(tok[1:])
        if c == 'G':
            A = np.zeros((D_in, D_out))
            D_min = min(D_in, D_out)
            A[:D_min, :D_min] = np.eye(D_min)
            mf = Linear(A, np.zeros(D_out))
            mf.b.set_trainable(False)
            kern = RBF(D_in, lengthscales=float(D_in) ** 0.5, variance=1.0, ARD=True)
            kern.variance.set_trainable(False)
            PP = np.zeros((D_out, d))
            PP[:, :min(d, DX)] = P[:, :min(d, DX)]
            ZZ = rng.standard_normal((M, D_in))
            ZZ[:, :min(D_in, DX)] = Z[:, :min(D_in, DX)]
            layer = GPLayer(SharedMixedMok(kern, PP), MixedKernelSharedMof(InducingPoints(ZZ)), d, mean_function=mf)
            if fix_linear:
                layer.kern.W.set_trainable(False)
                mf.set_trainable(False)
            layer.q_sqrt = layer.q_sqrt.read_value() * 1e-5
            layers.append(layer)
            D_in = D_out
        elif c == 'L':
            D_in += d
            layers.append(LatentVariableLayer(d, encoder=Encoder(d, DX + DY, [20, 20], seed=seed + len(layers))))
        else:
            raise ValueError(tok)
    kern = RBF(D_in, lengthscales=float(D_in) ** 0.5, variance=1.0, ARD=True)
    ZZ = rng.standard_normal((M, D_in))
    ZZ[:, :min(D_in, DX)] = Z[:, :min(D_in, DX)]
    layers.append(GPLayer(kern, InducingPoints(ZZ), DY))
    lik = Gaussian(variance=likelihood_variance)
    if mode == 'IWAE':
        return DGP_IWVI(X, Y, layers, lik, minibatch_size=minibatch_size, num_samples=num_IW_samples)
    if mode == 'VI':
        return DGP_VI(X, Y, layers, lik, minibatch_size=minibatch_size, num_samples=1)
    raise NotImplementedError('mode %s is outside the IW-ELBO hot path (SURVEY.md section 8)' % mode)


def spec_from_model(model):
    """Inverse of model_from_spec: constrained numpy values in the dict format the oracle consumes."""
    from .layers import GPLayer
    layers = []
    for layer in model.layers:
        if isinstance(layer, GPLayer):
            mix = hasattr(layer.kern, 'W')
            base = layer.kern.kernel if mix else layer.kern
            feat = layer.feature.feat if hasattr(layer.feature, 'feat') else layer.feature
            mf = layer.mean_function
            layers.append(dict(type='gp', kern=base.kind, variance=base.variance.read_value(),
                               lengthscales=base.lengthscales.read_value(), Z=feat.Z.read_value(),
                               q_mu=layer.q_mu.read_value(), q_sqrt=layer.q_sqrt.read_value(),
                               W=layer.kern.W.read_value() if mix else None, mf=mf.kind,
                               mf_A=mf.A.read_value() if mf.kind == 'Linear' else None,
                               mf_b=mf.b.read_value() if mf.kind == 'Linear' else None, jitter=layer.jitter,
                               active_dims=getattr(base, 'active_dims', None)))
        else:
            enc = layer.encoder
            layers.append(dict(type='lv', latent_dim=layer.latent_dim, Ws=[w.read_value() for w in enc.Ws],
                               bs=[b.read_value() for b in enc.bs], activation=enc.activation_func))
    return dict(num_data=model.num_data, num_samples=model.num_samples,
                lik_variance=model.likelihood.variance.read_value(), layers=layers)
