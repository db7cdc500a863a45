"""Shared test helpers: golden-fixture (de)serialisation and comparison utilities."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def canon(name):
    """Product parameter name -> oracle leaf name."""
    return (name.replace('.feature.feat.Z', '.Z').replace('.feature.Z', '.Z').replace('.kern.kernel.', '.kern.')
            .replace('.mean_function.', '.mf.'))


def spec_to_flat(spec):
    out = {'num_data': np.array(spec['num_data']), 'num_samples': np.array(spec['num_samples']),
           'lik_variance': np.asarray(spec['lik_variance']), 'n_layers': np.array(len(spec['layers']))}
    for i, ls in enumerate(spec['layers']):
        p = 'L%d.' % i
        out[p + 'type'] = np.array(ls['type'])
        if ls['type'] == 'lv':
            out[p + 'latent_dim'] = np.array(ls['latent_dim'])
            out[p + 'n'] = np.array(len(ls['Ws']))
            for j, (w, b) in enumerate(zip(ls['Ws'], ls['bs'])):
                out[p + 'W%d' % j] = w
                out[p + 'b%d' % j] = b
        else:
            for k in ('variance', 'lengthscales', 'Z', 'q_mu', 'q_sqrt', 'W', 'mf_A', 'mf_b'):
                if ls.get(k) is not None:
                    out[p + k] = np.asarray(ls[k])
            out[p + 'kern'] = np.array(ls['kern'])
            out[p + 'mf'] = np.array(ls['mf'])
            out[p + 'jitter'] = np.array(ls.get('jitter', 1e-6))
    return out


def spec_from_flat(f):
    layers = []
    for i in range(int(f['n_layers'])):
        p = 'L%d.' % i
        if str(f[p + 'type']) == 'lv':
            n = int(f[p + 'n'])
            layers.append(dict(type='lv', latent_dim=int(f[p + 'latent_dim']),
                               Ws=[f[p + 'W%d' % j] for j in range(n)], bs=[f[p + 'b%d' % j] for j in range(n)]))
        else:
            g = lambda k: f[p + k] if (p + k) in f else None
            layers.append(dict(type='gp', kern=str(f[p + 'kern']), variance=g('variance'), lengthscales=g('lengthscales'),
                               Z=g('Z'), q_mu=g('q_mu'), q_sqrt=g('q_sqrt'), W=g('W'), mf=str(f[p + 'mf']),
                               mf_A=g('mf_A'), mf_b=g('mf_b'), jitter=float(f[p + 'jitter'])))
    return dict(num_data=int(f['num_data']), num_samples=int(f['num_samples']), lik_variance=f['lik_variance'],
                layers=layers)


def save_golden(name, spec, X, Y, eps, elbo, grads, extra=None):
    d = spec_to_flat(spec)
    d['X'], d['Y'], d['elbo'] = X, Y, np.array(elbo)
    for i, e in enumerate(eps):
        if e is not None:
            d['eps%d' % i] = e
    for k, v in grads.items():
        d['grad:' + k] = np.asarray(v)
    for k, v in (extra or {}).items():
        d['extra:' + k] = np.asarray(v)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + '.npz'), **d)


def load_golden(name):
    f = dict(np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False))
    spec = spec_from_flat(f)
    eps = [f.get('eps%d' % i) for i in range(len(spec['layers']))]
    grads = {k[5:]: v for k, v in f.items() if k.startswith('grad:')}
    extra = {k[6:]: v for k, v in f.items() if k.startswith('extra:')}
    return spec, f['X'], f['Y'], eps, float(f['elbo']), grads, extra


def assert_grads_close(got, want, rtol, label=''):
    """got: product names -> arrays; want: oracle names -> arrays.  Relative to the largest entry of each tensor."""
    got = {canon(k): np.asarray(v) for k, v in got.items()}
    for k, w in want.items():
        w = np.asarray(w)
        assert k in got, '%s missing gradient %s' % (label, k)
        g = got[k].reshape(w.shape)
        scale = max(np.abs(w).max(), 1e-12)
        err = np.abs(g - w).max() / scale
        assert err < rtol, '%s grad %s: max err / max|ref| = %.3e' % (label, k, err)


def grad_report(got, want, rtol=1e-8, atol_rel=1e-12):
    """Per gradient tensor: the normwise error max|got - want| / max|want| AND the elementwise check north_star words
    as "gradients within rtol 1e-8": |got - want| <= rtol |want| + atol_rel max|want| entry by entry (the absolute floor
    is what float64 summation over thousands of points leaves on entries many orders below the tensor's largest).
    Returns {name: dict(normwise, n, n_fail, worst (largest |err| / allowed), worst_index)}."""
    got = {canon(k): np.asarray(v) for k, v in got.items()}
    out = {}
    for k, w in want.items():
        w = np.asarray(w, dtype=np.float64)
        g = got[k].reshape(w.shape)
        scale = max(np.abs(w).max(), 1e-300) if w.size else 1.0
        err = np.abs(g - w)
        allowed = rtol * np.abs(w) + atol_rel * scale
        ratio = err / allowed
        i = int(np.argmax(ratio)) if w.size else 0
        out[k] = dict(normwise=float(err.max() / scale) if w.size else 0.0, n=int(w.size),
                      n_fail=int((ratio > 1.0).sum()), worst=float(ratio.reshape(-1)[i]) if w.size else 0.0,
                      worst_index=i, max_abs=float(scale))
    return out


def assert_grads_close_elementwise(got, want, rtol=1e-8, atol_rel=1e-12, label=''):
    rep = grad_report(got, want, rtol, atol_rel)
    bad = {k: v for k, v in rep.items() if v['n_fail']}
    assert not bad, '%s: entries outside rtol %.0e + %.0e max|ref|: %s' % (
        label, rtol, atol_rel, {k: (v['n_fail'], v['n'], '%.2fx' % v['worst']) for k, v in bad.items()})
    return rep
