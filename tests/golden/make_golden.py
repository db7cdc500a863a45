"""Generates the golden vectors under tests/golden/ FROM THE ORACLE (oracle/iwvi_oracle.py, reference_style=True:
the final layer's full KxK covariance then its diagonal, exactly as reference models.py:123,133).  The reference
itself ships no vectors and cannot run here (no TensorFlow/GPflow), see the oracle's header.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from oracle import iwvi_oracle as O   # noqa: E402
from oracle import synthetic as S     # noqa: E402
import helpers as H                    # noqa: E402

CASES = {
    # name: (configuration, N(=B), D, M, K, kern, seed, final_mf)
    'iw_c1_demo_shape': ('L1', 60, 1, 50, 20, 'RBF', 1, 'Zero'),
    'iw_L1_G5': ('L1_G5', 48, 8, 100, 20, 'RBF', 2, 'Zero'),
    'iw_L1_G5_G5': ('L1_G5_G5', 24, 16, 70, 10, 'RBF', 3, 'Zero'),
    'iw_matern52_linear': ('L1_G3', 40, 3, 33, 7, 'Matern52', 4, 'Linear'),
}


def main():
    for name, (conf, N, D, M, K, kern, seed, fmf) in CASES.items():
        X, Y = (S.demo_data(seed)[0][:N], S.demo_data(seed)[1][:N]) if name.startswith('iw_c1') else S.make_data(N, D, seed)
        spec = S.make_spec(X, conf, M, K, seed=seed, perturb=0.2, inner_q_sqrt_scale=0.1, kern=kern, final_mf=fmf,
                           lik_variance=0.1 if name.startswith('iw_c1') else 0.01)
        spec['num_data'] = 10 * N   # exercise scale != 1
        eps = S.make_noise(spec, (N, K), seed=seed)
        elbo, grads = O.iw_elbo_and_grads(spec, X, Y, eps, reference_style=True)
        _, vgrads = O.vi_elbo_and_grads(spec, X, Y, [None if e is None else
                                                     np.transpose(e, (1, 0, 2)).reshape(K * N, -1) for e in eps])
        velbo = O.vi_elbo_and_grads(spec, X, Y, [None if e is None else
                                                 np.transpose(e, (1, 0, 2)).reshape(K * N, -1) for e in eps])[0]
        H.save_golden(name, spec, X, Y, eps, elbo.item(), {k: v.numpy() for k, v in grads.items()},
                      extra={'vi_elbo': velbo.item()})
        print(name, elbo.item(), velbo.item())


if __name__ == '__main__':
    main()
