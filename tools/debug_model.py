"""Debug aid: model-level gradient errors for one live-oracle configuration (prints every tensor's error)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import helpers as H
from oracle import iwvi_oracle as O
from oracle import synthetic as S
from dgps_with_iwvi_b200.build_models import model_from_spec
conf, N, D, M, K, kern = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
X, Y = S.make_data(N, D, seed=11)
spec = S.make_spec(X, conf, M, K, seed=11, perturb=0.3, inner_q_sqrt_scale=0.3, kern=kern)
eps = S.make_noise(spec, (N, K), seed=12)
e_ref, g_ref = O.iw_elbo_and_grads(spec, X, Y, eps, reference_style=True)
m = model_from_spec(spec, X, Y)
for rep in range(3):
    e, g = m.compute_log_likelihood_and_grads(X, Y, eps)
    got = {H.canon(k): np.asarray(v) for k, v in g.items()}
    print('rep', rep, 'elbo err', abs(e - e_ref.item()) / abs(e_ref.item()))
    for k, w in g_ref.items():
        w = w.numpy(); gg = got[k].reshape(w.shape)
        err = np.abs(gg - w).max() / max(np.abs(w).max(), 1e-12)
        if err > 1e-9 or rep == 0:
            print('  %-32s %.3e' % (k, err))
