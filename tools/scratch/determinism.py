import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from dgps_with_iwvi_b200.build_models import build_model
from dgps_with_iwvi_b200.engine import FlatParams, Engine
cfg = bench.CONFIGS['c3']
X, Y = bench.make_data(cfg['N'], cfg['D'], seed=0)
model = build_model(X, Y, cfg['configuration'], M=cfg['M'], num_IW_samples=cfg['K'], minibatch_size=cfg['B'],
                    likelihood_variance=cfg['lik_variance'], mode='IWAE', seed=0)
flat = FlatParams.of(model)
for split in (True, False):
    eng = Engine(model, cfg['B'], cfg['K'], 'iw', split_waves=split)
    ref = None
    for it in range(6):
        eng.elbo_and_grads(X[:cfg['B']], Y[:cfg['B']], None, seed=3, step=1, row0=0)
        torch.cuda.synchronize()
        g = flat.g.clone()
        if ref is None:
            ref = g
        else:
            bad = (g != ref)
            if bad.any():
                names = []
                for p in flat.params:
                    n, o, sz, shape, _ = flat.entries[id(p)]
                    nb = int(bad[o:o + sz].sum())
                    if nb:
                        names.append('%s:%d (max diff %.2e)' % (n, nb, float((g[o:o+sz] - ref[o:o+sz]).abs().max())))
                print('split', split, 'run', it, 'DIFFERS', names, flush=True)
            else:
                print('split', split, 'run', it, 'identical', flush=True)
