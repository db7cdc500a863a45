"""Runs a few training steps of a BASELINE config (for ncu captures): python tools/profile_step.py c3 3"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dgps_with_iwvi_b200.build_models import build_model  # noqa: E402
from dgps_with_iwvi_b200.training import Trainer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'c3'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = bench.CONFIGS[name]
X, Y = bench.make_data(cfg['N'], cfg['D'], seed=0)
model = build_model(X, Y, cfg['configuration'], M=cfg['M'], num_IW_samples=cfg['K'], minibatch_size=cfg['B'],
                    likelihood_variance=cfg['lik_variance'], mode='IWAE', seed=0)
tr = Trainer(model, cfg['B'])
for i in range(steps):
    idx = torch.arange(i * cfg['B'], (i + 1) * cfg['B'], device=model.X.device) % cfg['N']
    loss = tr.step_device(model.X[idx], model.Y[idx])
torch.cuda.synchronize()
print('elbo', float(loss.item()))
