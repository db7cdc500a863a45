"""c3-sized check that the pipelined, graph-captured Trainer (segment-wise Adam, next-step Cholesky on high-priority
streams, two point chains) trains bit-identically to the plain eager Trainer: any cross-stream hazard shows up as a
difference in the parameters after a few steps."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from dgps_with_iwvi_b200.build_models import build_model
from dgps_with_iwvi_b200.engine import FlatParams
from dgps_with_iwvi_b200.training import Trainer
name = sys.argv[1] if len(sys.argv) > 1 else 'c3'
cfg = bench.CONFIGS[name]
X, Y = bench.make_data(cfg['N'], cfg['D'], seed=0)
B = cfg['B']
def run(pipeline, graph, steps=7):
    model = build_model(X, Y, cfg['configuration'], M=cfg['M'], num_IW_samples=cfg['K'], minibatch_size=B,
                        likelihood_variance=cfg['lik_variance'], mode='IWAE', seed=0)
    tr = Trainer(model, B, lr=5e-3, seed=3, use_graph=graph, pipeline=pipeline)
    losses = []
    for i in range(steps):
        idx = (torch.arange(i * B, (i + 1) * B, device=model.X.device) * 7919) % cfg['N']
        losses.append(tr.step_device(model.X[idx], model.Y[idx]).clone())
    torch.cuda.synchronize()
    tr.engine.check_info()
    return FlatParams.of(model).x.clone(), torch.cat(losses).cpu().numpy()
x0, l0 = run(False, False)
for rep in range(4):
    x1, l1 = run(True, True)
    nd = int((x1 != x0).sum())
    print(name, 'rep', rep, 'params differing:', nd, 'of', x0.numel(), 'max rel diff %.3e' % float(((x1 - x0).abs().max() / x0.abs().max()).item()),
          'losses equal:', bool((l0 == l1).all()), flush=True)
