"""Times the parameter-contraction launch of one inner GP layer alone: float64 DMMA kernel vs the optional tcgen05 3xTF32
variant (IWVI_FLAG_FAST_REDUCE), on the buffers a training step left behind.  python tools/time_fast_reduce.py c3"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dgps_with_iwvi_b200 import _lib as LIB  # noqa: E402
from dgps_with_iwvi_b200 import capi  # noqa: E402
from dgps_with_iwvi_b200.build_models import build_model  # noqa: E402
from dgps_with_iwvi_b200.training import Trainer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'c3'
cfg = bench.CONFIGS[name]
X, Y = bench.make_data(min(cfg['N'], 100000), cfg['D'], seed=0)
model = build_model(X, Y, cfg['configuration'], M=cfg['M'], num_IW_samples=cfg['K'], minibatch_size=cfg['B'],
                    likelihood_variance=cfg['lik_variance'], mode='IWAE', seed=0)
tr = Trainer(model, cfg['B'], use_graph=False)
tr.step_device(model.X[:cfg['B']], model.Y[:cfg['B']])
eng = tr.engine
gps = [r for r in eng.recs if r['type'] == 'gp']
r = max(gps, key=lambda q: q['R'] * q['M'] * q['M'])
flat, layer, base, feat = eng.flat, r['layer'], r['base'], r['feat']
W = flat.cview(layer.kern.W) if r['mix'] else None
lin = r['mf'] == 'Linear'
mfA = flat.cview(layer.mean_function.A) if lin else None
mfb = flat.cview(layer.mean_function.b) if lin else None
z = lambda t: torch.zeros_like(t)
outs = [z(flat.gview(p)) for p in (feat.Z, base.lengthscales, base.variance, layer.q_mu, layer.q_sqrt)]
if outs[1].numel() != r['D']:
    outs[1] = torch.zeros(r['D'], dtype=torch.float64, device=eng.dev)
d_s = torch.randn(eng.T, r['P'], dtype=torch.float64, device=eng.dev)


def run(flags):
    d = capi.with_flags(r['d'], r['d'].flags | flags)
    capi.gp_rows_bwd(d, r['Lm'], r['aux'], r['save'], r['Fin'], W, mfA, mfb, r['eps'], d_s if r['sampled'] else None,
                     None if r['sampled'] else d_s, None if r['sampled'] else d_s, r['dX'], outs[0], outs[1], outs[2], outs[3],
                     outs[4], r['dLm'], None, None, None, r['bwd_ws'])


run(0)
res = {}
for tag, fl in (('exact', LIB.FLAG_ONLY_REDUCE), ('fast', LIB.FLAG_ONLY_REDUCE | LIB.FLAG_FAST_REDUCE)):
    run(fl | LIB.FLAG_ONLY_FINAL)
    res[tag] = (outs[3].clone(), outs[4].clone())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run(fl)
    e1.record()
    torch.cuda.synchronize()
    print('%s: %.3f ms per reduce launch (M=%d R=%d T=%d)' % (tag, e0.elapsed_time(e1) / 10, r['M'], r['R'], eng.T))
for i, nm in enumerate(('dq_mu', 'dq_sqrt')):
    a, b = res['exact'][i], res['fast'][i]
    print('%s: max|fast - exact| / max|exact| = %.2e' % (nm, (a - b).abs().max().item() / a.abs().max().item()))
