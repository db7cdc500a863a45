"""Per-launch timeline of one eager training step (all streams): wraps every capi entry point with CUDA events recorded
on the stream it is launched on, runs a few steps of a BASELINE config eagerly and prints, for the last step, each call's
start / end offset from the beginning of the step and the stream it ran on.  Eager launches keep the GPU fed at c3's
sizes (3.3 ms of device work per step against ~0.6 ms of host launch time), so the offsets show which chains are exposed.

    python tools/step_timeline.py c3 [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dgps_with_iwvi_b200 import capi  # noqa: E402
from dgps_with_iwvi_b200.build_models import build_model  # noqa: E402
from dgps_with_iwvi_b200.training import Trainer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'c3'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cfg = bench.CONFIGS[name]
X, Y = bench.make_data(cfg['N'], cfg['D'], seed=0)
model = build_model(X, Y, cfg['configuration'], M=cfg['M'], num_IW_samples=cfg['K'], minibatch_size=cfg['B'],
                    likelihood_variance=cfg['lik_variance'], mode='IWAE', seed=0)
tr = Trainer(model, cfg['B'], use_graph=False)

LOG = []
WRAP = ['gp_prologue_fwd', 'gp_rows_fwd', 'gp_rows_fwd_range', 'gp_rows_bwd', 'gp_prologue_bwd', 'lv_fwd', 'lv_bwd',
        'iwelbo_fwd', 'iwelbo_bwd', 'normal_fill', 'normal_fill_counter', 'adam_step_counter', 'positive_fwd']
FLAGS = {16: 'EPI', 32: 'TILE', 64: 'REDUCE', 128: 'FINAL', 256: 'A', 512: 'B', 1024: 'SKIPKL', 2048: 'ONLYKL'}


def wrap(fn_name):
    fn = getattr(capi, fn_name)

    def w(*a, **k):
        st = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        out = fn(*a, **k)
        e1.record(st)
        tag = fn_name
        if a and hasattr(a[0], 'flags') and fn_name in ('gp_rows_bwd', 'gp_prologue_bwd'):
            tag += '[' + '|'.join(v for b, v in FLAGS.items() if a[0].flags & b) + ']'
        if a and hasattr(a[0], 'R'):
            tag += ' R=%d' % a[0].R
        LOG.append((tag, st.cuda_stream, e0, e1))
        return out
    setattr(capi, fn_name, w)


for n in WRAP:
    wrap(n)
for i in range(steps):
    idx = torch.arange(i * cfg['B'], (i + 1) * cfg['B'], device=model.X.device) % cfg['N']
    if i == steps - 1:
        torch.cuda.synchronize()
        LOG.clear()
        t0 = torch.cuda.Event(enable_timing=True)
        t0.record()
    loss = tr.step_device(model.X[idx], model.Y[idx])
t1 = torch.cuda.Event(enable_timing=True)
t1.record()
torch.cuda.synchronize()
streams = {}
print('step: %.3f ms (eager)' % t0.elapsed_time(t1))
print('%-44s %6s %9s %9s %9s' % ('call', 'stream', 'start us', 'end us', 'dur us'))
for tag, s, e0, e1 in LOG:
    sid = streams.setdefault(s, len(streams))
    a, b = t0.elapsed_time(e0) * 1e3, t0.elapsed_time(e1) * 1e3
    print('%-44s %6d %9.1f %9.1f %9.1f' % (tag, sid, a, b, b - a))
