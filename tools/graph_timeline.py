"""Timeline of ONE replayed training-step graph (all streams): every capi launch is bracketed by two one-thread stamp
kernels (iwvi_debug_stamp: %globaltimer) on the stream it is issued on, so the stamps are captured into the step graph
with the same dependencies as the launches.  Prints start / end offsets of every call of the last replay.

    python tools/graph_timeline.py c3 [replays]"""
import ctypes as C
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
replays = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = bench.CONFIGS[name]
X, Y = bench.make_data(cfg['N'], cfg['D'], seed=0)
model = build_model(X, Y, cfg['configuration'], M=cfg['M'], num_IW_samples=cfg['K'], minibatch_size=cfg['B'],
                    likelihood_variance=cfg['lik_variance'], mode='IWAE', seed=0)
tr = Trainer(model, cfg['B'])
slots = torch.zeros(4096, dtype=torch.int64, device=model.X.device)
NAMES = []
WRAP = ['gp_prologue_fwd', 'gp_rows_fwd', 'gp_rows_fwd_range', 'gp_rows_bwd', 'gp_rows_bwd_range', 'gp_prologue_bwd',
        'lv_fwd', 'lv_bwd', 'iwelbo_fwd', 'iwelbo_bwd', 'normal_fill', 'normal_fill_counter', 'adam_step_counter',
        'adam_step_counter_part', 'positive_fwd', 'batch_gather']
FLAGS = {16: 'EPI', 32: 'TILE', 64: 'REDUCE', 128: 'FINAL', 256: 'A', 512: 'B', 1024: 'SKIPKL', 2048: 'ONLYKL',
         16384: 'HYP', 32768: 'Q'}
lib = LIB.load()
ON = [False]


def stamp(tag):
    st = torch.cuda.current_stream()
    NAMES.append((tag, st.cuda_stream))
    LIB.check(lib.iwvi_debug_stamp(slots.data_ptr(), len(NAMES) - 1, st.cuda_stream), 'stamp')


def wrap(fn_name):
    fn = getattr(capi, fn_name)

    def w(*a, **k):
        if not ON[0]:
            return fn(*a, **k)
        tag = fn_name
        if a and hasattr(a[0], 'flags'):
            tag += '[' + '|'.join(v for b, v in FLAGS.items() if a[0].flags & b) + ']'
        if a and hasattr(a[0], 'R'):
            tag += ' R=%d' % a[0].R
        if fn_name == 'gp_rows_bwd_range' or fn_name == 'gp_rows_fwd_range':
            tag += ' [%d,%d)' % (a[-2], a[-1])
        stamp(tag + '\ts')
        out = fn(*a, **k)
        stamp(tag + '\te')
        return out
    setattr(capi, fn_name, w)


for n in WRAP:
    wrap(n)
idx = lambda i: torch.arange(i * cfg['B'], (i + 1) * cfg['B'], device=model.X.device) % cfg['N']
for i in range(2):
    tr.step_indices(idx(i))
ON[0] = True          # the capture happens inside the next call
NAMES.clear()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(2, 2 + replays):
    torch.cuda.synchronize()
    e0.record()
    loss = tr.step_indices(idx(i))
    e1.record()
torch.cuda.synchronize()
t = slots.cpu().numpy()
recs = {}
for i, (tag, st) in enumerate(NAMES):
    nm, kind = tag.split('\t')
    recs.setdefault((nm, st, i // 2 if False else None), None)
streams = {}
pairs = []
open_ = {}
for i, (tag, st) in enumerate(NAMES):
    nm, kind = tag.split('\t')
    if kind == 's':
        open_[(nm, st)] = i
    else:
        pairs.append((nm, st, t[open_.pop((nm, st))], t[i]))
t0 = min(p[2] for p in pairs)
print('%s: last replay %.3f ms (graph with %d stamps; stamps add launch slots, so the total is a little above the bench)'
      % (name, e0.elapsed_time(e1), len(NAMES)))
print('%-52s %6s %9s %9s %9s' % ('call', 'stream', 'start us', 'end us', 'dur us'))
for nm, st, a, b in sorted(pairs, key=lambda p: p[2]):
    sid = streams.setdefault(st, len(streams))
    print('%-52s %6d %9.1f %9.1f %9.1f' % (nm, sid, (a - t0) / 1e3, (b - t0) / 1e3, (b - a) / 1e3))
