"""Finds the call that invalidates a stream capture: every capi entry point is wrapped with a capture-status probe."""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgps_with_iwvi_b200 import capi  # noqa: E402
from oracle import synthetic as S  # noqa: E402

state = {'bad': False}


def probe(where):
    if state['bad']:
        return
    try:
        torch.cuda.is_current_stream_capturing()
    except Exception as e:  # noqa: BLE001
        state['bad'] = True
        print('capture invalid %s: %s' % (where, str(e).splitlines()[0]), flush=True)


for name in dir(capi):
    fn = getattr(capi, name)
    if isinstance(fn, types.FunctionType) and not name.startswith('_') and name not in ('with_flags',):
        def mk(fn, name):
            def w(*a, **k):
                probe('before ' + name)
                try:
                    return fn(*a, **k)
                finally:
                    probe('after ' + name)
            return w
        setattr(capi, name, mk(fn, name))

from dgps_with_iwvi_b200.build_models import build_model  # noqa: E402
from dgps_with_iwvi_b200.training import Trainer  # noqa: E402

N, D, B, K = 400, 3, 32, 4
X, Y = S.make_data(N, D, seed=8)
model = build_model(X, Y, 'L1_G2', M=24, num_IW_samples=K, minibatch_size=B, mode='IWAE', seed=2)
tr = Trainer(model, B, lr=1e-2, lr_decay=0.98, seed=9, use_graph=True)
for i in range(4):
    tr.step(X[:B], Y[:B])
print('graphs', tr._graphs is not None)
Z = model.layers[1].feature.feat.Z
Z.set_trainable(False)
tr.step(X[:B], Y[:B])
print('step after set_trainable ok; graphs', tr._graphs is not None, 'pro_ready', tr._pro_ready)
try:
    tr.step(X[:B], Y[:B])
    print('second ok')
except Exception as e:  # noqa: BLE001
    print('FAILED', str(e).splitlines()[0])
