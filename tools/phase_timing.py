"""Per-phase cycle breakdown of gp_rows_fwd_kernel / gp_tile_bwd_kernel (thread 0 of every CTA, clock64()).
Build once with `python -m dgps_with_iwvi_b200.build --timing`, then run on the GPU box:
    IWVI_B200_LIB=dgps_with_iwvi_b200/lib/libiwvi_b200_timing.so python tools/phase_timing.py c3"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dgps_with_iwvi_b200 import _lib  # noqa: E402
from dgps_with_iwvi_b200.build_models import build_model  # noqa: E402
from dgps_with_iwvi_b200.training import Trainer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'c3'
cfg = bench.CONFIGS[name]
X, Y = bench.make_data(cfg['N'], cfg['D'], seed=0)
model = build_model(X, Y, cfg['configuration'], M=cfg['M'], num_IW_samples=cfg['K'], minibatch_size=cfg['B'],
                    likelihood_variance=cfg['lik_variance'], mode='IWAE', seed=0)
tr = Trainer(model, cfg['B'])
lib = C.CDLL(_lib.LIB_PATH)
buf = (C.c_ulonglong * 48)()
steps = 5
for i in range(steps + 1):
    idx = torch.arange(i * cfg['B'], (i + 1) * cfg['B'], device=model.X.device) % cfg['N']
    tr.step_device(model.X[idx], model.Y[idx])
    if i == 0:
        lib.iwvi_debug_phase_cycles(buf)   # discard the warm-up step
lib.iwvi_debug_phase_cycles(buf)
FWD = ['x tile', 'G gram+kernel fn', 'T forward subst', 'S fvar0/gmean + A save', 'U products + U save', 'E epilogue']
BWD = ['tile prologue', 'part 1 (A stream)', 'part 2 (Lq V)', 'back substitution', 'Bbar store', 'gram adjoint (loop exit)',
       'dX final', 'start-of-tile barrier', 'ga: Zt wait + gram GEMM', 'ga: kernel fn, G, row/col sums', 'ga: barrier',
       'ga: dX DMMAs', 'ga: dZ DMMAs + adds']
CHOL = ['gram block', 'updates (incl. waits for other rows)', 'diagonal factorisation', 'inversion + copies', 'wait for Dinv_k',
        'panel product / zero fill']
for k, (title, names) in enumerate([('gp_rows_fwd_kernel', FWD), ('gp_tile_bwd_kernel', BWD),
                                    ('gp_chol_kernel (last block row only)', CHOL)]):
    vals = [buf[k * 16 + j] for j in range(16)]
    tot = sum(vals)
    print('%s: all layers, %d steps, %.1f Mcycles on thread 0 of every CTA' % (title, steps, tot / 1e6))
    for j, n in enumerate(names):
        print('   %-26s %6.2f%%' % (n, 100.0 * vals[j] / max(tot, 1)))
