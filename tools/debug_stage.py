"""Debug aid: run tests/test_gpu_stages.test_gp_layer_stages for one shape printing every output's error."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_gpu_stages as TS

def close(name, got, want, rtol=1e-8):
    got = got.detach().cpu().numpy() if torch.is_tensor(got) else np.asarray(got)
    want = np.asarray(want)
    err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-300)
    bad = np.argwhere(~(np.abs(got - want) <= rtol * np.abs(want).max()))
    print('%-12s err %.3e  nbad %d of %d  first bad %s' % (name, err, len(bad), got.size, bad[:6].tolist()))
TS.close = close
for a in sys.argv[1:] or ['1']:
    shape = TS.SHAPES[int(a)] if a.isdigit() else eval(a)
    print('shape', shape)
    TS.test_gp_layer_stages(shape)
