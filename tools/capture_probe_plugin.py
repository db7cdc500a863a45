"""pytest plugin (python -m pytest -p tools.capture_probe_plugin ...): wraps every capi entry point with a probe that
reports the first call before / after which the current stream's capture is found invalidated."""
import types

import torch

from dgps_with_iwvi_b200 import capi

state = {'bad': False}


def probe(where):
    if state['bad']:
        return
    try:
        torch.cuda.is_current_stream_capturing()
    except Exception as e:  # noqa: BLE001
        state['bad'] = True
        print('\nCAPTURE INVALID %s: %s' % (where, str(e).splitlines()[0]), flush=True)


for _name in dir(capi):
    _fn = getattr(capi, _name)
    if isinstance(_fn, types.FunctionType) and not _name.startswith('_') and _name not in ('with_flags',):
        def _mk(fn, name):
            def w(*a, **k):
                probe('before ' + name)
                try:
                    return fn(*a, **k)
                finally:
                    probe('after ' + name)
            return w
        setattr(capi, _name, _mk(_fn, _name))
