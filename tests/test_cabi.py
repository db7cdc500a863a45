"""The C-ABI library (include/iwvi_b200.h) loads without a GPU and exports every symbol the header declares; the
size helpers and argument validation work on the host.  No compute entry point launches here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'iwvi_b200.h')


@pytest.fixture(scope='module')
def lib():
    from dgps_with_iwvi_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(iwvi_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from dgps_with_iwvi_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    raw = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), 'header declares %s but the library does not export it' % n
        assert n in _lib.SIGNATURES, 'no ctypes signature for %s' % n
    assert set(_lib.SIGNATURES) <= set(names), set(_lib.SIGNATURES) - set(names)


def test_version_and_struct_layout(lib):
    from dgps_with_iwvi_b200 import _lib
    assert lib.iwvi_version() == 100
    assert C.sizeof(_lib.GpDesc) == 48 and C.sizeof(_lib.ElboDesc) == 32
    assert _lib.GpDesc.jitter.offset == 40 and _lib.ElboDesc.scale.offset == 24
    assert C.sizeof(_lib.LvDesc) == 4 * 6 + 4 * 9 + 4 * 5 + 16


def test_size_helpers_on_host(lib):
    from dgps_with_iwvi_b200 import capi
    assert capi.gp_mp(50) == 64 and capi.gp_mp(256) == 256 and capi.gp_mp(257) == 320
    d = capi.gp_desc(25600, 256, 17, 5, 16, 'RBF', True, 'Linear', 3)
    NB, blk = 4, 64 * 68
    npairs = NB * (NB + 1) // 2
    assert capi.gp_aux_doubles(d) == npairs * blk * (1 + 5) + 256 * 20 + 256 + 256 * 8 + 64 + (8 + 8 * 36) + 8
    assert capi.gp_save_doubles(d) == (25600 // 64) * NB * blk * 6 + 2 * 25600 * 5
    assert capi.gp_bwd_ws_doubles(d) > 0 and capi.gp_pbwd_ws_doubles(d) == 2 * 256 * 256 + 256 * 32 + 256
    e = capi.elbo_desc(512, 50, 1, 1, True, True, 100.0)
    assert capi.elbo_ws_doubles(e) > 0
    lv = capi.lv_desc(512, 50, 16, 17, 1, [17, 20, 20, 2], True, True)
    assert capi.lv_param_doubles(lv) == 17 * 20 + 20 + 20 * 20 + 20 + 20 * 2 + 2


def test_bad_descriptors_are_rejected_before_any_launch(lib):
    from dgps_with_iwvi_b200 import _lib, capi
    too_big = capi.gp_desc(10, 513, 4, 1, 1, 'RBF', False, 'Zero')
    assert lib.iwvi_gp_aux_doubles(C.byref(too_big)) == -1
    mismatch = capi.gp_desc(10, 16, 4, 2, 3, 'RBF', False, 'Zero')     # P != R without mixing
    assert lib.iwvi_gp_rows_fwd(C.byref(mismatch), *([None] * 12)) == -1
    ok = capi.gp_desc(10, 16, 4, 2, 2, 'RBF', False, 'Zero')
    assert lib.iwvi_gp_rows_fwd(C.byref(ok), *([None] * 12)) == -4          # null pointers
    assert lib.iwvi_gp_prologue_fwd(C.byref(ok), *([None] * 10)) == -4
    assert lib.iwvi_gauss_kl_fwd(0, 1, None, None, None, None) == -1
    with pytest.raises(RuntimeError):
        _lib.check(-2, 'x')


def test_product_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from dgps_with_iwvi_b200 import capi
    with pytest.raises(RuntimeError, match='no CPU path'):
        capi.positive_fwd(torch.zeros(4, dtype=torch.float64), torch.zeros(4, dtype=torch.float64), 4)
