"""ctypes binding of the C ABI declared in include/iwvi_b200.h (lib/libiwvi_b200.so, built by build.py).

There is no CPU fallback: importing the product on a machine where the library is missing raises, and calling a
compute entry point without a CUDA device raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('IWVI_B200_LIB') or os.path.join(HERE, 'lib', 'libiwvi_b200.so')

MAX_ENC_LAYERS = 8

KERN_IDS = {'RBF': 0, 'Matern12': 1, 'Matern32': 2, 'Matern52': 3}
MF_IDS = {'Zero': 0, 'Identity': 1, 'Linear': 2}
ACT_IDS = {'tanh': 0, 'relu': 1, 'sigmoid': 2, 'softplus': 3, 'elu': 4, 'identity': 5}
FLAG_SAMPLE, FLAG_SAVE, FLAG_ACCUM = 1, 2, 4
FLAG_ONLY_EPI, FLAG_ONLY_TILE, FLAG_ONLY_REDUCE, FLAG_ONLY_FINAL = 16, 32, 64, 128
FLAG_ONLY_GRAM = 8
FLAG_PART_A, FLAG_PART_B, FLAG_SKIP_KL, FLAG_ONLY_KL = 256, 512, 1024, 2048
FLAG_NO_KDIAG = 4096
FLAG_TWO_CHAINS = 8192
FLAG_PRO_HYP, FLAG_PRO_Q = 16384, 32768
FLAG_FAST_REDUCE = 65536

ERRORS = {-1: 'bad descriptor', -2: 'unsupported size', -3: 'CUDA launch failure', -4: 'null pointer'}


class GpDesc(C.Structure):
    _fields_ = [('T', C.c_int32), ('M', C.c_int32), ('D', C.c_int32), ('R', C.c_int32), ('P', C.c_int32),
                ('kern', C.c_int32), ('mix', C.c_int32), ('mf', C.c_int32), ('flags', C.c_int32),
                ('reserved', C.c_int32), ('jitter', C.c_double)]


class LvDesc(C.Structure):
    _fields_ = [('Be', C.c_int32), ('Kt', C.c_int32), ('Df', C.c_int32), ('Dxy', C.c_int32), ('Lw', C.c_int32),
                ('n_layers', C.c_int32), ('dims', C.c_int32 * (MAX_ENC_LAYERS + 1)), ('sampled', C.c_int32),
                ('f_bcast', C.c_int32), ('prior', C.c_int32), ('act', C.c_int32), ('reserved', C.c_int32),
                ('prior_mu', C.c_double), ('prior_sigma', C.c_double)]


class ElboDesc(C.Structure):
    _fields_ = [('B', C.c_int32), ('K', C.c_int32), ('Dy', C.c_int32), ('Lw', C.c_int32), ('iw', C.c_int32),
                ('data_major', C.c_int32), ('scale', C.c_double)]


P = C.c_void_p  # device pointers and the stream travel as opaque addresses

SIGNATURES = {
    'iwvi_version': (C.c_int, []),
    'iwvi_last_cuda_error': (C.c_char_p, []),
    'iwvi_gp_mp': (C.c_int32, [C.c_int32]),
    'iwvi_gp_lda': (C.c_int32, [C.c_int32]),
    'iwvi_gp_aux_doubles': (C.c_int64, [C.POINTER(GpDesc)]),
    'iwvi_gp_save_doubles': (C.c_int64, [C.POINTER(GpDesc)]),
    'iwvi_gp_bwd_ws_doubles': (C.c_int64, [C.POINTER(GpDesc)]),
    'iwvi_gp_pbwd_ws_doubles': (C.c_int64, [C.POINTER(GpDesc)]),
    'iwvi_gp_prologue_fwd': (C.c_int, [C.POINTER(GpDesc)] + [P] * 10),
    'iwvi_gp_rows_fwd': (C.c_int, [C.POINTER(GpDesc)] + [P] * 12),
    'iwvi_gp_rows_fwd_range': (C.c_int, [C.POINTER(GpDesc)] + [P] * 11 + [C.c_int64, C.c_int64, P]),
    'iwvi_gp_tile_points': (C.c_int, [C.POINTER(GpDesc)]),
    'iwvi_gp_rows_bwd': (C.c_int, [C.POINTER(GpDesc)] + [P] * 23),
    'iwvi_gp_rows_bwd_range': (C.c_int, [C.POINTER(GpDesc)] + [P] * 22 + [C.c_int64, C.c_int64, P]),
    'iwvi_gp_bwd_tile_points': (C.c_int, [C.POINTER(GpDesc)]),
    'iwvi_gp_prologue_bwd': (C.c_int, [C.POINTER(GpDesc)] + [P] * 16),
    'iwvi_gp_fullcov_fwd': (C.c_int, [C.POINTER(GpDesc), C.c_int32, C.c_int32] + [P] * 5 + [C.c_double] + [P] * 5),
    'iwvi_gp_fullcov_ws_doubles': (C.c_int64, [C.POINTER(GpDesc), C.c_int32, C.c_int32]),
    'iwvi_gp_fullcov_bwd': (C.c_int, [C.POINTER(GpDesc), C.c_int32, C.c_int32] + [P] * 4 + [C.c_double] + [P] * 7),
    'iwvi_gauss_kl_fwd': (C.c_int, [C.c_int32, C.c_int32] + [P] * 4),
    'iwvi_gauss_kl_bwd': (C.c_int, [C.c_int32, C.c_int32] + [P] * 6),
    'iwvi_lv_param_doubles': (C.c_int64, [C.POINTER(LvDesc)]),
    'iwvi_lv_bwd_ws_doubles': (C.c_int64, [C.POINTER(LvDesc)]),
    'iwvi_lv_fwd': (C.c_int, [C.POINTER(LvDesc)] + [P] * 9),
    'iwvi_lv_bwd': (C.c_int, [C.POINTER(LvDesc)] + [P] * 14),
    'iwvi_elbo_ws_doubles': (C.c_int64, [C.POINTER(ElboDesc)]),
    'iwvi_iwelbo_fwd': (C.c_int, [C.POINTER(ElboDesc)] + [P] * 10),
    'iwvi_iwelbo_bwd': (C.c_int, [C.POINTER(ElboDesc)] + [P] * 12),
    'iwvi_normal_fill': (C.c_int, [P, C.c_int64, C.c_int32, C.c_int64, C.c_uint64, P]),
    'iwvi_normal_fill_counter': (C.c_int, [P, C.c_int64, C.c_int32, C.c_int64, C.c_uint64, C.c_int32, C.c_int64, P, P]),
    'iwvi_adam_step_counter': (C.c_int, [P] * 6 + [C.c_int64, C.c_int64, P, C.c_double, C.c_double, C.c_double, P, P]),
    'iwvi_adam_step_counter_part': (C.c_int, [P] * 6 + [C.c_int64, C.c_int64, P, C.c_double, C.c_double, C.c_double, P,
                                               C.c_int32, P]),
    'iwvi_positive_fwd': (C.c_int, [P, P, C.c_int64, P]),
    'iwvi_batch_gather': (C.c_int, [P, P, P, C.c_int32, C.c_int32, C.c_int32, P, P, P, P]),
    'iwvi_dp_push': (C.c_int, [P, P, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, P, P, P, P, P]),
    'iwvi_dp_reduce': (C.c_int, [P, P, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, P, P, P, P, P]),
    'iwvi_debug_stamp': (C.c_int, [P, C.c_int32, P]),
    'iwvi_probe_dmma': (C.c_int, [P, C.c_int32, C.c_int32, C.c_int32, P]),
    'iwvi_adam_step': (C.c_int, [P] * 6 + [C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double,
                                          C.c_int64, P]),
}

_lib = None


def load():
    """Loads the shared library (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                'dgps_with_iwvi_b200: %s is missing -- build it with `python -m dgps_with_iwvi_b200.build` '
                '(or __graft_entry__.build()); there is no CPU fallback' % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        detail = ''
        if rc == -3 and _lib is not None:
            detail = ' [%s]' % _lib.iwvi_last_cuda_error().decode()
        raise RuntimeError('%s failed: %s (code %d)%s' % (what, ERRORS.get(rc, 'unknown error'), rc, detail))
