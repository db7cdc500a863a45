"""Builds the C-ABI shared library (include/iwvi_b200.h) from csrc/*.cu for sm_100a, in-tree.

    python -m dgps_with_iwvi_b200.build [--force]

nvcc cross-compiles without a GPU.  The resulting lib/libiwvi_b200.so is git-ignored but travels with the
repository snapshot to the GPU box."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libiwvi_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden']


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps():
    hdr = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    hdr.append(os.path.join(os.path.dirname(HERE), 'include', 'iwvi_b200.h'))
    return hdr


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, timing=False, variant=None, extra=()):
    """timing=True: a second library, lib/libiwvi_b200_timing.so, compiled with -DIWVI_PHASE_TIMING (per-phase clock64()
    totals, tools/phase_timing.py); select it with the environment variable IWVI_B200_LIB.
    variant='name', extra=['-DX=1', ...]: a tuning build lib/libiwvi_b200_<name>.so with extra nvcc flags."""
    os.makedirs(LIBDIR, exist_ok=True)
    if timing:
        variant = 'timing'
    objdir = os.path.join(LIBDIR, 'obj_' + variant if variant else 'obj')
    os.makedirs(objdir, exist_ok=True)
    srcs = sources()
    hdrs = _deps()
    jobs = []
    objs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + '.o')
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append([NVCC] + FLAGS + (['-DIWVI_PHASE_TIMING', '-rdc=true'] if timing else []) + list(extra) +
                        (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed: %s\n%s\n%s' % (' '.join(cmd), r.stdout, r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        for out in ex.map(run, jobs):
            if verbose and out:
                print(out)
    lib = LIB.replace('.so', '_%s.so' % variant) if variant else LIB
    if jobs or force or _stale(lib, objs):
        run([NVCC, '-shared', '-o', lib] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
    return lib


if __name__ == '__main__':
    variant = sys.argv[sys.argv.index('--variant') + 1] if '--variant' in sys.argv else None
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, timing='--timing' in sys.argv, variant=variant,
                extra=[a for a in sys.argv[1:] if a.startswith('-D')]))
