#!/usr/bin/env python
"""IW-ELBO training throughput on B200 (BASELINE.json metric: IW-ELBO train steps/s and KxN samples/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3] [--impl b200|reference]

One step = IW-ELBO forward + hand-written backward on one minibatch + (N>1) one NCCL all-reduce of the flat fp64
gradient bucket + fused Adam update.  Synthetic data of the BASELINE config shape; float64 throughout.
Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for the flop/byte conventions."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs -> (configuration, N, D, M, K, B per GPU, lik_variance)
    'c1': dict(configuration='L1', N=200, D=1, M=50, K=20, B=200, lik_variance=0.1),
    'c2': dict(configuration='L1_G5', N=10000, D=8, M=100, K=20, B=512, lik_variance=0.01),
    'c3': dict(configuration='L1_G5_G5', N=100000, D=16, M=256, K=50, B=512, lik_variance=0.01),
    'c4': dict(configuration='L1_G5', N=16384, D=8, M=512, K=256, B=4096, lik_variance=0.01),
    'c5': dict(configuration='L1_G5_G5', N=1000000, D=8, M=256, K=50, B=512, lik_variance=0.01),
}
# Fallback FP64 tensor-pipe peak (profiles/FP64_PEAK_r01.md).  The figure the roofline uses is measured IN THE RUN by
# measure_fp64_peak() (DMMA issue-rate probe of the library + cuBLAS DGEMM); MEASURED_PEAKS.json has no fp64 entry.
FP64_DMMA_PEAK_TFLOPS = 37.05


def measure_fp64_peak(torch, capi, dev):
    """FP64 roofline denominator from this run: (a) the register-only DMMA.8x8x4 issue loop (csrc/probe.cu), one CTA of
    8 warps per SM, best of 5; (b) cuBLAS DGEMM 4096^3 through torch.matmul, best of 5.  CUDA events, after a warm-up."""
    nsm = torch.cuda.get_device_properties(dev).multi_processor_count
    out = torch.empty(nsm * 4 * 8 * 32, dtype=torch.float64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def best(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        b = 1e30
        for _ in range(reps):
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            b = min(b, e0.elapsed_time(e1))
        return b

    iters = 20000
    fl = [0.0]

    def probe():
        fl[0] = capi.probe_dmma(out, nsm * 4, 8, iters)
    ms = best(probe)
    dmma = fl[0] / (ms * 1e-3) / 1e12
    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty(n, n, dtype=torch.float64, device=dev)
    ms = best(lambda: torch.matmul(a, b, out=c))
    dgemm = 2.0 * n ** 3 / (ms * 1e-3) / 1e12
    return {'dmma_probe_tflops': dmma, 'cublas_dgemm_tflops': dgemm,
            'how': 'mma.sync.m8n8k4.f64 register-only issue loop, %d CTAs x 8 warps x %d iters x 8 accumulators, and '
                   'torch.matmul fp64 %d^3; CUDA events, best of 5, measured in this run' % (nsm * 4, iters, n)}


def _hbm_peak():
    """Measured HBM copy bandwidth of this pool's B200s (driver-written MEASURED_PEAKS.json), else the profiling recipe's
    fallback figure."""
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs'])
    except (OSError, KeyError, ValueError):
        return 6650.0   # "of fallback" (B200_PROFILING.md)


HBM_PEAK_GBS = _hbm_peak()


_STDOUT = None


def emit(obj):
    f = _STDOUT or sys.stdout
    f.write(json.dumps(obj) + '\n')
    f.flush()


def make_data(N, D, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D))
    w = rng.standard_normal((D, 1)) / np.sqrt(D)
    Y = np.sin(X @ w) + 0.1 * rng.standard_normal((N, 1))
    return X, Y


def layer_shapes(configuration, D):
    """[(D_in, R)] for every GP layer of a configuration string (experiments/build_models.py:201-241)."""
    out = []
    D_in = D
    for tok in [t for t in configuration.split('_') if t]:
        c, d = tok[0], int(tok[1:])
        if c == 'L':
            D_in += d
        else:
            out.append((D_in, d))
            D_in = D
    out.append((D_in, 1))
    return out


def fwd_flops(T, M, D, R):
    """SURVEY.md 8(d) triangular-aware convention, per GP layer forward."""
    return T * ((1 + R) * M * M + 2 * M * (D + 2 * R + 1))


def step_flops(cfg, B):
    T = B * cfg['K']
    f = sum(fwd_flops(T, cfg['M'], D, R) for D, R in layer_shapes(cfg['configuration'], cfg['D']))
    return 3 * f


class NvmlSampler:
    """In-process NVML polling (every ~4 ms) of SM clock, power and clock-event reasons.  The thread runs from before the
    warm-up; only samples whose timestamp falls inside [begin(), end()] -- the timed region -- are reported, so even a
    timed region of a few tens of milliseconds is covered (a freshly spawned nvidia-smi needs ~100 ms for its first line)."""
    MASKS = {'sw_power_cap': 0x4, 'hw_slowdown': 0x8, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40}

    def __init__(self, torch, local_rank):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        h = None
        try:
            uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + uuid) if not uuid.startswith('GPU-') else uuid)
        except Exception:
            h = None
        if h is None:
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            idx = int(vis.split(',')[local_rank]) if vis and all(v.strip().isdigit() for v in vis.split(',')) else local_rank
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        self.h = h
        self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        self.samples, self.stop_flag, self.t0, self.t1 = [], False, None, None
        self.thread = threading.Thread(target=self._poll, daemon=True)

    def start(self):
        self.thread.start()

    def _poll(self):
        nv, h = self.nv, self.h
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.samples.append((time.perf_counter(), sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.004)

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self.stop_flag = True
        self.thread.join(timeout=1)
        win = [x for x in self.samples if self.t0 is not None and self.t0 <= x[0] <= (self.t1 or 1e300)]
        if not win:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_sm, 'reasons': ['no samples']}
        reasons = sorted(n for n, m in self.MASKS.items() if any(x[3] & m for x in win))
        return {'sm_mhz': float(np.median([x[1] for x in win])), 'sm_max_mhz': self.max_sm,
                'power_w_max': float(max(x[2] for x in win)), 'samples': len(win), 'reasons': reasons,
                'source': 'nvml, 4 ms polling inside the timed region'}


class ClockSampler:
    def begin(self):
        pass

    def end(self):
        pass

    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '25'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'power_w_max': float(max(pw)),
                'samples': len(sm), 'reasons': sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (float64 torch-CPU restatement of the reference, op for op) -- the reference itself needs
# TensorFlow 1.x + GPflow 1.x, which cannot be installed here.
# ----------------------------------------------------------------------------------------------------------------
def cpu_eval_time(cfg, spec, X, Y, B_sample, reps, warm=1, reference_style=True):
    import torch
    from oracle import iwvi_oracle as O
    from oracle import synthetic as S
    K = cfg['K']
    Xb, Yb = X[:B_sample], Y[:B_sample]
    eps = S.make_noise(spec, (B_sample, K), seed=0)
    times = []
    for i in range(warm + reps):
        t0 = time.perf_counter()
        O.iw_elbo_and_grads(spec, Xb, Yb, eps, reference_style=reference_style)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    return times, torch.get_num_threads()


def cpu_sample_rows(cfg):
    """Rows of the minibatch the CPU arm evaluates per step, bounded so that one evaluation stays within seconds and
    the materialised [B,R,M,K] / [B,R,K,K] tensors of the reference formulation stay within host memory."""
    per_row = cfg['K'] * cfg['M'] * 6 * 8 * 6 + 2 * cfg['K'] * cfg['K'] * 8 * 6
    budget = 6e9
    return int(max(8, min(cfg['B'], budget // per_row)))


def oracle_spec(cfg, X, Y):
    from oracle import synthetic as S
    return S.make_spec(X, cfg['configuration'], cfg['M'], cfg['K'], lik_variance=cfg['lik_variance'], seed=0, perturb=0.0)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm runs on rank 0 alone and is meant to use every host
    # core the box gives this process, so the thread pools are sized from the affinity mask before torch is imported
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    for var in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[var] = str(cores)
    import torch
    torch.set_num_threads(cores)
    cfg = CONFIGS[args.config]
    X, Y = make_data(min(cfg['N'], 20000), cfg['D'], seed=0)
    spec = oracle_spec(cfg, X, Y)
    spec['num_data'] = cfg['N']
    Bs = cpu_sample_rows(cfg)
    times, threads = cpu_eval_time(cfg, spec, X, Y, Bs, reps=args.steps, warm=max(args.warmup, 1))
    sec = float(np.mean(times))
    value = Bs * cfg['K'] / sec
    out = {
        'impl': 'reference', 'metric': 'iw_elbo_train_KxN_samples_per_s', 'value': value, 'unit': 'KxN samples/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3,
        'steps_per_s': (Bs / cfg['B']) / sec, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.config, cfg, args.gpus),
        'arm': arm_description('reference', args.gpus),
        'cpu_baseline': {'value': value, 'unit': 'KxN samples/s', 'cores': threads, 'kind': 'port',
                         'sample': 'oracle/iwvi_oracle.py (torch-CPU fp64 op-for-op restatement, reference-style KxK final '
                                   'layer) forward+autograd on %d of %d minibatch rows x K=%d, no optimiser step; the '
                                   'reference needs TF1/GPflow1 which cannot be installed here' % (Bs, cfg['B'], cfg['K'])},
        'e2e': {'value': value, 'unit': 'KxN samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    emit(out)


def workload_config(name, cfg, gpus):
    """The workload -- identical for both arms, and nothing but the workload (what each arm DOES with it is `arm`)."""
    return {'workload': '%s: %s N=%d D=%d M=%d K=%d B=%d/GPU (BASELINE.json configs)' % (
        name, cfg['configuration'], cfg['N'], cfg['D'], cfg['M'], cfg['K'], cfg['B']),
        'global_batch_rows': cfg['B'] * gpus, 'points_per_gpu_step': cfg['B'] * cfg['K'],
        'step': 'IW-ELBO forward + backward over one minibatch of B rows x K importance samples',
        'l2': 'inputs larger than L2: each step streams its own A/U panels (about 1 GB at c3 against 126 MB of L2) and '
              'rewrites every buffer; no explicit flush'}


def arm_description(arm, gpus, exchange=None):
    if arm == 'reference':
        return {'parallelism': 'none: rank 0 alone, all host cores (torch intra-op threads)',
                'optimizer': 'none (forward + autograd of the IW-ELBO only: this arm does LESS work per step than the GPU arm)',
                'launch': 'oracle/iwvi_oracle.py, float64 torch-CPU op-for-op restatement of the reference (TF1/GPflow1 '
                          'cannot be installed here); each step = one evaluation on a bounded row sample of the minibatch'}
    return {'parallelism': 'dp%d rows sharded; per-layer segments of the packed fp64 gradient bucket exchanged inside the step '
                           'graph, %s' % (gpus, exchange or 'no exchange at one rank'),
            'optimizer': 'adam (fused kernel, segment-wise as each layer\'s gradients are final)',
            'launch': 'whole step replayed as a CUDA graph'}


# ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--config', default='c3', choices=sorted(CONFIGS))
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-reference-iteration', action='store_true',
                    help='skip the extra timing of the reference-faithful iteration (NatGrad + Adam, N=1 only)')
    ap.add_argument('--global-rows', type=int, default=0,
                    help='strong scaling: fix the GLOBAL minibatch (rows) and split it over the GPUs (SURVEY.md 8(d): 4096 '
                         'for c3); default 0 = weak scaling with the config\'s rows per GPU')
    args = ap.parse_args()
    # the contract is ONE JSON line on stdout: libraries that chat on stdout (NCCL prints its version there) are sent to
    # stderr for the duration of the run, and the line is written to the saved descriptor
    global _STDOUT
    sys.stdout.flush()
    _STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # the gradient segments are exchanged while compute kernels with hundreds of queued CTAs are running: NCCL's
        # kernels go on a high-priority stream so that they are dispatched as soon as SMs free up
        try:
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank), pg_options=opts)
        except (AttributeError, TypeError):
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from dgps_with_iwvi_b200 import capi
    from dgps_with_iwvi_b200 import _lib as LIB
    from dgps_with_iwvi_b200.build_models import build_model, spec_from_model
    from dgps_with_iwvi_b200.models import Minibatch
    from dgps_with_iwvi_b200.training import Trainer

    cfg = dict(CONFIGS[args.config])
    strong = args.global_rows > 0
    if strong:
        if args.global_rows % world:
            raise SystemExit('bench.py: --global-rows %d is not divisible by %d GPUs' % (args.global_rows, world))
        cfg['B'] = args.global_rows // world
    B, K, N = cfg['B'], cfg['K'], cfg['N']
    Bg = B * world
    X, Y = make_data(N, cfg['D'], seed=0)
    model = build_model(X, Y, cfg['configuration'], M=cfg['M'], num_IW_samples=K, minibatch_size=Bg,
                        likelihood_variance=cfg['lik_variance'], mode='IWAE', seed=0)
    if os.environ.get('IWVI_FAST_REDUCE'):          # quick experiments; the reported fast-path line is fast_path_timing()
        model.fast_reduce = True
    trainer = Trainer(model, B, lr=5e-3, seed=0)
    eng = trainer.engine
    dev = eng.dev
    stream_idx = Minibatch(N, Bg, seed=0)          # same stream on every rank; each takes its contiguous slice

    def next_idx():
        return stream_idx.next()[rank * B:(rank + 1) * B]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident inputs: the dataset lives in HBM, minibatches are gathered there ----
    Xd, Yd = model.X, model.Y

    # minibatch indices of every resident-mode step are uploaded once: a per-step host->device copy of pageable memory
    # would block the host until the previous step has drained and expose the launch latency of the next one
    n_res = args.warmup + args.steps
    idx_all = torch.as_tensor(np.stack([next_idx() for _ in range(n_res)]), device=dev)

    def step_resident(i):
        return trainer.step_indices(idx_all[i])          # rows gathered on the device by iwvi_batch_gather

    # The clock sampler is constructed and started BEFORE the warm-up, on rank 0, so that nothing host-side happens on
    # any rank between the pre-timing barrier and ev0.record(): a rank that reaches its first all-reduce while another
    # is still setting up would wait there, and the bench takes the MAX over ranks (round-1 SCALE lost 28 % to this).
    sampler = None
    if rank == 0:
        try:
            sampler = NvmlSampler(torch, local_rank)
        except Exception:
            sampler = ClockSampler(local_rank)      # nvidia-smi -lms 25 in a subprocess
        sampler.start()
    for i in range(args.warmup):
        step_resident(i)
    barrier()
    l0 = capi.LAUNCHES
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler is not None:
        sampler.begin()
    ev0.record()
    for i in range(args.steps):
        loss = step_resident(args.warmup + i)
    ev1.record()
    barrier()
    if sampler is not None:
        sampler.end()
    ms = ev0.elapsed_time(ev1)
    launches = capi.LAUNCHES - l0
    clocks = sampler.stop() if sampler is not None else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    eng.check_info()
    last_elbo = float(loss.item())

    # ---- end to end: pinned host minibatch in, ELBO (float) out, every step ----
    Xh, Yh = torch.as_tensor(X).pin_memory(), torch.as_tensor(Y).pin_memory()
    stage = [(torch.empty(B, cfg['D'], dtype=torch.float64).pin_memory(), torch.empty(B, 1, dtype=torch.float64).pin_memory())
             for _ in range(2)]

    stage.append((torch.empty_like(stage[0][0]).pin_memory(), torch.empty_like(stage[0][1]).pin_memory()))
    e2e_elbos = []

    def step_e2e(i):
        idx = torch.as_tensor(next_idx())
        xs, ys = stage[i % 3]              # (a host buffer is free again once the NEXT call has returned: three in rotation)
        torch.index_select(Xh, 0, idx, out=xs)
        torch.index_select(Yh, 0, idx, out=ys)
        # the training-loop API: this step's H2D copies + step + D2H of its ELBO are all enqueued here; the float that
        # comes back is the PREVIOUS step's ELBO (read from pinned memory), so the host never idles the GPU
        return trainer.step_pipelined(xs, ys)

    for i in range(max(args.warmup // 2, 3)):
        step_e2e(i)
    trainer.flush()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for i in range(args.steps):
        v = step_e2e(i)
        if i > 0:                          # (call 0 hands back the last warm-up step's ELBO)
            e2e_elbos.append(v)
    e2e_elbos.append(trainer.flush())      # the last step's ELBO is read inside the timed region too
    ev1.record()
    barrier()
    ms_e2e = max(ev0.elapsed_time(ev1), 0.0)
    t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())

    # ---- N > 1, outside the timed region: the data-parallel evaluation (row shards + ONE all-reduce of the packed
    # bucket, exactly the Trainer's exchange) against the same global minibatch evaluated by rank 0 alone ----
    dp_parity = dp_parity_check(trainer, model, Xd, Yd, B, K, world, rank, torch, dist) if world > 1 else None

    if rank != 0:
        teardown(trainer, world, torch, dist)
        return

    # ---- per-kernel timing of the three DMMA kernels of one GP layer (CUDA events on the launching stream) ----
    peak = measure_fp64_peak(torch, capi, dev)
    peak_tf = peak['dmma_probe_tflops']
    kern = kernel_timings(eng, capi, LIB, torch, peak_tf)
    sec = ms / 1e3 / args.steps
    value = Bg * K / sec
    flops = step_flops(cfg, B)
    dom = max(kern, key=lambda k: k['ms'] * k['launches_per_step'])
    # self-check: the end-to-end step does strictly more than the resident one (H2D + D2H every step), so a resident
    # time ABOVE it means something host-side leaked into the resident region
    ok_order = ms <= ms_e2e * 1.02
    out = {
        'metric': 'iw_elbo_train_KxN_samples_per_s', 'value': value, 'unit': 'KxN samples/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'steps_per_s': 1.0 / sec,
        'higher_is_better': True, 'scaling': 'strong' if strong else 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': workload_config(args.config, cfg, world),
        'arm': arm_description('b200', world, None if world == 1 else (
            'one-shot all-reduce over NVLink peer memory (iwvi_dp_push / iwvi_dp_reduce: pack, transfer, rank-ordered sum and '
            'unpack in two launches)' if trainer.gbucket.p2p is not None else
            'NCCL all-reduce between a pack and an unpack launch (peer mapping unavailable: %s)' % trainer.gbucket.p2p_error)),
        'clocks': clocks,
        'e2e': {'value': Bg * K / (ms_e2e / 1e3 / args.steps), 'unit': 'KxN samples/s',
                'ms_per_step': ms_e2e / args.steps, 'h2d_bytes_per_step': B * (cfg['D'] + 1) * 8 + 0,
                'd2h_bytes_per_step': 8,
                'how': 'Trainer.step_pipelined: pinned host minibatch -> device staging (copy stream) -> one gather launch -> '
                       'step graph -> ELBO to a pinned word; every step has its own H2D copy and its own D2H read, the float '
                       'is handed to the caller one call later (flush() for the last one, inside the timed region)',
                'elbos_read': len(e2e_elbos)},
        'gpu_launches': launches,
        'self_check': {'resident_le_e2e': bool(ok_order), 'resident_ms': ms / args.steps, 'e2e_ms': ms_e2e / args.steps},
        'step_algorithmic_gflop_per_gpu': flops / 1e9,
        'step_tflops_per_gpu': flops / sec / 1e12,
        'step_frac_of_fp64_dmma_peak': flops / sec / 1e12 / peak_tf,
        'fp64_peak': peak,
        'roofline': {'bound': 'tensor', 'kernel': dom['kernel'], 'achieved': dom['tflops'], 'peak': peak_tf,
                     'unit': 'TFLOP/s', 'frac': dom['tflops'] / peak_tf,
                     'traffic': ncu_traffic(args.config, dom['kernel']),
                     'traffic_scope': 'dram bytes of ONE launch in the committed ncu capture (profiles/ncu_%s_summary.json): the '
                                      'per-point kernels run there as the chain of full waves (296 of 400 tiles at c3), '
                                      'the timed launch above covers all tiles' % args.config,
                     'peak_source': 'FP64 DMMA issue-rate probe measured in this run (fp64_peak; round-1 pool figure '
                                    '%.2f); MEASURED_PEAKS.json carries only bf16 and HBM' % FP64_DMMA_PEAK_TFLOPS},
        'kernels': kern,
        'hbm_kernels': hbm_kernel_timings(eng, capi, LIB, torch, HBM_PEAK_GBS),
        'hbm_peak_gbs': HBM_PEAK_GBS,
        'elbo_last': last_elbo,
    }
    if dp_parity is not None:
        out['dp_parity_max_err'] = dp_parity['bucket_max_err_over_max']
        out['dp_parity'] = dp_parity
    if not ok_order:
        sys.stderr.write('bench.py: SELF-CHECK FAILED: resident %.3f ms/step > end-to-end %.3f ms/step\n'
                         % (ms / args.steps, ms_e2e / args.steps))
    if world == 1 and not args.no_cpu_baseline:
        spec = spec_from_model(model)
        Bs = cpu_sample_rows(cfg)
        t_probe, threads = cpu_eval_time(cfg, spec, X, Y, Bs, reps=1, warm=1)
        reps = int(min(30, max(3, 12.0 / max(t_probe[0], 1e-3))))      # about 10-15 s of CPU work in total
        times, threads = cpu_eval_time(cfg, spec, X, Y, Bs, reps=reps, warm=0)
        v = Bs * K / float(np.mean(times))
        out['cpu_baseline'] = {'value': v, 'unit': 'KxN samples/s', 'cores': threads, 'kind': 'port',
                               'ms_per_eval': float(np.mean(times)) * 1e3,
                               'sample': 'oracle (torch-CPU fp64 restatement of the reference, KxK final layer) '
                                         'forward+autograd on %d of %d minibatch rows x K=%d, %d evals after warm-up, no '
                                         'optimiser step' % (Bs, B, K, reps)}
        # the same oracle with the final layer's variance taken directly (the maths the GPU path runs)
        t_diag, _ = cpu_eval_time(cfg, spec, X, Y, Bs, reps=3, warm=1, reference_style=False)
        out['cpu_baseline']['value_diag_only'] = Bs * K / float(np.mean(t_diag))
    if world == 1 and not args.no_reference_iteration:
        out['reference_iteration'] = reference_iteration_timing(cfg, X, Y, torch)
    emit(out)
    teardown(trainer, world, torch, dist)


def teardown(trainer, world, torch, dist):
    """Leaves the process group in order.  The step graph holds captured NCCL kernels: it is released before the
    communicator is destroyed, and because a rank may wait in ncclCommDestroy for a peer that is still measuring (rank 0
    times kernels after the other ranks are done), a watchdog ends the process -- the JSON line is already out -- if the
    teardown has not finished within 20 s."""
    if world <= 1:
        return
    sys.stdout.flush()
    sys.stderr.flush()
    threading.Timer(20.0, lambda: os._exit(0)).start()
    trainer._graphs = None
    torch.cuda.synchronize()
    try:
        dist.destroy_process_group()
    finally:
        os._exit(0)


def dp_parity_check(trainer, model, Xd, Yd, B, K, world, rank, torch, dist):
    """Correctness figure of the N-GPU line, outside the timed region: every rank evaluates its row shard of ONE fixed
    global minibatch and the Trainer's exchange (packed bucket, one NCCL all-reduce) runs; rank 0 then evaluates the same
    global minibatch alone (world size 1) and reports max|dp - single| / max|single| over the exchanged gradient entries
    and the relative ELBO difference.  Noise is keyed by the global point index, so the two must agree to summation order."""
    from dgps_with_iwvi_b200.models import Minibatch
    eng, flat = trainer.engine, trainer.flat
    N = Xd.shape[0]
    Bg = B * world
    gidx = torch.as_tensor(Minibatch(N, Bg, seed=123).next(), device=Xd.device)
    mine = gidx[rank * B:(rank + 1) * B]
    eng.elbo_and_grads(Xd[mine], Yd[mine], None, seed=77, step=1, row0=rank * B)
    trainer.allreduce_grads()
    torch.cuda.synchronize()
    live = trainer.gbucket.index
    g_dp = flat.g[live].clone()
    out = None
    if rank == 0:
        eng1 = model.engine(Bg, K, None, 1, 0)
        eng1.elbo_and_grads(Xd[gidx], Yd[gidx], None, seed=77, step=1, row0=0)
        eng1.check_info()
        g_1 = flat.g[live].clone()
        scale = g_1[:-1].abs().max().item()
        out = {'bucket_max_err_over_max': (g_dp[:-1] - g_1[:-1]).abs().max().item() / scale,
               'elbo_rel_err': abs(g_dp[-1].item() - g_1[-1].item()) / abs(g_1[-1].item()),
               'elbo_dp': g_dp[-1].item(), 'elbo_single': g_1[-1].item(), 'entries': int(live.numel()),
               'what': 'dp%d (row shards + one all-reduce of the packed bucket) vs the same %d-row global minibatch on '
                       'rank 0 alone' % (world, Bg)}
        model.drop_engine(eng1)
        del eng1
        torch.cuda.empty_cache()
    dist.barrier()
    return out


def reference_iteration_timing(cfg, X, Y, torch, iters=20, warm=3):
    """the reference-faithful training iteration (experiments/build_models.py:284-300) -- a NatGrad step
    on the last layer's q(u) evaluated on one minibatch, then an Adam step on everything else evaluated on a second one:
    two IW-ELBO forward+backward passes per iteration.  Reported beside the headline (one pass + Adam per step)."""
    from dgps_with_iwvi_b200.build_models import build_model
    from dgps_with_iwvi_b200.models import Minibatch
    from dgps_with_iwvi_b200.training import ReferenceIterationTrainer
    B, K, N = cfg['B'], cfg['K'], cfg['N']
    model = build_model(X, Y, cfg['configuration'], M=cfg['M'], num_IW_samples=K, minibatch_size=B,
                        likelihood_variance=cfg['lik_variance'], mode='IWAE', seed=0)
    tr = ReferenceIterationTrainer(model, B, lr=5e-3, gamma=1e-2, seed=0)
    dev = tr.engine.dev
    mb = Minibatch(N, B, seed=1)
    idx = torch.as_tensor(np.stack([mb.next() for _ in range(2 * (iters + warm))]), device=dev)
    Xd, Yd = model.X, model.Y

    def it(i):
        a, b = idx[2 * i], idx[2 * i + 1]
        return tr.iteration(Xd[a], Yd[a], Xd[b], Yd[b])

    for i in range(warm):
        it(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        elbo_ng, elbo_adam = it(warm + i)
    e1.record()
    torch.cuda.synchronize()
    tr.engine.check_info()
    sec = e0.elapsed_time(e1) / 1e3 / iters
    return {'ms_per_iteration': sec * 1e3, 'iterations_per_s': 1.0 / sec, 'KxN_samples_per_s': 2 * B * K / sec,
            'iterations': iters, 'elbo_last': float(elbo_adam.item()),
            'what': 'NatGrad(gamma=1e-2) on the last GP layer\'s q_mu/q_sqrt + Adam(5e-3) on the rest, two minibatches and '
                    'two forward+backward passes per iteration (build_models.py:284-300); eager launches, no CUDA graph'}


def ncu_traffic(config, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture
    of this config (profiles/ncu_<config>_summary.json, written by tools/ncu_to_profiles.py); None if not captured."""
    path = os.path.join(ROOT, 'profiles', 'ncu_%s_summary.json' % config)
    try:
        with open(path) as f:
            return json.load(f)['kernels'][kernel]['dram_bytes_per_launch']
    except (OSError, KeyError, ValueError):
        return None


def kernel_timings(eng, capi, LIB, torch, peak_tf, reps=5):
    """Times gp_rows_fwd_kernel, gp_tile_bwd_kernel and gp_reduce_bwd_kernel of the widest inner GP layer alone, on the
    buffers the last step left behind.  Algorithmic flops per launch:
      rows_fwd : T[(1+R)M^2 + 2M(D+2R+1)]           tile_bwd : T[(1+R)M^2 + 2M(3D+R+1)]
      reduce   : T(1+R)M^2 + 2TMR   (lower triangles only)
    All three use SURVEY.md 8(d)'s triangular-aware count; the blocked algorithms execute more (diagonal blocks are
    only skipped at 8x8x4 granularity), so 1.0 is not reachable."""
    gps = [r for r in eng.recs if r['type'] == 'gp']
    r = max(gps, key=lambda q: q['R'] * q['M'] * q['M'])
    n_like = sum(1 for q in gps if (q['R'], q['M']) == (r['R'], r['M']))   # layers of the timed shape (the R=1 final layer is cheaper)
    T, M, D, R = eng.T, r['M'], r['D'], r['R']
    NB = r['Mp'] // 64
    flat, layer, base, feat = eng.flat, r['layer'], r['base'], r['feat']
    W = flat.cview(layer.kern.W) if r['mix'] else None
    lin = r['mf'] == 'Linear'
    mfA = flat.cview(layer.mean_function.A) if lin else None
    mfb = flat.cview(layer.mean_function.b) if lin else None
    scratch = [torch.zeros_like(flat.gview(p)) for p in (feat.Z, base.lengthscales, base.variance, layer.q_mu, layer.q_sqrt)]
    if scratch[1].numel() != D:
        scratch[1] = torch.zeros(D, dtype=torch.float64, device=eng.dev)
    dW = torch.zeros_like(W) if r['mix'] else None
    dA = torch.zeros_like(mfA) if lin else None
    db = torch.zeros_like(mfb) if lin else None
    d_s = torch.randn(T, r['P'], dtype=torch.float64, device=eng.dev)

    def fwd():
        capi.gp_rows_fwd(r['d'], r['Lm'], r['aux'], r['Fin'], W, mfA, mfb, r['eps'], r['sample'], r['mean'], r['var'],
                         r['save'])

    def bwd(flag):
        d = capi.with_flags(r['d'], r['d'].flags | flag)
        capi.gp_rows_bwd(d, r['Lm'], r['aux'], r['save'], r['Fin'], W, mfA, mfb, r['eps'],
                         d_s if r['sampled'] else None, None if r['sampled'] else d_s, None if r['sampled'] else d_s,
                         r['dX'], scratch[0], scratch[1], scratch[2], scratch[3], scratch[4], r['dLm'], dW, dA, db,
                         r['bwd_ws'])

    bwd(0)
    cases = [
        ('gp_rows_fwd_kernel', fwd, T * ((1 + R) * M * M + 2 * M * (D + 2 * R + 1))),
        ('gp_tile_bwd_kernel', lambda: bwd(LIB.FLAG_ONLY_TILE), T * ((1 + R) * M * M + 2 * M * (R + 1))),
        ('gp_gram_bwd_kernel', lambda: bwd(LIB.FLAG_ONLY_GRAM), T * 2 * M * 3 * D),
        ('gp_reduce_bwd_kernel', lambda: bwd(64), T * (1 + R) * M * M + 2 * T * M * R),
    ]
    out = []
    for name, fn, fl in cases:
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out.append({'kernel': name, 'layer': 'M=%d D=%d R=%d T=%d' % (M, D, R, T), 'ms': ms, 'gflop': fl / 1e9,
                    'tflops': fl / (ms * 1e-3) / 1e12, 'frac_of_fp64_dmma_peak': fl / (ms * 1e-3) / 1e12 / peak_tf,
                    'launches_per_step': n_like})
    return out


def hbm_kernel_timings(eng, capi, LIB, torch, hbm_gbs, reps=20):
    """The fused elementwise / reduction stages of the step (HBM-bound), each timed alone with CUDA events
    on the buffers the last step left behind.  Algorithmic bytes = unique input + output bytes of the launch.  At c3 the
    working sets are 0.2-8 MB (L2-resident, a few microseconds per launch): the fractions say how far launch latency
    keeps these from the HBM copy peak, they are not where the step's time goes."""
    T, B, K = eng.T, eng.B, eng.K
    out = []

    def timed(name, fn, nbytes, note):
        # `reps` launches replayed as one CUDA graph: an eager loop of microsecond kernels times the host, not the GPU
        fn()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(reps):
                fn()
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        gbs = nbytes / (us * 1e-6) / 1e9
        out.append({'kernel': name, 'us': us, 'algorithmic_mb': nbytes / 1e6, 'gbs': gbs,
                    'frac_of_hbm_peak': gbs / hbm_gbs, 'bytes': note})

    lv = eng.lv_recs[0] if eng.lv_recs else None
    if lv is not None and lv.get('bcast'):
        Lw, Df, Dxy = lv['Lw'], lv['Df'], eng.Dx + eng.Dy
        timed('normal_fill_kernel', lambda: capi.normal_fill(lv['eps'], T, Lw, 0, 1234), 8 * T * Lw, '8*T*Lw written')
        timed('lv_fwd_kernel',
              lambda: capi.lv_fwd(lv['d'], eng.X, eng.XY, lv['params'], lv['eps'], lv['samples'], lv['kl'], lv['mu'],
                                  lv['sigma']),
              8 * (T * (Lw + Df + Lw + Lw) + B * (Df + Dxy + 2 * Lw)),
              '8*[T*(eps Lw + samples Df+Lw + kl Lw) + B*(X + XY + mu,sigma)]')
    gps = [r for r in eng.recs if r['type'] == 'gp']
    inner = [r for r in gps if r['sampled']]
    if inner:
        r = inner[-1]
        timed('normal_fill_kernel[T,R]', lambda: capi.normal_fill(r['eps'], T, r['R'], 0, 99), 8 * T * r['R'], '8*T*R written')
    last = gps[-1]
    lik = eng.flat.cview(eng.model.likelihood.variance)
    kl_local = eng.lv_recs[0]['kl'] if len(eng.lv_recs) == 1 else eng.kl_cat
    Lw_t, Dy = eng.Lw_total, eng.Dy
    timed('elbo_fwd_kernel(+final)',
          lambda: capi.iwelbo_fwd(eng.ed, last['mean'], last['var'], eng.Y, lik, kl_local, eng.elbo_data, eng.logp, eng.w,
                                  eng.elbo_ws),
          8 * (T * (2 * Dy + Lw_t + 1) + B * (Dy + 1)), '8*[T*(mean,var 2Dy + kl Lw + w 1) + B*(Y + logp)]')
    gl = torch.zeros_like(eng.flat.gview(eng.model.likelihood.variance))
    timed('elbo_bwd_kernel(+final)',
          lambda: capi.iwelbo_bwd(eng.ed, last['mean'], last['var'], eng.Y, lik, eng.w, eng.one, eng.dmean, eng.dvar,
                                  eng.dkl_local, gl, eng.elbo_ws),
          8 * (T * (2 * Dy + 1 + 2 * Dy + Lw_t) + B * Dy), '8*[T*(mean,var,w read; dmean,dvar,dkl written) + B*Y]')
    return out


if __name__ == '__main__':
    main()
