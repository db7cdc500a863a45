"""Training step around the hot path: IW-ELBO forward + backward (engine.Engine), ONE all-reduce of the flat float64
gradient bucket across the data-parallel ranks (NCCL over NVLink on GPUs; the last slot of the bucket carries the
ELBO), and the fused Adam + positive-transform kernel.  This is the Adam half of the reference's training iteration
(experiments/build_models.py:284-300); minibatch rows are sharded contiguously across ranks, noise is keyed by the
global row index so the result does not depend on the number of GPUs (up to summation order)."""
import torch
import torch.distributed as dist

from . import capi
from . import natgrad as NG
from .engine import FlatParams


def shard_rows(B_global, world_size, rank):
    """Contiguous row range of `rank` in a global minibatch of B_global rows (must divide evenly)."""
    if B_global % world_size:
        raise ValueError('global minibatch %d is not divisible by world size %d' % (B_global, world_size))
    b = B_global // world_size
    return rank * b, (rank + 1) * b


def staircase_decay(base, step, decay_steps=1000, rate=0.98):
    """tf.train.exponential_decay(base, global_step, 1000, rate, staircase=True) (build_models.py:289-292)."""
    return base * rate ** (step // decay_steps)


class GradBucket:
    """The exchange step of data-parallel training: ONE all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests) of
    the PACKED gradient bucket.  What crosses the wire: trainable entries only (frozen W / mean-function / kernel-variance
    slots never form a gradient, build_models.py:209,213,225-227); of every LowerTriangular parameter only the lower
    triangle (GPflow stores q_sqrt packed; the strict upper triangle of the dense [R, M, M] storage gets an exactly-zero
    gradient); `always_reduce` parameters whatever their flag (the NatGrad half needs the last layer's q_mu / q_sqrt
    gradients, which Adam's mask excludes); and the ELBO slot.  At c3: 0.37 M doubles instead of the dense 0.74 M."""

    def __init__(self, flat, always_reduce=(), process_group=None):
        from .params import LowerTriangular
        self.flat, self.pg = flat, process_group
        flat.refresh_mask()
        keep = flat.mask != 0
        for p in always_reduce:
            _, o, sz, _, _ = flat.entries[id(p)]
            keep[o:o + sz] = True
        for p in flat.params:
            if isinstance(p.transform, LowerTriangular):
                _, o, sz, shape, _ = flat.entries[id(p)]
                tri = torch.ones(shape[-2:], dtype=torch.bool, device=flat.device).tril().expand(shape).reshape(-1)
                keep[o:o + sz] &= tri
        idx = torch.nonzero(keep).flatten()
        self.index = torch.cat([idx, torch.tensor([flat.n], dtype=idx.dtype, device=flat.device)]).contiguous()
        self.buf = torch.zeros(self.index.numel(), dtype=torch.float64, device=flat.device)

    def allreduce(self):
        """pack -> all_reduce(SUM) -> unpack, in stream order on the current stream."""
        g = self.flat.g
        torch.index_select(g, 0, self.index, out=self.buf)
        dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=self.pg)
        g.index_copy_(0, self.index, self.buf)


class Trainer:
    """use_graph=True (default): after two eager steps the whole step -- noise, forward, backward, optimiser -- is
    captured ONCE as a CUDA graph and replayed; the step count and learning rate live in device memory
    (iwvi_normal_fill_counter / iwvi_adam_step_counter) so every replay draws fresh noise and applies the right bias
    correction.  With several ranks the NCCL all-reduce stays an eager call between two graphs."""

    def __init__(self, model, B_local, lr=5e-3, lr_decay=0.98, beta1=0.9, beta2=0.999, eps=1e-8, seed=0,
                 process_group=None, use_graph=True, always_reduce=None):
        self.model = model
        self.pg = process_group
        self.distributed = dist.is_available() and dist.is_initialized()
        self.world_size = dist.get_world_size(process_group) if self.distributed else 1
        self.rank = dist.get_rank(process_group) if self.distributed else 0
        self.B_local = int(B_local)
        self.engine = model.engine(self.B_local, model.num_samples, None, self.world_size, self.rank)
        self.flat = FlatParams.of(model)
        self.flat.refresh_mask()
        self._trainable = tuple(p.trainable for p in self.flat.params)
        self.always_reduce = list(always_reduce or [])
        self._build_bucket()
        self.m = torch.zeros_like(self.flat.x)
        self.v = torch.zeros_like(self.flat.x)
        self.lr, self.lr_decay = lr, lr_decay
        self.betas, self.eps = (beta1, beta2), eps
        self.seed = seed
        self.t = 0
        dev = self.flat.device
        self.state = torch.zeros(2, dtype=torch.int64, device=dev)          # [0] = optimiser steps completed
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float64, device=dev)
        self._lr_host = float(lr)
        self.use_graph = bool(use_graph)
        self._graphs = None            # (graph_fwd_bwd, graph_update) once captured
        self._graph_launches = 0
        self._graph_row0 = None
        self.eager_steps_before_capture = 2

    # ---- the exchange step: ONE all-reduce of the PACKED gradient bucket ----
    def _build_bucket(self):
        self.gbucket = GradBucket(self.flat, self.always_reduce, self.pg)

    def allreduce_grads(self):
        if self.world_size > 1:
            self.gbucket.allreduce()

    # ---- the two halves of a step, written against device-side state only (capturable) ----
    def _fwd_bwd(self, row0):
        self.engine.draw_noise(None, seed=self.seed, row0=row0, state=self.state)
        self.engine.forward(join=False)
        self.engine.backward()

    def _update(self):
        f = self.flat
        capi.adam_step_counter(f.x, f.g, self.m, self.v, f.mask, f.theta_pos, f.n, f.n_pos, self.lr_dev, self.betas[0],
                               self.betas[1], self.eps, self.state)

    def _capture(self, row0):
        l0 = capi.LAUNCHES
        torch.cuda.synchronize()
        ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(ga):
            self._fwd_bwd(row0)
        with torch.cuda.graph(gb, pool=ga.pool()):
            self._update()
        self._graphs = (ga, gb)
        self._graph_launches = capi.LAUNCHES - l0
        self._graph_row0 = row0
        capi.LAUNCHES = l0             # capture launched nothing; replays are counted in step_device

    def step_device(self, X_local, Y_local, row0_global=None):
        """One training step on this rank's rows; returns the global ELBO as a 1-element device tensor (no sync)."""
        self.t += 1
        row0 = self.rank * self.B_local if row0_global is None else row0_global
        # the reference increments global_step BEFORE the optimiser ops of an iteration (build_models.py:297-300), so
        # iteration t = 1, 2, ... runs with exponential_decay(..., global_step=t, 1000, rate, staircase=True)
        lr = staircase_decay(self.lr, self.t, 1000, self.lr_decay)
        trainable = tuple(p.trainable for p in self.flat.params)
        if trainable != self._trainable:       # set_trainable() after construction: new mask, and the captured graphs
            self.flat.refresh_mask()           # (which skip the reductions of frozen W / mean-function slots) are stale
            self._trainable = trainable
            self._graphs = None
            self._build_bucket()
        if lr != self._lr_host:
            self.lr_dev.fill_(lr)
            self._lr_host = lr
        self.engine.set_batch(X_local, Y_local)
        graph = self.use_graph and self.t > self.eager_steps_before_capture
        if graph and (self._graphs is None or self._graph_row0 != row0):
            self._capture(row0)
        if graph:
            self._graphs[0].replay()
        else:
            self._fwd_bwd(row0)
        self.allreduce_grads()
        if graph:
            self._graphs[1].replay()
            capi.LAUNCHES += self._graph_launches
        else:
            self._update()
        return self.flat.loss_slot

    def step(self, X_host, Y_host):
        """End-to-end call: host (pinned) minibatch in, ELBO (python float) out."""
        return float(self.step_device(X_host, Y_host).item())


class ReferenceIterationTrainer(Trainer):
    """The reference's full training iteration (experiments/build_models.py:284-300): a natural-gradient step with rate
    gamma on the LAST GP layer's (q_mu, q_sqrt) evaluated on one minibatch, then an Adam step on every other trainable
    parameter evaluated on a second minibatch with fresh noise; both rates decay by `decay` every 1000 iterations
    (staircase).  Two ELBO forward+backward passes per iteration, as in the reference."""

    def __init__(self, model, B_local, lr=5e-3, gamma=1e-2, lr_decay=0.98, gamma_decay=0.98, **kw):
        from .layers import GPLayer
        last = [l for l in model.layers if isinstance(l, GPLayer)][-1]
        last.q_mu.set_trainable(False)          # handed to the natural-gradient optimiser (build_models.py:284-287)
        last.q_sqrt.set_trainable(False)
        kw['use_graph'] = False
        kw['always_reduce'] = [last.q_mu, last.q_sqrt]
        super().__init__(model, B_local, lr=lr, lr_decay=lr_decay, **kw)
        self.ng_layer = last
        self.gamma, self.gamma_decay = gamma, gamma_decay
        self.it = 0

    NG_STEP_BASE = 1 << 40   # noise-step offset of the NatGrad evaluations: never collides with the Adam evaluations

    def iteration(self, X1, Y1, X2, Y2):
        """Returns (ELBO of the NatGrad evaluation, ELBO of the Adam evaluation) as 1-element device tensors."""
        f, layer = self.flat, self.ng_layer
        row0 = self.rank * self.B_local
        # ---- NatGrad half
        self.engine.set_batch(X1, Y1)
        self.engine.draw_noise(None, seed=self.seed, step=self.NG_STEP_BASE + self.it, row0=row0)
        self.engine.forward()
        self.engine.backward()
        self.allreduce_grads()
        elbo_ng = f.loss_slot.clone()
        gamma = staircase_decay(self.gamma, self.it + 1, 1000, self.gamma_decay)   # global_step = iteration, 1-based
        mu_new, L_new = NG.natgrad_step(f.cview(layer.q_mu), f.cview(layer.q_sqrt), f.gview(layer.q_mu),
                                        f.gview(layer.q_sqrt), gamma)
        f.cview(layer.q_mu).copy_(mu_new)       # both are stored untransformed (q_sqrt: full array, tril on read)
        f.cview(layer.q_sqrt).copy_(L_new)
        # ---- Adam half (its own minibatch and noise); Trainer.step_device advances the step counter
        elbo_adam = self.step_device(X2, Y2)
        self.it += 1
        return elbo_ng, elbo_adam
