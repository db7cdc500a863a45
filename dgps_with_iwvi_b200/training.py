"""Training step around the hot path: IW-ELBO forward + backward (engine.Engine), ONE all-reduce of the flat float64
gradient bucket across the data-parallel ranks (NCCL over NVLink on GPUs; the last slot of the bucket carries the
ELBO), and the fused Adam + positive-transform kernel.  This is the Adam half of the reference's training iteration
(experiments/build_models.py:284-300); minibatch rows are sharded contiguously across ranks, noise is keyed by the
global row index so the result does not depend on the number of GPUs (up to summation order)."""
import torch
import torch.distributed as dist

from . import capi
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


class Trainer:
    def __init__(self, model, B_local, lr=5e-3, lr_decay=0.98, beta1=0.9, beta2=0.999, eps=1e-8, seed=0,
                 process_group=None):
        self.model = model
        self.pg = process_group
        self.distributed = dist.is_available() and dist.is_initialized()
        self.world_size = dist.get_world_size(process_group) if self.distributed else 1
        self.rank = dist.get_rank(process_group) if self.distributed else 0
        self.B_local = int(B_local)
        self.engine = model.engine(self.B_local, model.num_samples, None, self.world_size, self.rank)
        self.flat = FlatParams.of(model)
        self.flat.refresh_mask()
        self.m = torch.zeros_like(self.flat.x)
        self.v = torch.zeros_like(self.flat.x)
        self.lr, self.lr_decay = lr, lr_decay
        self.betas, self.eps = (beta1, beta2), eps
        self.seed = seed
        self.t = 0
        self.launches_per_step = None

    def step_device(self, X_local, Y_local, row0_global=None):
        """One training step on this rank's rows; returns the global ELBO as a 1-element device tensor (no sync)."""
        self.t += 1
        row0 = self.rank * self.B_local if row0_global is None else row0_global
        self.engine.elbo_and_grads(X_local, Y_local, None, seed=self.seed, step=self.t, row0=row0)
        if self.world_size > 1:
            dist.all_reduce(self.flat.g, op=dist.ReduceOp.SUM, group=self.pg)
        f = self.flat
        lr = staircase_decay(self.lr, self.t - 1, 1000, self.lr_decay)
        capi.adam_step(f.x, f.g, self.m, self.v, f.mask, f.theta_pos, f.n, f.n_pos, lr, self.betas[0], self.betas[1],
                       self.eps, self.t)
        return f.loss_slot

    def step(self, X_host, Y_host):
        """End-to-end call: host (pinned) minibatch in, ELBO (python float) out."""
        return float(self.step_device(X_host, Y_host).item())
