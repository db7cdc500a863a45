"""Training step around the hot path: IW-ELBO forward + backward (engine.Engine), ONE all-reduce of the flat float64
gradient bucket across the data-parallel ranks (NCCL over NVLink on GPUs; the last slot of the bucket carries the
ELBO), and the fused Adam + positive-transform kernel.  This is the Adam half of the reference's training iteration
(experiments/build_models.py:284-300); minibatch rows are sharded contiguously across ranks, noise is keyed by the
global row index so the result does not depend on the number of GPUs (up to summation order)."""
import torch
import torch.distributed as dist

from . import _lib as LIB
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
    """The exchange step of data-parallel training: all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests) of the
    PACKED gradient bucket.  What crosses the wire: trainable entries only (frozen W / mean-function / kernel-variance
    slots never form a gradient, build_models.py:209,213,225-227); of every LowerTriangular parameter only the lower
    triangle (GPflow stores q_sqrt packed; the strict upper triangle of the dense [R, M, M] storage gets an exactly-zero
    gradient); `always_reduce` parameters whatever their flag (the NatGrad half needs the last layer's q_mu / q_sqrt
    gradients, which Adam's mask excludes); and the ELBO slot.  At c3: 0.37 M doubles instead of the dense 0.74 M.

    `groups` = [(tag, [parameters])]: the packed buffer is laid out segment by segment in that order, the remaining
    parameters and the ELBO slot last (tag 'final'), so that a segment can be all-reduced on its own as soon as the backward
    pass has completed it (allreduce(tag)) -- or the whole bucket in one call (allreduce())."""

    def __init__(self, flat, always_reduce=(), process_group=None, groups=(), p2p=False):
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
        seg_of = torch.full((flat.n,), len(groups), dtype=torch.int64, device=flat.device)     # default: 'final'
        for gi, (_, params) in enumerate(groups):
            for p in params:
                _, o, sz, _, _ = flat.entries[id(p)]
                seg_of[o:o + sz] = gi
        parts, self.segments, off = [], {}, 0
        for gi, tag in enumerate([t for t, _ in groups] + ['final']):
            idx = torch.nonzero(keep & (seg_of == gi)).flatten()
            if tag == 'final':
                idx = torch.cat([idx, torch.tensor([flat.n], dtype=idx.dtype, device=flat.device)])
            parts.append(idx)
            self.segments[tag] = (off, off + idx.numel())
            off += idx.numel()
        self.index = torch.cat(parts).contiguous()
        self.buf = torch.zeros(self.index.numel(), dtype=torch.float64, device=flat.device)
        self.seg_ids = {tag: i for i, tag in enumerate(self.segments)}
        self.p2p = None
        self.p2p_error = None
        if p2p and flat.device.type == 'cuda' and dist.is_available() and dist.is_initialized() \
                and dist.get_world_size(process_group) > 1 and dist.get_backend(process_group) == 'nccl':
            self._setup_p2p()

    def _setup_p2p(self):
        """One-shot all-reduce over NVLink peer memory (csrc/dp_exchange.cu): a symmetric receive buffer
        [2][world][bucket] + flag words per (segment, rank), mapped into every peer by torch's symmetric memory
        (allocation + handle exchange only; the data path is iwvi_dp_push / iwvi_dp_reduce).  Falls back to NCCL, and
        says so in `p2p_error`, if the mapping cannot be established."""
        import ctypes as C
        world, rank = dist.get_world_size(self.pg), dist.get_rank(self.pg)
        n, nseg = self.index.numel(), len(self.segments) + 1            # (+1: the whole bucket as one segment)
        ok = torch.ones(1, device=self.flat.device)
        try:
            if world > 16:
                raise RuntimeError('more than IWVI_DP_MAX_RANKS ranks')
            import torch.distributed._symmetric_memory as symm_mem
            recv_doubles = 2 * world * n
            t = symm_mem.empty(recv_doubles + nseg * world, dtype=torch.float64, device=self.flat.device)
            hdl = symm_mem.rendezvous(t, self.pg if self.pg is not None else dist.group.WORLD)
            t.zero_()
            ptrs = [int(p) for p in hdl.buffer_ptrs]
        except Exception as e:  # noqa: BLE001
            self.p2p_error = '%s: %s' % (type(e).__name__, str(e).splitlines()[0] if str(e) else '')
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.pg)       # all ranks or none (and: everybody has zeroed)
        torch.cuda.synchronize()
        if ok.item() == 0:
            if self.p2p_error is None:
                self.p2p_error = 'a peer could not map the symmetric buffer'
            return
        self.p2p = {
            'world': world, 'rank': rank, 'n': n, 'buf': t, 'hdl': hdl,
            'recv_ptrs': (C.c_uint64 * world)(*ptrs),
            'flag_ptrs': (C.c_uint64 * world)(*[p + 8 * recv_doubles for p in ptrs]),
            'recv_local': ptrs[rank], 'flags_local': ptrs[rank] + 8 * recv_doubles,
            'epoch': torch.zeros(nseg, dtype=torch.int64, device=self.flat.device),
            'counters': torch.zeros(2 * nseg, dtype=torch.int32, device=self.flat.device)}

    def allreduce(self, tag=None):
        """pack -> all_reduce(SUM) -> unpack of one segment (or of the whole bucket), in stream order on the current
        stream."""
        a, b = (0, self.index.numel()) if tag is None else self.segments[tag]
        if a == b:
            return
        g, idx, buf = self.flat.g, self.index[a:b], self.buf[a:b]
        if self.p2p is not None:
            # pack + transfer + rank-ordered sum + unpack in two launches over the peers' memory
            P = self.p2p
            seg = len(self.segments) if tag is None else self.seg_ids[tag]
            capi.dp_push(g, idx, b - a, a, P['n'], seg, P['rank'], P['world'], P['recv_ptrs'], P['flag_ptrs'],
                         P['epoch'], P['counters'])
            capi.dp_reduce(g, idx, b - a, a, P['n'], seg, P['world'], P['recv_local'], P['flags_local'],
                           P['epoch'], P['counters'])
            return
        torch.index_select(g, 0, idx, out=buf)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg)
        g.index_copy_(0, idx, buf)


class Trainer:
    """use_graph=True (default): after two eager steps the whole step -- noise, forward, backward, gradient exchange,
    optimiser -- is captured ONCE as a CUDA graph and replayed; the step count and learning rate live in device memory
    (iwvi_normal_fill_counter / iwvi_adam_step_counter) so every replay draws fresh noise and applies the right bias
    correction.

    Several ranks: the packed gradient bucket is exchanged in segments, each all-reduced (NCCL) on the side stream that
    completed it while the backward pass of the layers below is still running (overlap_comm): the upper GP layers as
    soon as their Cholesky / gram adjoints are done, the first GP layer's q_mu / q_sqrt while its own adjoint chain runs;
    only the small remainder (first-layer Z / kernel parameters, encoder, likelihood variance, ELBO slot) is exchanged
    after the backward pass.  The NCCL calls are captured into the step graph (graph_comm); if the capture is refused
    the all-reduce falls back to ONE eager call of the whole bucket between two graphs."""

    def __init__(self, model, B_local, lr=5e-3, lr_decay=0.98, beta1=0.9, beta2=0.999, eps=1e-8, seed=0,
                 process_group=None, use_graph=True, always_reduce=None, overlap_comm=True, graph_comm=True,
                 distributed=None, pipeline=True):
        self.model = model
        self.pg = process_group
        self.distributed = (dist.is_available() and dist.is_initialized()) if distributed is None else bool(distributed)
        self.world_size = dist.get_world_size(process_group) if self.distributed else 1
        self.rank = dist.get_rank(process_group) if self.distributed else 0
        self.B_local = int(B_local)
        self.engine = model.engine(self.B_local, model.num_samples, None, self.world_size, self.rank)
        self.flat = FlatParams.of(model)
        self.flat.refresh_mask()
        self._trainable = tuple(p.trainable for p in self.flat.params)
        self.always_reduce = list(always_reduce or [])
        self.graph_comm = bool(graph_comm)
        # segments are issued from side streams inside the backward pass: with CUDA graphs that needs the capture
        self.overlap_comm = bool(overlap_comm) and self.world_size > 1 and (self.graph_comm or not use_graph)
        # Segment-wise step: every segment of parameters is updated (Adam) on the side stream that completed -- and, on
        # several ranks, exchanged -- its gradients, and the NEXT step's once-per-layer stage (Cholesky of Kuu, padded
        # copies, KL) of that layer follows at once, overlapped with the rest of the backward pass, instead of opening the
        # next step with ~115 us of single-CTA chains.  Needs the segments' gradients to be final when the hook fires:
        # one rank, or overlap_comm.
        self.pipeline = bool(pipeline) and (self.world_size == 1 or self.overlap_comm)
        self._pro_ready = False        # the engine's once-per-layer stages match the current parameters
        self.m = torch.zeros_like(self.flat.x)
        self.v = torch.zeros_like(self.flat.x)
        self._build_bucket()
        self.lr, self.lr_decay = lr, lr_decay
        self.betas, self.eps = (beta1, beta2), eps
        self.seed = seed
        self.t = 0
        dev = self.flat.device
        self.state = torch.zeros(2, dtype=torch.int64, device=dev)          # [0] = optimiser steps completed
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float64, device=dev)
        self._lr_host = float(lr)
        self.use_graph = bool(use_graph)
        self._updating = False
        self._graphs = None            # (graph_fwd_bwd, graph_update) once captured
        self._graph_launches = 0
        self._graph_row0 = None
        self.eager_steps_before_capture = 2

    # ---- the exchange step (packed gradient bucket, in segments) and the segment-wise update ----
    def _build_bucket(self):
        groups = []
        if self.overlap_comm or self.pipeline:
            gps = [r for r in self.engine.recs if r['type'] == 'gp']
            for r in reversed(gps):                      # the order in which the backward pass completes them
                layer, base, feat = r['layer'], r['base'], r['feat']
                hyp = [feat.Z, base.lengthscales, base.variance]
                if r['mix']:
                    hyp.append(layer.kern.W)
                if r['mf'] == 'Linear':
                    hyp += [layer.mean_function.A, layer.mean_function.b]
                if r['gi'] == 0:
                    groups.append((('gp_h', 0), hyp))
                    groups.append((('gp_q', 0), [layer.q_mu, layer.q_sqrt]))
                else:
                    groups.append((('gp', r['gi']), hyp + [layer.q_mu, layer.q_sqrt]))
        import os
        self.gbucket = GradBucket(self.flat, self.always_reduce, self.pg, groups, p2p=self.world_size > 1 and not os.environ.get('IWVI_DP_NCCL'))
        # Adam masks: trainable entries of each segment (the packed index of a segment also holds always_reduce entries
        # and, for 'final', the ELBO slot: the trainable mask filters those out)
        f = self.flat
        self.seg_mask = {}
        for tag, (a, b) in self.gbucket.segments.items():
            idx = self.gbucket.index[a:b]
            idx = idx[idx < f.n]
            mk = torch.zeros_like(f.mask)
            mk[idx] = f.mask[idx]
            self.seg_mask[tag] = mk
        self._gp_by_gi = {r['gi']: r for r in self.engine.recs if r['type'] == 'gp'}
        self.engine.grad_hook = self._segment_done if (self.overlap_comm or self.pipeline) else None
        self._pro_ready = False

    def _adam(self, mask, advance):
        f = self.flat
        capi.adam_step_counter_part(f.x, f.g, self.m, self.v, mask, f.theta_pos, f.n, f.n_pos, self.lr_dev,
                                    self.betas[0], self.betas[1], self.eps, self.state, advance)

    def _segment_done(self, tag):
        """Engine.backward calls this on the side stream that has just completed the gradients of segment `tag`."""
        if self.overlap_comm:
            self.gbucket.allreduce(tag)
        if not self.pipeline or not self._updating:
            return
        self._adam(self.seg_mask[tag], advance=False)
        eng = self.engine
        torch.cuda.current_stream().wait_event(eng.ev_loss)      # this step's loss assembly has read the KLs
        r = self._gp_by_gi[tag[1]]
        eng.prologue_layer(r, {'gp': 0, 'gp_h': LIB.FLAG_PRO_HYP, 'gp_q': LIB.FLAG_PRO_Q}[tag[0]])

    def allreduce_grads(self):
        """What is left to exchange after backward(): everything, or with overlap_comm the 'final' segment."""
        if self.world_size > 1:
            self.gbucket.allreduce('final' if self.overlap_comm else None)

    # ---- the two halves of a step, written against device-side state only (capturable) ----
    def _fwd_bwd(self, row0):
        eng = self.engine
        eng.draw_noise(None, seed=self.seed, row0=row0, state=self.state)
        if self.pipeline and not self._pro_ready:
            # first step (or parameters were changed from outside): the once-per-layer stages, all at once
            eng.forward(join=False, prologue=True)
        else:
            eng.forward(join=False, prologue=not self.pipeline)
        self._updating = True
        eng.backward()
        self._updating = False

    def _update(self):
        if self.pipeline:
            self._adam(self.seg_mask['final'], advance=True)      # encoder, likelihood variance; closes the step
            self._pro_ready = True
        else:
            self._adam(self.flat.mask, advance=True)

    def invalidate(self):
        """Call after changing parameter values from outside the trainer (assignment, loading a checkpoint): the next
        step recomputes the once-per-layer stages instead of using the ones prepared at the end of the previous step."""
        self._pro_ready = False
        self._graphs = None

    def _capture(self, row0):
        # The cyclic garbage collector must not run inside the capture: a finaliser that destroys CUDA objects of an
        # earlier engine (graphs, streams, cached plans) is an operation the global capture mode forbids, and it
        # invalidates the capture at whatever launch happens to come next.
        import gc
        gc.collect()
        was_enabled = gc.isenabled()
        gc.disable()
        try:
            self._capture_impl(row0)
        finally:
            if was_enabled:
                gc.enable()

    def _capture_impl(self, row0):
        l0 = capi.LAUNCHES
        torch.cuda.synchronize()
        self._graphs = None
        if self.world_size == 1 or self.graph_comm:
            # the whole step, exchange included, as ONE graph
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._fwd_bwd(row0)
                    self.allreduce_grads()
                    self._update()
                self._graphs = (g, None)
            except RuntimeError as e:          # NCCL refused the capture: two graphs around an eager all-reduce
                if self.world_size == 1:
                    raise
                import warnings
                warnings.warn('dgps_with_iwvi_b200: NCCL all-reduce could not be captured into the step graph (%s); '
                              'falling back to an eager all-reduce between two graphs' % str(e).splitlines()[0])
                torch.cuda.synchronize()
                self.graph_comm = False
                if self.overlap_comm:          # segments issued from side streams need the capture: one bucket instead
                    self.overlap_comm = False
                    self.pipeline = False
                    self._build_bucket()
                capi.LAUNCHES = l0
        if self._graphs is None:
            ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(ga):
                self._fwd_bwd(row0)
            with torch.cuda.graph(gb, pool=ga.pool()):
                self._update()
            self._graphs = (ga, gb)
        self._graph_launches = capi.LAUNCHES - l0
        self._graph_row0 = row0
        capi.LAUNCHES = l0             # capture launched nothing; replays are counted in step_device

    def step_indices(self, idx_local, row0_global=None):
        """step_device with this rank's minibatch rows given as int64 device indices into the model's resident data
        (model.X, model.Y): the gather is one launch instead of torch indexing + copies."""
        return self.step_device(None, None, row0_global, idx=idx_local)

    def step_device(self, X_local, Y_local, row0_global=None, idx=None, data=None):
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
        if idx is not None:
            Xd, Yd = data if data is not None else (self.model.X, self.model.Y)
            self.engine.set_batch_indices(Xd, Yd, idx)
        else:
            self.engine.set_batch(X_local, Y_local)
        # (the graph is captured once the pipelined once-per-layer stages are in place: a step that still has to open
        #  with them -- the first one, or after invalidate() -- runs eagerly)
        graph = self.use_graph and self.t > self.eager_steps_before_capture and (not self.pipeline or self._pro_ready)
        if graph and (self._graphs is None or self._graph_row0 != row0):
            self._capture(row0)
        if graph and self._graphs[1] is None:          # one graph: noise, forward, backward, exchange, optimiser
            self._graphs[0].replay()
            capi.LAUNCHES += self._graph_launches
            return self.flat.loss_slot
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

    def step_pipelined(self, X_host, Y_host):
        """End-to-end call for a training LOOP: host (pinned) minibatch in, ELBO (python float) of the PREVIOUS call out
        (None on the first call; `flush()` returns the last one).  The host does not wait for the step it has just
        enqueued: the minibatch goes to a device staging buffer on a copy stream while the previous step is still
        running, one gather launch moves it into the plan's buffers when that step has finished, and the ELBO comes
        back through a pinned word that is read one call later.  Every step still pays its own host-to-device copy
        and its own device-to-host read; what is hidden is their latency and the launch latency of the step graph.
        The caller may reuse X_host / Y_host as soon as the NEXT call has returned."""
        if getattr(self, '_pipe', None) is None:
            eng, dev = self.engine, self.flat.device
            z = lambda *sh: torch.zeros(*sh, dtype=torch.float64, device=dev)
            self._pipe = {
                'X': [z(eng.B, eng.Dx) for _ in range(2)], 'Y': [z(eng.B, eng.Dy) for _ in range(2)],
                'loss': [torch.zeros(1, dtype=torch.float64).pin_memory() for _ in range(2)],
                'h2d': [torch.cuda.Event() for _ in range(2)], 'free': [torch.cuda.Event() for _ in range(2)],
                'done': [torch.cuda.Event() for _ in range(2)], 'stream': torch.cuda.Stream(device=dev),
                'idx': torch.arange(eng.B, dtype=torch.int64, device=dev), 'i': 0}
        P = self._pipe
        i = P['i']
        k = i % 2
        main = torch.cuda.current_stream()
        cs = P['stream']
        if i >= 2:
            cs.wait_event(P['free'][k])            # the gather of call i - 2 has read this staging buffer
        with torch.cuda.stream(cs):
            P['X'][k].copy_(torch.as_tensor(X_host, dtype=torch.float64).reshape(P['X'][k].shape), non_blocking=True)
            P['Y'][k].copy_(torch.as_tensor(Y_host, dtype=torch.float64).reshape(P['Y'][k].shape), non_blocking=True)
            P['h2d'][k].record(cs)
        main.wait_event(P['h2d'][k])
        loss = self.step_device(None, None, idx=P['idx'], data=(P['X'][k], P['Y'][k]))
        P['free'][k].record(main)
        P['loss'][k].copy_(loss, non_blocking=True)
        P['done'][k].record(main)
        P['i'] = i + 1
        if i == 0:
            return None
        P['done'][1 - k].synchronize()
        return float(P['loss'][1 - k][0])

    def flush(self):
        """ELBO of the last step_pipelined call (waits for it)."""
        P = getattr(self, '_pipe', None)
        if P is None or P['i'] == 0:
            return None
        k = (P['i'] - 1) % 2
        P['done'][k].synchronize()
        return float(P['loss'][k][0])


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
        kw['pipeline'] = False          # the NatGrad half rewrites q(u) of the last layer between two Adam evaluations
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
