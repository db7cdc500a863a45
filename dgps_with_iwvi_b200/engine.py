"""Execution plan of the IW-ELBO hot path: one object per (model, B, K, mode) that owns every device buffer the C ABI
needs (the library itself never allocates) and strings the stages together, forward and backward, without autograd:

    forward : gp_prologue_fwd (all GP layers) -> [lv_fwd | gp_rows_fwd]* -> iwelbo_fwd
    backward: iwelbo_bwd -> [gp_rows_bwd + gp_prologue_bwd | lv_bwd]* (reversed)

This is reference models.py:112-150 (DGP_IWVI._build_likelihood), models.py:49-86 (DGP_VI) and models.py:93-107
(prediction) plus the tf.gradients call the optimisers make (build_models.py:293-295), on a flat parameter buffer:
gradients with respect to the *constrained* values land in one flat float64 bucket (`flat.g`) whose last slot carries
the ELBO itself, so that data-parallel training needs a single all-reduce (training.py)."""
import numpy as np
import torch

from . import _lib as LIB
from . import capi
from .params import Log1pe

F64 = torch.float64


def _ceil_even(n):
    return (n + 1) // 2 * 2


class FlatParams:
    """All parameters of a model in one unconstrained float64 buffer `x` (positive-transformed entries first),
    their constrained values (`theta_pos` for the positive block, `x` itself elsewhere), the gradient bucket `g`
    (+1 slot for the ELBO) and the trainable mask.  Parameter objects are re-bound to views of `x`."""

    def __init__(self, model):
        named = list(model.named_parameters())
        pos = [(n, p) for n, p in named if isinstance(p.transform, Log1pe)]
        rest = [(n, p) for n, p in named if not isinstance(p.transform, Log1pe)]
        self.entries = {}
        off = 0
        for n, p in pos:
            self.entries[id(p)] = (n, off, p.size, p.shape, True)
            off += p.size
        self.n_pos = off
        off = _ceil_even(off)
        for n, p in rest:
            self.entries[id(p)] = (n, off, p.size, p.shape, False)
            off = _ceil_even(off + p.size)
        self.n = off
        dev = named[0][1].unconstrained.device
        self.device = dev
        self.x = torch.zeros(self.n, dtype=F64, device=dev)
        self.theta_pos = torch.zeros(max(self.n_pos, 1), dtype=F64, device=dev)
        self.g = torch.zeros(self.n + 2, dtype=F64, device=dev)
        self.mask = torch.zeros(self.n, dtype=F64, device=dev)
        self.params = [p for _, p in pos + rest]
        for p in self.params:
            _, o, sz, shape, _ = self.entries[id(p)]
            p._rebind(self.x[o:o + sz].view(shape))
        self.refresh_mask()

    @staticmethod
    def of(model):
        flat = model.__dict__.get('_flat')
        if flat is None:
            flat = FlatParams(model)
            object.__setattr__(model, '_flat', flat)
        return flat

    def refresh_mask(self):
        self.mask.zero_()
        for p in self.params:
            _, o, sz, _, _ = self.entries[id(p)]
            if p.trainable:
                self.mask[o:o + sz] = 1.0

    def refresh_constrained(self):
        if self.n_pos:
            capi.positive_fwd(self.x, self.theta_pos, self.n_pos)

    def cview(self, p):
        """Constrained value of parameter p as a view (valid after refresh_constrained)."""
        _, o, sz, shape, is_pos = self.entries[id(p)]
        src = self.theta_pos if is_pos else self.x
        return src[o:o + sz].view(shape if shape else (1,))

    def gview(self, p):
        _, o, sz, shape, _ = self.entries[id(p)]
        return self.g[o:o + sz].view(shape if shape else (1,))

    @property
    def loss_slot(self):
        return self.g[self.n:self.n + 1]

    def grads_by_name(self):
        return {self.entries[id(p)][0]: self.gview(p).detach().clone() for p in self.params}


def layer_seed(seed, step, layer):
    return (int(seed) * 0x9E3779B97F4A7C15 + int(step) * 0xBF58476D1CE4E5B9 + (layer + 1) * 0x94D049BB133111EB) \
        & 0xFFFFFFFFFFFFFFFF


class Engine:
    """mode 'iw'      : DGP_IWVI._build_likelihood, points = [B, K] data-major (models.py:113-116)
       mode 'vi'      : DGP_VI._build_likelihood,   points = [S*B] sample-major (models.py:50-53), K := S
       mode 'predict' : propagate without amortisation inputs, points = [S, N] (models.py:93-107), B := N, K := S"""

    def __init__(self, model, B, K, mode='iw', world_size=1, rank=0, split_waves=True, fast_reduce=False):
        from .layers import GPLayer, LatentVariableLayer
        if not torch.cuda.is_available():
            raise RuntimeError('dgps_with_iwvi_b200: no CUDA device -- the IW-ELBO path has no CPU fallback')
        LIB.load()
        assert mode in ('iw', 'vi', 'predict')
        self.model, self.B, self.K, self.mode = model, int(B), int(K), mode
        self.world_size, self.rank = int(world_size), int(rank)
        # OPTIONAL reduced-precision fast path (off by default; never used by the parity tests): the parameter contractions of
        # the backward pass on tcgen05 / TMEM as 3xTF32 products (include/iwvi_b200.h, IWVI_FLAG_FAST_REDUCE)
        self.fast_reduce = bool(fast_reduce)
        self.T = T = self.B * self.K
        self.flat = flat = FlatParams.of(model)
        dev = flat.device
        self.dev = dev
        self.Dx, self.Dy = int(model.Dx), int(model.Dy)
        self.train = mode != 'predict'
        z = lambda *s: torch.zeros(*s, dtype=F64, device=dev)
        self.X = z(self.B, self.Dx)
        self.Y = z(self.B, self.Dy)
        self.XY = z(self.B, self.Dx + self.Dy)
        layers = list(model.layers)
        self.n_gp = sum(isinstance(l, GPLayer) for l in layers)
        self.kls = z(max(self.n_gp, 1))
        self.infos = torch.zeros(max(self.n_gp, 1), dtype=torch.int32, device=dev)
        self.one = torch.ones(1, dtype=F64, device=dev)
        self.dkl = torch.full((1,), -1.0 / self.world_size, dtype=F64, device=dev)
        self.recs = []
        D_cur = self.Dx
        gi = 0
        max_bwd_ws = max_pbwd_ws = max_lv_ws = 0
        for li, layer in enumerate(layers):
            last = li == len(layers) - 1
            if isinstance(layer, LatentVariableLayer):
                Lw = layer.latent_dim
                first = li == 0
                prior = mode == 'predict'
                bcast = first and mode == 'iw'
                Be, Kt = (self.B, self.K) if bcast else (T, 1)
                enc = layer.encoder
                d = capi.lv_desc(Be, Kt, D_cur, self.Dx + self.Dy, Lw, None if prior else enc.layer_dims,
                                 sampled=(mode == 'iw'), f_bcast=bcast, prior=prior,
                                 prior_mu=layer.prior_mu, prior_sigma=layer.prior_sigma,
                                 act=getattr(enc, 'activation_func', 'tanh'))
                r = dict(type='lv', layer=layer, d=d, idx=li, Lw=Lw, Df=D_cur, bcast=bcast, Be=Be, first=first,
                         samples=z(T, D_cur + Lw), kl=z(T, Lw), eps=z(T, Lw), mu=z(Be, Lw), sigma=z(Be, Lw),
                         enc_in=None if (bcast or prior) else z(T, self.Dx + self.Dy))
                if not prior:
                    ps = [p for pair in zip(enc.Ws, enc.bs) for p in pair]
                    offs = [flat.entries[id(p)][1] for p in ps]
                    sizes = [_ceil_even(flat.entries[id(p)][2]) for p in ps]
                    n_par = capi.lv_param_doubles(d)
                    contiguous = all(offs[i] + sizes[i] == offs[i + 1] for i in range(len(ps) - 1)) and \
                        all(flat.entries[id(p)][2] % 2 == 0 for p in ps[:-1])
                    r['packed_view'] = contiguous
                    if contiguous:
                        r['params'] = flat.x[offs[0]:offs[0] + n_par]
                        r['d_params'] = flat.g[offs[0]:offs[0] + n_par]
                    else:   # odd-sized tensors leave alignment gaps in the flat buffer: pack/unpack around the call
                        r['params'] = z(n_par)
                        r['d_params'] = z(n_par)
                        r['plist'] = ps
                    if self.train:
                        max_lv_ws = max(max_lv_ws, capi.lv_bwd_ws_doubles(d))
                        r['dF'] = None if first else z(T, D_cur)
                D_cur += Lw
            else:
                kern = layer.kern
                mix = hasattr(kern, 'W')
                base = kern.kernel if mix else kern
                feat = layer.feature.feat if hasattr(layer.feature, 'feat') else layer.feature
                M, R = layer.num_inducing, layer.num_outputs
                P = kern.W.shape[0] if mix else R
                D = D_cur
                assert feat.Z.shape == (M, D), 'inducing inputs %s do not match layer input width %d' % (feat.Z.shape, D)
                assert base.input_dim <= D, 'kernel input_dim %d exceeds the layer input width %d' % (base.input_dim, D)
                mfk = layer.mean_function.kind
                if mode == 'iw' and not last and not mix:
                    # reference models.py:118-125 asks plain-kernel inner layers for full_cov=True over K, a branch whose
                    # sampler (temp_workaround.py:92-96) cannot execute as written; this plan draws independently over K.
                    import warnings
                    warnings.warn('DGP_IWVI: inner GP layer %d has a plain kernel (no SharedMixedMok): the reference would '
                                  'request a joint draw over the K importance samples (models.py:118-125, a branch that '
                                  'fails in the reference itself); this plan samples independently over K.  The joint '
                                  'draw is available at the operator level (multisample_sample_conditional(full_cov=True)).'
                                  % li, RuntimeWarning, stacklevel=3)
                sample = not last   # the final layer's sample is never consumed (models.py:133-150, :96-97)
                flags = (LIB.FLAG_SAMPLE if sample else 0) | (LIB.FLAG_SAVE if self.train else 0)
                d = capi.gp_desc(T, M, D, R, P, base.kind, mix, mfk, flags, layer.jitter)
                Mp = capi.gp_mp(M)
                # lengthscales reach the kernels as a width-D vector: the parameter itself when it already is one (ARD over all
                # D columns, or D == 1), else an expanded copy (a shared lengthscale; +inf outside the kernel's active_dims)
                direct = (base.ARD or D == 1) and getattr(base, 'active_dims', None) is None and base.input_dim == D
                adims = getattr(base, 'active_dims', None)
                if adims is None and base.input_dim != D:
                    adims = list(range(base.input_dim))
                r = dict(type='gp', layer=layer, d=d, idx=li, gi=gi, M=M, R=R, P=P, D=D, mix=mix, mf=mfk, Mp=Mp,
                         sampled=sample, base=base, feat=feat, ard=direct, adims=adims,
                         adims_t=None if adims is None else torch.as_tensor(adims, dtype=torch.int64, device=dev),
                         Lm=z(Mp, Mp), aux=z(capi.gp_aux_doubles(d)), kl=self.kls[gi:gi + 1],
                         info=self.infos[gi:gi + 1], mean=z(T, P), var=z(T, P),
                         sample=z(T, P) if sample else None, eps=z(T, R) if sample else None,
                         ls_vec=None if direct else z(D))
                if self.train:
                    r['save'] = z(capi.gp_save_doubles(d))
                    r['dLm'] = z(Mp, Mp)
                    r['dX'] = z(T, D)
                    r['dls_vec'] = None if r['ard'] else z(D)
                    max_bwd_ws = max(max_bwd_ws, capi.gp_bwd_ws_doubles(d))
                    r['bwd_ws'] = z(capi.gp_bwd_ws_doubles(d))    # per layer: the parameter reductions of layer l run on
                    #                                               its side stream while layer l-1's tile kernel runs
                    r['pbwd_ws'] = z(capi.gp_pbwd_ws_doubles(d))   # per layer: prologue adjoints overlap on side streams
                gi += 1
                D_cur = P
            self.recs.append(r)
        assert self.recs[-1]['type'] == 'gp', 'the last layer must be a GPLayer'
        assert D_cur == self.Dy, 'final layer width %d != Dy %d' % (D_cur, self.Dy)
        self.lv_recs = [r for r in self.recs if r['type'] == 'lv']
        self.Lw_total = sum(r['Lw'] for r in self.lv_recs)
        if mode != 'predict':
            scale = float(model.num_data) / float(self.B * self.world_size)
            self.ed = capi.elbo_desc(self.B, self.K, self.Dy, self.Lw_total, mode == 'iw', mode == 'iw', scale)
            self.w = z(self.B, self.K)
            self.logp = z(self.B)
            self.elbo_data = z(1)
            self.elbo_ws = z(capi.elbo_ws_doubles(self.ed))
            self.kl_cat = z(T, self.Lw_total) if len(self.lv_recs) > 1 else None
            self.dmean, self.dvar = z(T, self.Dy), z(T, self.Dy)
            self.dkl_local = z(T, self.Lw_total) if self.Lw_total else None
            self.lv_ws = z(max(max_lv_ws, 1))
        # once-per-layer prologue stages (Cholesky, KL and their adjoints) depend only on the parameters: they run on
        # side streams, concurrently with each other and with the per-point stages of other layers
        self.side = [torch.cuda.Stream(device=dev) for _ in range(self.n_gp)]
        # The latency chains (Cholesky of Kuu; Cholesky / gram adjoints; the trainer's segment update + next-step
        # factorisation) are a few CTAs each and sit on the step's critical path, while the kernels they overlap with
        # queue hundreds of CTAs: on a HIGH-PRIORITY stream their blocks are dispatched as soon as an SM frees up instead
        # of after the last wave of whatever throughput kernel was launched before them.
        self.hp = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(self.n_gp)]
        self.ev_red = [torch.cuda.Event() for _ in range(self.n_gp)]
        self.ev_start = torch.cuda.Event()
        self.ev_pro = [torch.cuda.Event() for _ in range(self.n_gp)]
        self.ev_rows = [torch.cuda.Event() for _ in range(self.n_gp)]
        self.ev_pbwd = [torch.cuda.Event() for _ in range(self.n_gp)]
        self.side_b = torch.cuda.Stream(device=dev)      # second half of the first GP layer's reductions (backward)
        self.ev_part_b = torch.cuda.Event()
        self.ev_part_a = torch.cuda.Event()
        self.ev_elbo, self.ev_loss = torch.cuda.Event(), torch.cuda.Event()
        self._loss_pending = False
        # training: called as grad_hook(tag) on the stream where the gradients named by `tag` have just been completed --
        # ('gp', gi): every parameter of GP layer gi > 0; ('gp_h', 0): inducing inputs / kernel / mixing / mean-function
        # parameters of the first GP layer; ('gp_q', 0): its q_mu / q_sqrt -- so that their all-reduce, their update and the
        # next step's factorisation overlap with the rest of the backward pass (training.Trainer); None: no hook
        self.grad_hook = None
        self.X_tiled = None
        if self.recs[0]['type'] == 'gp' or not self.recs[0].get('bcast', False):
            self.X_tiled = z(T, self.Dx)
        # Points are independent through the whole chain of GP layers, and a layer's forward kernel is persistent over
        # tiles of 64 (32) points: T / tile points is rarely a multiple of the SM count, so the last wave of every layer
        # leaves SMs idle (c3: 400 tiles on 148 SMs = 2.7 waves run as 3).  The forward pass therefore runs the layer
        # chain of the full waves on the main stream and the chain of the remaining tiles on a second stream; the two
        # never exchange data before the likelihood, and the remainder's CTAs fill the SMs the other chain leaves free.
        self.split = None
        self.split_bwd = None
        self.ev_bwd0 = torch.cuda.Event()
        self.ev_rows_b = [torch.cuda.Event() for _ in range(self.n_gp)]
        self.side_f = torch.cuda.Stream(device=dev)
        self.ev_fork, self.ev_join = torch.cuda.Event(), torch.cuda.Event()
        gps = [r for r in self.recs if r['type'] == 'gp']
        first_gp = self.recs.index(gps[0])
        if all(r['type'] == 'gp' for r in self.recs[first_gp:]) and split_waves:
            tps = {capi.gp_tile_points(r['d']) for r in gps}
            if len(tps) == 1:
                tp = tps.pop()
                nsm = torch.cuda.get_device_properties(dev).multi_processor_count
                tiles = -(-T // tp)
                full = tiles // nsm * nsm
                if full > 0 and tiles - full >= nsm // 8:
                    self.split = full * tp
            # the same for the per-point half of the backward pass (its tile width may differ from the forward's)
            if self.train:
                tpb = {capi.gp_bwd_tile_points(r['d']) for r in gps}
                if len(tpb) == 1:
                    tp = tpb.pop()
                    nsm = torch.cuda.get_device_properties(dev).multi_processor_count
                    tiles = -(-T // tp)
                    full = tiles // nsm * nsm
                    if full > 0 and tiles - full >= nsm // 8:
                        self.split_bwd = full * tp

    # ------------------------------------------------------------------------------------------
    def _cv(self, p):
        return self.flat.cview(p)

    def set_batch(self, X, Y=None):
        """X [B, Dx], Y [B, Dy]: device or host tensors / arrays (copied into the plan's buffers)."""
        X = torch.as_tensor(X, dtype=F64)
        self.X.copy_(X.reshape(self.B, self.Dx), non_blocking=True)
        if Y is not None:
            Y = torch.as_tensor(Y, dtype=F64)
            self.Y.copy_(Y.reshape(self.B, self.Dy), non_blocking=True)
            capi.batch_gather(self.X, self.Y, None, self.B, self.Dx, self.Dy, None, None, self.XY)

    def set_batch_indices(self, Xdata, Ydata, idx):
        """The minibatch as row indices (int64 device tensor [B]) into device-resident data arrays X [N, Dx], Y [N, Dy]
        (gpflow.params.Minibatch, models.py:25-26): X, Y and [X, Y] are gathered in ONE launch."""
        assert idx.dtype == torch.int64 and idx.numel() == self.B
        capi.batch_gather(Xdata, Ydata, idx, self.B, self.Dx, self.Dy, self.X, self.Y, self.XY)

    def _tile(self, A, out):
        if self.mode == 'iw':      # data-major: point = n*K + k
            out.view(self.B, self.K, -1).copy_(A[:, None, :].expand(self.B, self.K, A.shape[1]))
        else:                      # sample-major: point = s*B + n
            out.view(self.K, self.B, -1).copy_(A[None, :, :].expand(self.K, self.B, A.shape[1]))

    def draw_noise(self, eps=None, seed=0, step=0, row0=0, state=None):
        """eps: list with one entry per layer (None where no noise is consumed) of arrays shaped [*, C] in point
        order, or None to draw counter-based noise keyed by (seed, step, layer) and the GLOBAL point index.
        state: optional int64 device tensor whose element 0 counts completed optimiser steps; the step is then read on
        the device (step = state[0] + 1), which makes the call replayable inside a CUDA graph."""
        for r in self.recs:
            buf = r.get('eps')
            if buf is None:
                continue
            if eps is not None:
                e = eps[r['idx']]
                buf.copy_(torch.as_tensor(np.asarray(e), dtype=F64).reshape(buf.shape), non_blocking=True)
                continue
            first = int(row0) * self.K if self.mode == 'iw' else 0
            step_add = 0 if self.mode == 'iw' else 7919 * int(row0)
            if state is not None:
                capi.normal_fill_counter(buf, self.T, buf.shape[1], first, seed, r['idx'], step_add, state)
            else:
                capi.normal_fill(buf, self.T, buf.shape[1], first, layer_seed(seed, step + step_add, r['idx']))

    def prologue_layer(self, r, part=0):
        """Once-per-step stage of GP layer `r` on the CURRENT stream: scaled inducing inputs, Cholesky of Kuu, padded
        variational parameters, KL.  part: 0 everything, LIB.FLAG_PRO_HYP / LIB.FLAG_PRO_Q one half (include/iwvi_b200.h)."""
        base, feat, layer = r['base'], r['feat'], r['layer']
        if r['ard']:
            ls = self._cv(base.lengthscales)
        else:
            if part != LIB.FLAG_PRO_Q:
                r['ls_vec'].copy_(base.full_lengthscales(self._cv(base.lengthscales).reshape(-1), r['D']))
            ls = r['ls_vec']
        r['ls'] = ls
        d = r['d'] if not part else capi.with_flags(r['d'], r['d'].flags | part)
        capi.gp_prologue_fwd(d, self._cv(feat.Z), ls, self._cv(base.variance), self._cv(layer.q_mu),
                             self._cv(layer.q_sqrt), r['Lm'], r['aux'], r['kl'], r['info'])

    def forward(self, join=True, prologue=True):
        """join=False: the two-kernel assembly of the loss slot (data term minus the global KLs) is left running on a side
        stream and joined at the end of backward(), off the critical path between the forward and the backward pass.
        prologue=False: the once-per-layer stages (constrained values, Cholesky factors, KLs) are already in place for the
        current parameters -- training.Trainer prepares them at the end of the previous step, as soon as each layer's
        update is final, where they overlap with the rest of the backward pass instead of opening the next step."""
        flat = self.flat
        main = torch.cuda.current_stream()
        if prologue:
            flat.refresh_constrained()
            self.ev_start.record(main)
        for r in self.recs:
            if r['type'] != 'gp' or not prologue:
                continue
            side = self.hp[r['gi']]
            side.wait_event(self.ev_start)
            with torch.cuda.stream(side):
                self.prologue_layer(r)
                self.ev_pro[r['gi']].record(side)
        F = None
        if self.X_tiled is not None:
            self._tile(self.X, self.X_tiled)
            F = self.X_tiled
        for r in self.recs:
            if r['type'] == 'lv':
                layer = r['layer']
                if r['d'].prior:
                    capi.lv_fwd(r['d'], F, None, None, r['eps'], r['samples'], r['kl'], r['mu'], r['sigma'])
                else:
                    if not r['packed_view']:
                        torch.cat([self._cv(p).reshape(-1) for p in r['plist']], out=r['params'])
                    if r['bcast']:
                        Fin, enc_in = self.X, self.XY
                    else:
                        self._tile(self.XY, r['enc_in'])
                        Fin, enc_in = F, r['enc_in']
                    r['Fin'] = Fin
                    capi.lv_fwd(r['d'], Fin, enc_in, r['params'], r['eps'], r['samples'], r['kl'], r['mu'], r['sigma'])
                F = r['samples']
            else:
                layer, base = r['layer'], r['base']
                r['Fin'] = F
                if prologue:
                    main.wait_event(self.ev_pro[r['gi']])
                args = (r['d'], r['Lm'], r['aux'], F,
                        self._cv(layer.kern.W) if r['mix'] else None,
                        self._cv(layer.mean_function.A) if r['mf'] == 'Linear' else None,
                        self._cv(layer.mean_function.b) if r['mf'] == 'Linear' else None,
                        r['eps'], r['sample'], r['mean'], r['var'], r.get('save'))
                if self.split is None:
                    capi.gp_rows_fwd(*args)
                else:
                    if r['gi'] == 0:          # the chain's inputs (latent-variable layer output, noise) are complete
                        self.ev_fork.record(main)
                        self.side_f.wait_event(self.ev_fork)
                    if prologue:
                        self.side_f.wait_event(self.ev_pro[r['gi']])
                    capi.gp_rows_fwd_range(*args, 0, self.split)
                    with torch.cuda.stream(self.side_f):
                        capi.gp_rows_fwd_range(*args, self.split, self.T)
                F = r['sample']
        if self.split is not None:
            self.ev_join.record(self.side_f)
            main.wait_event(self.ev_join)
        last = self.recs[-1]
        if self.mode == 'predict':
            return last['mean'], last['var']
        if len(self.lv_recs) > 1:
            torch.cat([r['kl'] for r in self.lv_recs], 1, out=self.kl_cat)
            kl_local = self.kl_cat
        elif self.lv_recs:
            kl_local = self.lv_recs[0]['kl']
        else:
            kl_local = None
        lik = self._cv(self.model.likelihood.variance)
        capi.iwelbo_fwd(self.ed, last['mean'], last['var'], self.Y, lik, kl_local, self.elbo_data, self.logp, self.w,
                        self.elbo_ws)
        # ELBO of this rank's shard; the global KL enters with weight 1/world_size so that a SUM all-reduce is exact
        self.ev_elbo.record(main)
        self.side_b.wait_event(self.ev_elbo)
        with torch.cuda.stream(self.side_b):
            torch.sub(self.elbo_data, self.kls[:self.n_gp].sum().reshape(1), alpha=1.0 / self.world_size,
                      out=flat.loss_slot)
            self.ev_loss.record(self.side_b)
        self._loss_pending = True
        if join:
            self._join_loss()
        return flat.loss_slot

    def _fold_dls(self, r):
        """Gradient of the width-D lengthscale vector the kernels saw -> gradient of the lengthscale parameter: the sum for
        a shared lengthscale, the active columns for an ARD kernel restricted by active_dims (inactive columns carry an
        exact zero: their 1 / lengthscale is 0)."""
        base, g = r['base'], self.flat.gview(r['base'].lengthscales)
        if base.ARD:
            torch.index_select(r['dls_vec'], 0, r['adims_t'], out=g)
        else:
            torch.sum(r['dls_vec'], 0, keepdim=True, out=g)

    def _join_loss(self):
        if self._loss_pending:
            torch.cuda.current_stream().wait_event(self.ev_loss)
            self._loss_pending = False

    def backward(self):
        flat = self.flat
        last = self.recs[-1]
        lik_p = self.model.likelihood.variance
        capi.iwelbo_bwd(self.ed, last['mean'], last['var'], self.Y, self._cv(lik_p), self.w, self.one, self.dmean,
                        self.dvar, self.dkl_local, flat.gview(lik_p), self.elbo_ws)
        two = self.split_bwd is not None
        if two:                # the remainder's chain of per-point launches starts here, on its own stream
            self.ev_bwd0.record(torch.cuda.current_stream())
            self.side_f.wait_event(self.ev_bwd0)
        d_next = None          # cotangent of the current layer's output samples
        lv_off = self.Lw_total
        for r in reversed(self.recs):
            if r['type'] == 'gp':
                layer, base, feat = r['layer'], r['base'], r['feat']
                is_last = r is last
                gls = flat.gview(base.lengthscales) if r['ard'] else r['dls_vec']
                W = self._cv(layer.kern.W) if r['mix'] else None
                lin = r['mf'] == 'Linear'
                mfA = self._cv(layer.mean_function.A) if lin else None
                mfb = self._cv(layer.mean_function.b) if lin else None
                outs = (flat.gview(feat.Z), gls, flat.gview(base.variance), flat.gview(layer.q_mu),
                        flat.gview(layer.q_sqrt))
                # gradients of frozen W / mean-function parameters are not formed (set_trainable(False),
                # build_models.py:209,225-227): passing NULL skips their reductions
                gW = flat.gview(layer.kern.W) if (r['mix'] and layer.kern.W.trainable) else None
                gA = flat.gview(layer.mean_function.A) if (lin and layer.mean_function.A.trainable) else None
                gb = flat.gview(layer.mean_function.b) if (lin and layer.mean_function.b.trainable) else None
                # The per-point half (epilogue adjoint + tile kernel: dX for the layer below, Bbar, per-CTA partials) runs on
                # the main stream; the parameter half (split-K reductions over the points, fixed-order sums, then the
                # Cholesky / gram adjoint) follows on the layer's side stream, where it overlaps with the tile kernel of the
                # layer below -- each fills the SMs the other leaves idle in its last wave.
                args = (r['Lm'], r['aux'], r['save'], r['Fin'], W, mfA, mfb, r['eps'],
                        d_next if r['sampled'] else None,
                        self.dmean if is_last else None, self.dvar if is_last else None,
                        r['dX'], outs[0], outs[1], outs[2], outs[3], outs[4], r['dLm'], gW, gA, gb, r['bwd_ws'])
                fl = r['d'].flags
                main = torch.cuda.current_stream()
                gi = r['gi']
                side = self.side[gi]
                if two:
                    # Points are independent through the backward chain of GP layers too: the full waves of tiles run on
                    # the main stream, the remaining tiles (one per CTA) as a second chain whose CTAs fill the SMs the
                    # last wave of every layer leaves idle; the parameter reductions wait for both.
                    dpt = capi.with_flags(r['d'], fl | LIB.FLAG_ONLY_EPI | LIB.FLAG_ONLY_TILE | LIB.FLAG_ONLY_GRAM)
                    capi.gp_rows_bwd_range(dpt, *args, 0, self.split_bwd)
                    with torch.cuda.stream(self.side_f):
                        capi.gp_rows_bwd_range(dpt, *args, self.split_bwd, self.T)
                        self.ev_rows_b[gi].record(self.side_f)
                    fl |= LIB.FLAG_TWO_CHAINS
                    side.wait_event(self.ev_rows_b[gi])
                    if gi == 0:
                        self.side_b.wait_event(self.ev_rows_b[gi])
                else:
                    capi.gp_rows_bwd(capi.with_flags(r['d'], fl | LIB.FLAG_ONLY_EPI | LIB.FLAG_ONLY_TILE | LIB.FLAG_ONLY_GRAM), *args)
                self.ev_rows[gi].record(main)
                side.wait_event(self.ev_rows[gi])
                pargs = (r['Lm'], r['aux'], self._cv(feat.Z), r['ls'], self._cv(base.variance), self._cv(layer.q_mu),
                         self._cv(layer.q_sqrt), r['dLm'], self.dkl, outs[0], outs[1], outs[2], outs[3], outs[4],
                         r['pbwd_ws'])
                red = fl | LIB.FLAG_ONLY_REDUCE | LIB.FLAG_ONLY_FINAL | (LIB.FLAG_FAST_REDUCE if self.fast_reduce else 0)
                if gi == 0:
                    # The first GP layer comes last in the backward pass, so nothing is left to hide its Cholesky / gram
                    # adjoint chain behind -- except its own reductions: dLm first (part A), then the chain on this stream
                    # while dq_sqrt / dq_mu (part B) are formed on a second one; the KL adjoint adds into part B's outputs.
                    # Part A goes on the HIGH-PRIORITY stream: launched together with part B at equal priority its 140 CTAs
                    # interleave with part B's 700 and finish when those do, which leaves the whole adjoint chain (and
                    # the trainer's update + next-step factorisation behind it) exposed at the end of the step.
                    hp = self.hp[gi]
                    hp.wait_event(self.ev_rows[gi])
                    if two:
                        hp.wait_event(self.ev_rows_b[gi])
                    with torch.cuda.stream(hp):
                        capi.gp_rows_bwd(capi.with_flags(r['d'], red | LIB.FLAG_PART_A), *args)
                        self.ev_part_a.record(hp)
                        capi.gp_prologue_bwd(capi.with_flags(r['d'], LIB.FLAG_ACCUM | LIB.FLAG_SKIP_KL), *pargs)
                        if not r['ard']:
                            self._fold_dls(r)
                        if self.grad_hook is not None:
                            self.grad_hook(('gp_h', 0))
                        self.ev_pbwd[gi].record(hp)
                    self.side_b.wait_event(self.ev_rows[gi])
                    # Part B starts when part A's reduce + finalize launches are done (its adjoint chain still overlaps):
                    # the two reduce launches must NOT run side by side.  With both in flight (part A's 140 CTAs on the
                    # high-priority stream arriving while part B's CTAs stream their operands) part B's sums came out
                    # wrong by about one point's contribution in about one cold evaluation out of two once the reduce
                    # kernel's product loop got faster (round 2; bisected on the GPU: any ordering that keeps the two
                    # reduce grids apart is clean, any that lets them overlap is not; the memory they touch is disjoint,
                    # the mechanism is not understood).  Part A is one short wave, so nothing is lost by the order.
                    self.side_b.wait_event(self.ev_part_a)
                    with torch.cuda.stream(self.side_b):
                        capi.gp_rows_bwd(capi.with_flags(r['d'], red | LIB.FLAG_PART_B), *args)
                        capi.gp_prologue_bwd(capi.with_flags(r['d'], LIB.FLAG_ACCUM | LIB.FLAG_ONLY_KL), *pargs)
                        if self.grad_hook is not None:
                            self.grad_hook(('gp_q', 0))
                        self.ev_part_b.record(self.side_b)
                else:
                    hp = self.hp[gi]
                    with torch.cuda.stream(side):
                        capi.gp_rows_bwd(capi.with_flags(r['d'], red), *args)
                        self.ev_red[gi].record(side)
                    hp.wait_event(self.ev_red[gi])
                    with torch.cuda.stream(hp):
                        capi.gp_prologue_bwd(capi.with_flags(r['d'], LIB.FLAG_ACCUM), *pargs)
                        if not r['ard']:
                            self._fold_dls(r)
                        if self.grad_hook is not None:
                            self.grad_hook(('gp', gi))
                        self.ev_pbwd[gi].record(hp)
                d_next = r['dX']
            else:
                if two:        # (a latent-variable layer only ever sits below the GP chain when the pass is split)
                    torch.cuda.current_stream().wait_event(self.ev_rows_b[0])
                    two = False
                Lw = r['Lw']
                lv_off -= Lw
                if len(self.lv_recs) > 1:
                    d_kl = self.dkl_local[:, lv_off:lv_off + Lw].contiguous()
                else:
                    d_kl = self.dkl_local
                enc_in = self.XY if r['bcast'] else r['enc_in']
                capi.lv_bwd(r['d'], r['Fin'], enc_in, r['params'], r['eps'], r['mu'], r['sigma'], d_next, d_kl, None,
                            None, r['d_params'], r.get('dF'), self.lv_ws)
                if not r['packed_view']:
                    o = 0
                    for p in r['plist']:
                        flat.gview(p).copy_(r['d_params'][o:o + p.size].view(flat.gview(p).shape))
                        o += p.size
                d_next = r.get('dF')
        main = torch.cuda.current_stream()
        if two:                # no latent-variable layer below the GP chain: join the remainder's chain here
            main.wait_event(self.ev_rows_b[0])
        for ev in self.ev_pbwd:
            main.wait_event(ev)
        if self.n_gp:
            main.wait_event(self.ev_part_b)
        self._join_loss()
        return flat.g

    def elbo_and_grads(self, X, Y, eps=None, seed=0, step=0, row0=0):
        self.set_batch(X, Y)
        self.draw_noise(eps, seed, step, row0)
        loss = self.forward(join=False)
        self.backward()
        return loss

    def check_info(self):
        """Raises if a Cholesky failed (LAPACK-style info from the device).  Synchronises."""
        info = self.infos[:self.n_gp].cpu().numpy()
        if (info != 0).any():
            i = int(np.nonzero(info)[0][0])
            raise RuntimeError('Cholesky of Kuu failed in GP layer %d: leading minor of order %d is not positive '
                               'definite' % (i, int(info[i])))
