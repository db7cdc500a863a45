"""Operator-level boundary of the hot path: the functions the reference's layers call (reference
dgps_with_iwvi/temp_workaround.py -- same module name, same function names and argument meaning):

    multisample_sample_conditional(Xnew, feat, kern, f, full_cov=, full_output_cov=, q_sqrt=, white=)   :118-161
    independent_multisample_sample_conditional(...)                                                    :12-98
    gauss_kl(q_mu, q_sqrt, K=None)                                                                     :167-188

plus the two helpers layers.py needs (Encoder.__call__ layers.py:137-152, LatentVariableLayer.propagate :72-105).
Each is a torch.autograd.Function whose forward and backward call the C ABI (include/iwvi_b200.h) on CUDA tensors,
so `layer.propagate` composes with torch autograd exactly as the reference's composes with tf.gradients.  Whole
models train through engine.Engine instead (same C ABI, no autograd graph, preallocated buffers).

Extra keyword arguments the reference gets implicitly: `eps` (the N(0,1) draw of tf.random_normal at :89 / layers.py:86;
drawn with torch.randn when omitted), `mean_function` and `jitter` (the reference adds the mean function in
GPLayer.propagate, layers.py:46-48; here it is fused into the per-point kernel).

Not on the hot path and therefore not in CUDA (raises NotImplementedError): full_output_cov=True.  white=False (:63-65,
never taken by GPLayer, layers.py:42) is served by the same per-point kernels after an M x M change of variables
(Lm^-1 f, Lm^-1 tril(q_sqrt)), see independent_multisample_sample_conditional.  q_sqrt=None (:174-184, SGHMC) and the 2-D diagonal q_sqrt (:72-73) are mapped
onto the same kernels (zero / diagonal Cholesky factors).  full_cov=True
(:55-57,82-83, whose joint sampler :92-96 has a shape bug and is dead code in training) is served
by iwvi_gp_fullcov_fwd / _bwd (csrc/gp_fullcov.cu: DMMA Gram products on the saved A / U panels, shared-memory
Cholesky and its adjoint, corrected joint draw; batched blocked Cholesky over global memory beyond 64 points), differentiable for groups of up to
256 points (BASELINE config c4: K = 256)."""
import numpy as np
import torch

from . import _lib as LIB
from . import capi
from .params import Parameter

F64 = torch.float64


def _t(v, dev=None):
    """Constrained value of a Parameter, or a tensor / array, as a float64 tensor on the CUDA device."""
    if v is None:
        return None
    if isinstance(v, Parameter):
        v = v.value
    v = torch.as_tensor(v, dtype=F64) if not torch.is_tensor(v) else v
    if not v.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError('dgps_with_iwvi_b200: no CUDA device -- the conditional has no CPU fallback')
        v = v.cuda()
    return v


def _c(t):
    return None if t is None else t.detach().contiguous()


class _GPConditional(torch.autograd.Function):
    """iwvi_gp_prologue_fwd + iwvi_gp_rows_fwd; backward = iwvi_gp_rows_bwd + iwvi_gp_prologue_bwd."""

    @staticmethod
    def forward(ctx, X, Z, ls, variance, q_mu, q_sqrt, W, mfA, mfb, eps, meta):
        LIB.load()
        dev = X.device
        T, D = X.shape
        M, R = q_mu.shape
        mix = W is not None
        P = W.shape[0] if mix else R
        sample = eps is not None
        flags = (LIB.FLAG_SAMPLE if sample else 0) | LIB.FLAG_SAVE
        d = capi.gp_desc(T, M, D, R, P, meta['kern'], mix, meta['mf'], flags, meta['jitter'])
        z = lambda *s: torch.zeros(*s, dtype=F64, device=dev)
        Mp = capi.gp_mp(M)
        Lm, aux, kl = z(Mp, Mp), z(capi.gp_aux_doubles(d)), z(1)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        Xc, Zc, lsc, vc, qmc, qsc = _c(X), _c(Z), _c(ls), _c(variance).reshape(1), _c(q_mu), _c(q_sqrt)
        Wc, Ac, bc, ec = _c(W), _c(mfA), _c(mfb), _c(eps)
        capi.gp_prologue_fwd(d, Zc, lsc, vc, qmc, qsc, Lm, aux, kl, info)
        mean, var = z(T, P), z(T, P)
        smp = z(T, P) if sample else None
        save = z(capi.gp_save_doubles(d))
        capi.gp_rows_fwd(d, Lm, aux, Xc, Wc, Ac, bc, ec, smp, mean, var, save)
        if meta.get('check', True):
            i = int(info.item())
            if i:
                raise RuntimeError('Cholesky of Kuu failed: leading minor of order %d is not positive definite' % i)
        ctx.d, ctx.meta = d, meta
        ctx.saved = (Xc, Zc, lsc, vc, qmc, qsc, Wc, Ac, bc, ec, Lm, aux, save)
        ctx.var_shape = variance.shape
        if smp is None:
            smp = mean.new_zeros(0)
        ctx.mark_non_differentiable(kl)
        return smp, mean, var, kl

    @staticmethod
    def backward(ctx, d_sample, d_mean, d_var, _d_kl):
        Xc, Zc, lsc, vc, qmc, qsc, Wc, Ac, bc, ec, Lm, aux, save = ctx.saved
        d = ctx.d
        dev = Xc.device
        z = lambda *s: torch.zeros(*s, dtype=F64, device=dev)
        T, D = Xc.shape
        M, R = qmc.shape
        Mp = Lm.shape[0]
        sampled = ec is not None
        dX, dZ, dls, dv, dqm, dqs, dLm = z(T, D), z(M, D), z(D), z(1), z(M, R), z(R, M, M), z(Mp, Mp)
        dW = z(*Wc.shape) if Wc is not None else None
        dA = z(*Ac.shape) if Ac is not None else None
        db = z(*bc.shape) if bc is not None else None
        ws = z(capi.gp_bwd_ws_doubles(d))
        capi.gp_rows_bwd(d, Lm, aux, save, Xc, Wc, Ac, bc, ec,
                         _c(d_sample) if (sampled and d_sample is not None and d_sample.numel()) else None,
                         _c(d_mean), _c(d_var), dX, dZ, dls, dv, dqm, dqs, dLm, dW, dA, db, ws)
        # (the KL is a separate operator, gauss_kl: its adjoint is skipped here)
        capi.gp_prologue_bwd(capi.with_flags(d, d.flags | LIB.FLAG_ACCUM | LIB.FLAG_SKIP_KL), Lm, aux, Zc, lsc, vc, qmc, qsc, dLm, z(1),
                             dZ, dls, dv, dqm, dqs, z(capi.gp_pbwd_ws_doubles(d)))
        return dX, dZ, dls, dv.reshape(ctx.var_shape), dqm, dqs, dW, dA, db, None, None


class _GPFullCov(torch.autograd.Function):
    """iwvi_gp_prologue_fwd + iwvi_gp_rows_fwd + iwvi_gp_fullcov_fwd: mean [T, R], covariance over the inner axis
    [S, R, N, N] and (with noise z [S, R, N]) the joint draw [T, R].  backward = iwvi_gp_fullcov_bwd, which recasts the
    N x N cotangents as panels the per-point backward kernels understand, then iwvi_gp_rows_bwd + iwvi_gp_prologue_bwd."""

    @staticmethod
    def forward(ctx, X, Z, ls, variance, q_mu, q_sqrt, mfA, mfb, z, meta):
        LIB.load()
        dev = X.device
        T, D = X.shape
        M, R = q_mu.shape
        S_, N = meta['S'], meta['N']
        d = capi.gp_desc(T, M, D, R, R, meta['kern'], False, meta['mf'], LIB.FLAG_SAVE, meta['jitter'])
        zr = lambda *s: torch.zeros(*s, dtype=F64, device=dev)
        Mp = capi.gp_mp(M)
        Lm, aux, kl = zr(Mp, Mp), zr(capi.gp_aux_doubles(d)), zr(1)
        info = torch.zeros(2, dtype=torch.int32, device=dev)
        Xc, Zc, lsc, vc, qmc, qsc = _c(X), _c(Z), _c(ls), _c(variance).reshape(1), _c(q_mu), _c(q_sqrt)
        Ac, bc, zc = _c(mfA), _c(mfb), _c(z)
        capi.gp_prologue_fwd(d, Zc, lsc, vc, qmc, qsc, Lm, aux, kl, info[:1])
        mean, var = zr(T, R), zr(T, R)
        save = zr(capi.gp_save_doubles(d))
        capi.gp_rows_fwd(d, Lm, aux, Xc, None, Ac, bc, None, None, mean, var, save)
        cov = zr(S_, R, N, N)
        smp = zr(T, R) if zc is not None else None
        n_ws = capi.gp_fullcov_ws_doubles(d, S_, N)
        ws = zr(n_ws) if n_ws else None
        capi.gp_fullcov_fwd(d, S_, N, aux, Xc, save, mean, zc, meta['chol_jitter'], cov, smp, info[1:], ws)
        i0, i1 = (int(v) for v in info.tolist())
        if i0:
            raise RuntimeError('Cholesky of Kuu failed: leading minor of order %d is not positive definite' % i0)
        if i1:
            raise RuntimeError('Cholesky of the covariance over the inner axis failed: leading minor of order %d is not '
                               'positive definite' % i1)
        ctx.d, ctx.meta = d, meta
        ctx.saved = (Xc, Zc, lsc, vc, qmc, qsc, Ac, bc, zc, Lm, aux, save)
        ctx.var_shape = variance.shape
        return (smp if smp is not None else mean.new_zeros(0)), mean, cov

    @staticmethod
    def backward(ctx, d_sample, d_mean, d_cov):
        Xc, Zc, lsc, vc, qmc, qsc, Ac, bc, zc, Lm, aux, save = ctx.saved
        d, meta = ctx.d, ctx.meta
        S_, N = meta['S'], meta['N']
        if N > 256:
            raise NotImplementedError('the adjoint of the covariance over the inner axis is built for at most 256 points '
                                      'per group (N = %d)' % N)
        dev = Xc.device
        zr = lambda *s: torch.zeros(*s, dtype=F64, device=dev)
        T, D = Xc.shape
        M, R = qmc.shape
        Mp = Lm.shape[0]
        ds = _c(d_sample) if (zc is not None and d_sample is not None and d_sample.numel()) else None
        save2, dXk, part = torch.zeros_like(save), zr(T, D), zr(S_, 40)
        n_ws = capi.gp_fullcov_ws_doubles(d, S_, N)
        capi.gp_fullcov_bwd(d, S_, N, aux, Xc, save, zc, meta['chol_jitter'], ds, _c(d_cov), save2, dXk, part,
                            zr(n_ws) if n_ws else None)
        gm = zr(T, R) if d_mean is None else d_mean.detach().clone()
        if ds is not None:
            gm += ds
        ones = torch.ones(T, R, dtype=F64, device=dev)
        dX, dZ, dls, dv, dqm, dqs, dLm = zr(T, D), zr(M, D), zr(D), zr(1), zr(M, R), zr(R, M, M), zr(Mp, Mp)
        dA = zr(*Ac.shape) if Ac is not None else None
        db = zr(*bc.shape) if bc is not None else None
        ws = zr(capi.gp_bwd_ws_doubles(d))
        fl = d.flags | LIB.FLAG_NO_KDIAG
        args = (Lm, aux, save2, Xc, None, Ac, bc, None, None, gm, ones, dX, dZ, dls, dv, dqm, dqs, dLm, None, dA, db, ws)
        capi.gp_rows_bwd(capi.with_flags(d, fl | LIB.FLAG_ONLY_EPI | LIB.FLAG_ONLY_TILE | LIB.FLAG_ONLY_GRAM), *args)
        # the A panel is the first slot of the save layout (csrc/common.cuh SaveLayout): [Tp/64][Mp/64][64][68]
        n_a = ((T + 127) // 128 * 2) * (Mp // 64) * 64 * 68
        save2[:n_a].copy_(save[:n_a])
        capi.gp_rows_bwd(capi.with_flags(d, fl | LIB.FLAG_ONLY_REDUCE | LIB.FLAG_ONLY_FINAL), *args)
        dX += dXk
        dls += part[:, :D].sum(0)
        dv += part[:, 32].sum()
        capi.gp_prologue_bwd(capi.with_flags(d, d.flags | LIB.FLAG_ACCUM | LIB.FLAG_SKIP_KL), Lm, aux, Zc, lsc, vc, qmc, qsc,
                             dLm, zr(1), dZ, dls, dv, dqm, dqs, zr(capi.gp_pbwd_ws_doubles(d)))
        return dX, dZ, dls, dv.reshape(ctx.var_shape), dqm, dqs, dA, db, None, None


class _KuuCholesky(torch.autograd.Function):
    """Lm = chol(Kuu + jitter I) [M, M] as a differentiable function of (Z, lengthscales, variance):
    iwvi_gp_prologue_fwd (temp_workaround.py:39,48) forward, iwvi_gp_prologue_bwd (Cholesky + gram adjoint) backward."""

    @staticmethod
    def forward(ctx, Z, ls, variance, meta):
        LIB.load()
        dev = Z.device
        M, D = Z.shape
        d = capi.gp_desc(0, M, D, 1, 1, meta['kern'], False, 'Zero', 0, meta['jitter'])
        z = lambda *s: torch.zeros(*s, dtype=F64, device=dev)
        Mp = capi.gp_mp(M)
        Lm, aux, kl = z(Mp, Mp), z(capi.gp_aux_doubles(d)), z(1)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        Zc, lsc, vc = _c(Z), _c(ls), _c(variance).reshape(1)
        q_mu, q_sqrt = z(M, 1), torch.eye(M, dtype=F64, device=dev)[None].contiguous()
        capi.gp_prologue_fwd(d, Zc, lsc, vc, q_mu, q_sqrt, Lm, aux, kl, info)
        i = int(info.item())
        if i:
            raise RuntimeError('Cholesky of Kuu failed: leading minor of order %d is not positive definite' % i)
        ctx.d, ctx.saved, ctx.var_shape = d, (Zc, lsc, vc, q_mu, q_sqrt, Lm, aux), variance.shape
        return Lm[:M, :M].contiguous()

    @staticmethod
    def backward(ctx, dL):
        Zc, lsc, vc, q_mu, q_sqrt, Lm, aux = ctx.saved
        d = ctx.d
        dev = Zc.device
        z = lambda *s: torch.zeros(*s, dtype=F64, device=dev)
        M, D = Zc.shape
        Mp = Lm.shape[0]
        dLm = z(Mp, Mp)
        dLm[:M, :M] = torch.tril(dL)
        dZ, dls, dv = z(M, D), z(D), z(1)
        capi.gp_prologue_bwd(capi.with_flags(d, LIB.FLAG_SKIP_KL), Lm, aux, Zc, lsc, vc, q_mu, q_sqrt, dLm, z(1),
                             dZ, dls, dv, z(M, 1), z(1, M, M), z(capi.gp_pbwd_ws_doubles(d)))
        return dZ, dls, dv.reshape(ctx.var_shape), None


def _kern_parts(kern):
    mix = hasattr(kern, 'W')
    base = kern.kernel if mix else kern
    return mix, base


def independent_multisample_sample_conditional(Xnew, feat, kern, f, *, full_cov=False, q_sqrt=None, white=False,
                                               eps=None, mean_function=None, jitter=1e-6, W=None, sample=True):
    """Reference temp_workaround.py:12-98.  Xnew [S, N, D] (or [T, D]); f = q_mu [M, R]; q_sqrt [R, M, M].
    Returns sample, mean [S, N, P], var [S, N, P] (full_cov=False) or [S, R, N, N] (full_cov=True, forward only)."""
    X = _t(Xnew)
    lead = X.shape[:-1]
    D = X.shape[-1]
    X2 = X.reshape(-1, D)
    T = X2.shape[0]
    Z = _t(feat.feat.Z if hasattr(feat, 'feat') else feat.Z)
    q_mu = _t(f)
    M, R = q_mu.shape
    if q_sqrt is None:
        # temp_workaround.py:71 skipped (the SGHMC case, q(u) a point mass): tril(0)^T A = 0 adds nothing to the variance
        q_sq = torch.zeros(R, M, M, dtype=F64, device=q_mu.device)
    else:
        q_sq = _t(q_sqrt)
        if q_sq.dim() == 2:
            # diagonal form [M, R] (temp_workaround.py:72-73): LTA = A * q_sqrt^T, i.e. a diagonal Cholesky factor per output
            q_sq = torch.diag_embed(q_sq.t())
        elif q_sq.dim() != 3:
            raise ValueError('Bad dimension for q_sqrt: %s' % str(q_sq.dim()))
    ls, variance = _t(kern.lengthscales), _t(kern.variance)
    ls_vec = kern.full_lengthscales(ls, D)      # width D; +inf (1 / lengthscale = 0) outside the kernel's active_dims
    if not white:
        # temp_workaround.py:63-65: A <- Lm^-T A before the mean (:68) and the q_sqrt projection (:78), the prior
        # variance term (:59) keeping the first A.  With Lm lower-triangular,
        #     (Lm^-T A)^T f = A^T (Lm^-1 f)   and   tril(q_sqrt)^T Lm^-T A = (Lm^-1 tril(q_sqrt))^T A,
        # and Lm^-1 tril(q_sqrt) is itself lower-triangular: the unwhitened conditional IS the whitened one at
        # (Lm^-1 f, Lm^-1 tril(q_sqrt)).  That change of variables is M x M work once per layer (two triangular solves,
        # library TRSM), differentiable through Lm (its adjoint reaches Z and the kernel parameters through
        # iwvi_gp_prologue_bwd); the per-point kernels run unchanged.
        Lm = _KuuCholesky.apply(Z, ls_vec, variance, dict(kern=kern.kind, jitter=float(jitter)))
        q_mu = torch.linalg.solve_triangular(Lm, q_mu, upper=False)
        if q_sqrt is not None:
            q_sq = torch.linalg.solve_triangular(Lm, torch.tril(q_sq), upper=False)
    Wt = _t(W)
    mf_kind = mean_function.kind if mean_function is not None else 'Zero'
    mfA = _t(mean_function.A) if mf_kind == 'Linear' else None
    mfb = _t(mean_function.b) if mf_kind == 'Linear' else None
    P = Wt.shape[0] if Wt is not None else R
    meta = dict(kern=kern.kind, mf=mf_kind, jitter=float(jitter))
    if not full_cov:
        e = None
        if sample:
            e = torch.randn(T, R, dtype=F64, device=X.device) if eps is None else _t(eps).reshape(T, R)
        smp, mean, var, _ = _GPConditional.apply(X2, Z, ls_vec, variance, q_mu, q_sq, Wt, mfA, mfb, e, meta)
        return (smp.reshape(*lead, P) if e is not None else None), mean.reshape(*lead, P), var.reshape(*lead, P)
    # ---- full covariance over the inner axis and the joint draw: iwvi_gp_fullcov_fwd / _bwd on the saved panels
    if Wt is not None:
        raise NotImplementedError('the Mok branch forces full_cov=False (temp_workaround.py:125-129)')
    S_ = int(np.prod(lead[:-1])) if len(lead) > 1 else 1
    N = lead[-1]
    if sample and N > 256:
        raise NotImplementedError('the joint draw is built for inner axes of at most 256 points (N = %d); the covariance '
                                  'itself (sample=False) has no such limit' % N)
    z = None
    if sample:
        # the joint draw the reference intends at :92-96, in its [S, R, N, 1] noise order (:93-94)
        z = torch.randn(S_, R, N, dtype=F64, device=X.device) if eps is None else _t(eps).reshape(S_, R, N)
    # 3-D inputs: tf.cholesky(fvar) as is (:95); 2-D inputs go through gpflow's _sample_mvn, which adds the jitter
    meta.update(S=S_, N=N, chol_jitter=0.0 if len(lead) > 1 else float(jitter))
    smp, mean, cov = _GPFullCov.apply(X2, Z, ls_vec, variance, q_mu, q_sq, mfA, mfb, z, meta)
    if len(lead) == 1:
        cov = cov[0]
    return (smp.reshape(*lead, R) if sample else None), mean.reshape(*lead, R), cov


def multisample_sample_conditional(Xnew, feat, kern, f, *, full_cov=False, full_output_cov=False, q_sqrt=None,
                                   white=False, eps=None, mean_function=None, jitter=1e-6, sample=True):
    """Reference temp_workaround.py:118-161.  SharedMixedMok + MixedKernelSharedMof: the latent GPs are evaluated with
    full_cov forced to False (:125-129) and mixed by W (mean, sample: @ W^T; variance: @ (W^2)^T, :142-145) -- fused in
    the per-point kernel.  2-D inputs (the gpflow sample_conditional branch, :133-140,156-161) run the same kernel."""
    if full_output_cov:
        raise NotImplementedError('full_output_cov=True is not used on the IW-ELBO path')
    mix, base = _kern_parts(kern)
    if mix:
        return independent_multisample_sample_conditional(
            Xnew, feat, base, f, full_cov=False, q_sqrt=q_sqrt, white=white, eps=eps, mean_function=mean_function,
            jitter=jitter, W=kern.W, sample=sample)
    return independent_multisample_sample_conditional(
        Xnew, feat, base, f, full_cov=full_cov, q_sqrt=q_sqrt, white=white, eps=eps, mean_function=mean_function,
        jitter=jitter, sample=sample)


class _GaussKL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q_mu, q_sqrt):
        LIB.load()
        qm, qs = _c(q_mu), _c(q_sqrt)
        M, R = qm.shape
        kl = torch.zeros(1, dtype=F64, device=qm.device)
        capi.gauss_kl_fwd(M, R, qm, qs, kl)
        ctx.saved = (qm, qs)
        return kl.reshape(())

    @staticmethod
    def backward(ctx, dkl):
        qm, qs = ctx.saved
        M, R = qm.shape
        dqm, dqs = torch.zeros_like(qm), torch.zeros_like(qs)
        capi.gauss_kl_bwd(M, R, qm, qs, dkl.detach().reshape(1).contiguous(), dqm, dqs)
        return dqm, dqs


def gauss_kl(q_mu, q_sqrt, K=None, jitter=1e-6):
    """Reference temp_workaround.py:167-188.  K=None (what GPLayer passes, layers.py:44): gpflow's whitened gauss_kl on
    the CUDA kernel.  K given (the unwhitened prior N(0, K), never used by the layers): with L = chol(K),
    KL[N(q_mu, Lq Lq^T) || N(0, K)] equals the whitened KL at (L^-1 q_mu, L^-1 tril(q_sqrt)) -- an M x M change of variables
    (library Cholesky + TRSM, once per call), then the same kernel.  q_sqrt=None (:174-184, SGHMC): minus the log density
    of q_mu under N(0, K + jitter I) (or N(0, I)), summed over outputs."""
    qm = _t(q_mu)
    if q_sqrt is None:
        M = qm.shape[0]
        if K is None:
            alpha, logdet = qm, 0.0
        else:
            Kt = _t(K)
            L = torch.linalg.cholesky(Kt + jitter * torch.eye(M, dtype=F64, device=Kt.device))     # :182
            alpha = torch.linalg.solve_triangular(L, qm, upper=False)
            logdet = torch.log(torch.diagonal(L)).sum()
        # -sum_r log N(q_mu_r; 0, L L^T)   (gpflow.logdensities.multivariate_normal, :184)
        return 0.5 * (alpha * alpha).sum() + qm.shape[1] * (0.5 * M * float(np.log(2.0 * np.pi)) + logdet)
    q_sq = _t(q_sqrt)
    if q_sq.dim() == 2:
        q_sq = torch.diag_embed(q_sq.t())      # gpflow gauss_kl's diagonal form [M, R]
    if K is not None:
        L = torch.linalg.cholesky(_t(K))
        qm = torch.linalg.solve_triangular(L, qm, upper=False)
        q_sq = torch.linalg.solve_triangular(L, torch.tril(q_sq), upper=False)
    return _GaussKL.apply(qm, q_sq)


# ------------------------------------------------------------------------------------------------------------------
# LatentVariableLayer / Encoder
# ------------------------------------------------------------------------------------------------------------------
class _LVPropagate(torch.autograd.Function):
    """iwvi_lv_fwd / iwvi_lv_bwd.  F [T, Df], enc_in [T, Dxy] or None (prior branch), params packed W0,b0,W1,b1,..."""

    @staticmethod
    def forward(ctx, F, enc_in, params, eps, meta):
        LIB.load()
        dev = eps.device
        T, Lw = eps.shape
        Df = 0 if F is None else F.shape[1]
        prior = enc_in is None
        d = capi.lv_desc(T, 1, Df, 0 if prior else enc_in.shape[1], Lw, None if prior else meta['dims'],
                         sampled=meta['sampled'], f_bcast=False, prior=prior, prior_mu=meta.get('prior_mu', 0.0),
                         prior_sigma=meta.get('prior_sigma', 1.0), act=meta.get('act', 'tanh'))
        z = lambda *s: torch.zeros(*s, dtype=F64, device=dev)
        samples, kl, mu, sigma = z(T, Df + Lw), z(T, Lw), z(T, Lw), z(T, Lw)
        Fc, Ec, Pc, ec = _c(F), _c(enc_in), _c(params), _c(eps)
        capi.lv_fwd(d, Fc, Ec, Pc, ec, samples, kl, mu, sigma)
        ctx.d = d
        ctx.saved = (Fc, Ec, Pc, ec, mu, sigma)
        return samples, kl, mu, sigma

    @staticmethod
    def backward(ctx, d_samples, d_kl, d_mu, d_sigma):
        Fc, Ec, Pc, ec, mu, sigma = ctx.saved
        d = ctx.d
        dev = ec.device
        z = lambda *s: torch.zeros(*s, dtype=F64, device=dev)
        dF = z(*Fc.shape) if Fc is not None else None
        if d.prior:
            capi.lv_bwd(d, Fc, None, None, ec, None, None, _c(d_samples), _c(d_kl), None, None, None, dF, None)
            return dF, None, None, None, None
        dP = z(Pc.numel())
        ws = z(capi.lv_bwd_ws_doubles(d))
        capi.lv_bwd(d, Fc, Ec, Pc, ec, mu, sigma, _c(d_samples), _c(d_kl), _c(d_mu), _c(d_sigma), dP, dF, ws)
        return dF, None, dP, None, None


def _packed_encoder_params(enc):
    return torch.cat([_t(p).reshape(-1) for pair in zip(enc.Ws, enc.bs) for p in pair])


def encoder_forward(enc, Z):
    """Encoder.__call__ (reference layers.py:137-152): Z [..., input_dim] -> (q_mu, q_sqrt) [..., latent_dim]."""
    Zt = _t(Z)
    lead = Zt.shape[:-1]
    Z2 = Zt.reshape(-1, Zt.shape[-1])
    eps0 = torch.zeros(Z2.shape[0], enc.latent_dim, dtype=F64, device=Z2.device)
    meta = dict(dims=enc.layer_dims, sampled=False, act=getattr(enc, 'activation_func', 'tanh'))
    _, _, mu, sigma = _LVPropagate.apply(None, Z2, _packed_encoder_params(enc), eps0, meta)
    return mu.reshape(*lead, enc.latent_dim), sigma.reshape(*lead, enc.latent_dim)


def latent_variable_propagate(layer, F, inference_amorization_inputs=None, is_sampled_local_regularizer=False, eps=None):
    """LatentVariableLayer.propagate (reference layers.py:72-105) -> (samples, mean, cov, kl):
    samples = [F, W], mean = [F, q_mu], cov = [0, q_sqrt^2], kl [..., latent_dim] = log q(W) - log p(W) per sample
    (is_sampled_local_regularizer) or the closed-form KL[q||N(0,1)]."""
    Ft = _t(F)
    lead = Ft.shape[:-1]
    Df, Lw = Ft.shape[-1], layer.latent_dim
    F2 = Ft.reshape(-1, Df)
    T = F2.shape[0]
    e = torch.randn(T, Lw, dtype=F64, device=F2.device) if eps is None else _t(eps).reshape(T, Lw)
    meta = dict(sampled=bool(is_sampled_local_regularizer), prior_mu=layer.prior_mu, prior_sigma=layer.prior_sigma)
    if inference_amorization_inputs is None:
        samples, kl, mu, sigma = _LVPropagate.apply(F2, None, None, e, meta)
    else:
        XY = _t(inference_amorization_inputs)
        meta['dims'] = layer.encoder.layer_dims
        meta['act'] = getattr(layer.encoder, 'activation_func', 'tanh')
        samples, kl, mu, sigma = _LVPropagate.apply(F2, XY.reshape(T, XY.shape[-1]),
                                                    _packed_encoder_params(layer.encoder), e, meta)
    mean = torch.cat([F2, mu], 1)
    cov = torch.cat([torch.zeros_like(F2), sigma * sigma], 1)
    r = lambda a, c: a.reshape(*lead, c)
    return r(samples, Df + Lw), r(mean, Df + Lw), r(cov, Df + Lw), r(kl, Lw)


def full_cov_conditional(layer, X):
    """Single-GPLayer _build_predict(X, full_cov=True) (reference models.py:89-91 via GPModel.predict_f_full_cov, pinned
    by tests/test_gp_layer.py:53-54): mean [N, R], cov [R, N, N]."""
    _, mean, cov = multisample_sample_conditional(
        X, layer.feature, layer.kern, layer.q_mu, full_cov=True, q_sqrt=layer.q_sqrt, white=True,
        mean_function=layer.mean_function, jitter=layer.jitter, sample=False)
    return mean, cov
