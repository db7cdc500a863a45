"""Tensor-level wrappers over the C ABI: every function takes contiguous float64 CUDA tensors (or None), passes
their device addresses and torch's current CUDA stream, and raises on a non-zero return code.  No arithmetic
happens here and nothing is allocated except where a docstring says so."""
import ctypes as C

import torch

from . import _lib as L

F64 = torch.float64

# kernels launched so far through this module (each entry point's launch count is fixed by csrc/*.cu)
LAUNCHES = 0


def _count(n):
    global LAUNCHES
    LAUNCHES += n


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('dgps_with_iwvi_b200 has no CPU path: tensor is on %s' % t.device)
    if t.dtype not in (torch.float64, torch.int32, torch.int64):
        raise TypeError('expected float64/int32/int64, got %s' % t.dtype)
    if not t.is_contiguous():
        raise ValueError('tensor must be contiguous')
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def gp_desc(T, M, D, R, P, kern, mix, mf, flags=0, jitter=1e-6):
    return L.GpDesc(int(T), int(M), int(D), int(R), int(P), L.KERN_IDS[kern] if isinstance(kern, str) else int(kern),
                    int(bool(mix)), L.MF_IDS[mf] if isinstance(mf, str) else int(mf), int(flags), 0, float(jitter))


def with_flags(d, flags, T=None):
    return L.GpDesc(d.T if T is None else int(T), d.M, d.D, d.R, d.P, d.kern, d.mix, d.mf, int(flags), 0, d.jitter)


def lv_desc(Be, Kt, Df, Dxy, Lw, dims, sampled, f_bcast, prior=False, prior_mu=0.0, prior_sigma=1.0, act='tanh'):
    arr = (C.c_int32 * (L.MAX_ENC_LAYERS + 1))()
    dims = list(dims or [])
    for i, v in enumerate(dims):
        arr[i] = int(v)
    return L.LvDesc(int(Be), int(Kt), int(Df), int(Dxy), int(Lw), max(len(dims) - 1, 0), arr, int(bool(sampled)),
                    int(bool(f_bcast)), int(bool(prior)), L.ACT_IDS[act] if isinstance(act, str) else int(act), 0,
                    float(prior_mu), float(prior_sigma))


def elbo_desc(B, K, Dy, Lw, iw, data_major, scale):
    return L.ElboDesc(int(B), int(K), int(Dy), int(Lw), int(bool(iw)), int(bool(data_major)), float(scale))


def _size(fn, d):
    n = fn(C.byref(d))
    if n < 0:
        raise RuntimeError('bad descriptor for %s' % fn.__name__)
    return int(n)


def gp_mp(M):
    return int(L.load().iwvi_gp_mp(int(M)))


def gp_aux_doubles(d):
    return _size(L.load().iwvi_gp_aux_doubles, d)


def gp_save_doubles(d):
    return _size(L.load().iwvi_gp_save_doubles, d)


def gp_bwd_ws_doubles(d):
    return _size(L.load().iwvi_gp_bwd_ws_doubles, d)


def gp_pbwd_ws_doubles(d):
    return _size(L.load().iwvi_gp_pbwd_ws_doubles, d)


def lv_param_doubles(d):
    return _size(L.load().iwvi_lv_param_doubles, d)


def lv_bwd_ws_doubles(d):
    return _size(L.load().iwvi_lv_bwd_ws_doubles, d)


def elbo_ws_doubles(d):
    return _size(L.load().iwvi_elbo_ws_doubles, d)


def gp_prologue_fwd(d, Z, ls, variance, q_mu, q_sqrt, Lm, aux, kl, info):
    _count(2)
    L.check(L.load().iwvi_gp_prologue_fwd(C.byref(d), _ptr(Z), _ptr(ls), _ptr(variance), _ptr(q_mu), _ptr(q_sqrt),
                                          _ptr(Lm), _ptr(aux), _ptr(kl), _ptr(info), _stream()), 'iwvi_gp_prologue_fwd')


def gp_rows_fwd(d, Lm, aux, X, W, mfA, mfb, eps, sample, mean, var, save):
    _count(2 if d.flags & L.FLAG_SAVE else 1)      # row kernel (+ the per-point epilogue kernel of a saving call)
    L.check(L.load().iwvi_gp_rows_fwd(C.byref(d), _ptr(Lm), _ptr(aux), _ptr(X), _ptr(W), _ptr(mfA), _ptr(mfb),
                                      _ptr(eps), _ptr(sample), _ptr(mean), _ptr(var), _ptr(save), _stream()),
            'iwvi_gp_rows_fwd')


def gp_rows_fwd_range(d, Lm, aux, X, W, mfA, mfb, eps, sample, mean, var, save, point_begin, point_end):
    _count(2 if d.flags & L.FLAG_SAVE else 1)
    L.check(L.load().iwvi_gp_rows_fwd_range(C.byref(d), _ptr(Lm), _ptr(aux), _ptr(X), _ptr(W), _ptr(mfA), _ptr(mfb),
                                            _ptr(eps), _ptr(sample), _ptr(mean), _ptr(var), _ptr(save),
                                            int(point_begin), int(point_end), _stream()), 'iwvi_gp_rows_fwd_range')


def gp_tile_points(d):
    tp = L.load().iwvi_gp_tile_points(C.byref(d))
    if tp < 0:
        L.check(tp, 'iwvi_gp_tile_points')
    return tp


def gp_rows_bwd(d, Lm, aux, save, X, W, mfA, mfb, eps, d_sample, d_mean, d_var, dX, dZ, dls, dvariance, dq_mu, dq_sqrt,
                dLm, dW, dmfA, dmfb, ws):
    only = d.flags & 248
    _count(bin(only).count('1') if only else 5)
    L.check(L.load().iwvi_gp_rows_bwd(C.byref(d), _ptr(Lm), _ptr(aux), _ptr(save), _ptr(X), _ptr(W), _ptr(mfA),
                                      _ptr(mfb), _ptr(eps), _ptr(d_sample), _ptr(d_mean), _ptr(d_var), _ptr(dX),
                                      _ptr(dZ), _ptr(dls), _ptr(dvariance), _ptr(dq_mu), _ptr(dq_sqrt), _ptr(dLm),
                                      _ptr(dW), _ptr(dmfA), _ptr(dmfb), _ptr(ws), _stream()), 'iwvi_gp_rows_bwd')


def gp_rows_bwd_range(d, Lm, aux, save, X, W, mfA, mfb, eps, d_sample, d_mean, d_var, dX, dZ, dls, dvariance, dq_mu,
                      dq_sqrt, dLm, dW, dmfA, dmfb, ws, point_begin, point_end):
    only = d.flags & 248
    _count(bin(only).count('1'))
    L.check(L.load().iwvi_gp_rows_bwd_range(C.byref(d), _ptr(Lm), _ptr(aux), _ptr(save), _ptr(X), _ptr(W), _ptr(mfA),
                                            _ptr(mfb), _ptr(eps), _ptr(d_sample), _ptr(d_mean), _ptr(d_var), _ptr(dX),
                                            _ptr(dZ), _ptr(dls), _ptr(dvariance), _ptr(dq_mu), _ptr(dq_sqrt), _ptr(dLm),
                                            _ptr(dW), _ptr(dmfA), _ptr(dmfb), _ptr(ws), int(point_begin), int(point_end),
                                            _stream()), 'iwvi_gp_rows_bwd_range')


def gp_bwd_tile_points(d):
    tp = L.load().iwvi_gp_bwd_tile_points(C.byref(d))
    if tp < 0:
        L.check(tp, 'iwvi_gp_bwd_tile_points')
    return tp


def gp_prologue_bwd(d, Lm, aux, Z, ls, variance, q_mu, q_sqrt, dLm, dkl, dZ, dls, dvariance, dq_mu, dq_sqrt, ws):
    _count(1 if d.flags & L.FLAG_ONLY_KL else (5 if d.flags & L.FLAG_SKIP_KL else 6))
    L.check(L.load().iwvi_gp_prologue_bwd(C.byref(d), _ptr(Lm), _ptr(aux), _ptr(Z), _ptr(ls), _ptr(variance),
                                          _ptr(q_mu), _ptr(q_sqrt), _ptr(dLm), _ptr(dkl), _ptr(dZ), _ptr(dls),
                                          _ptr(dvariance), _ptr(dq_mu), _ptr(dq_sqrt), _ptr(ws), _stream()),
            'iwvi_gp_prologue_bwd')


def gp_fullcov_ws_doubles(d, S, N):
    n = L.load().iwvi_gp_fullcov_ws_doubles(C.byref(d), int(S), int(N))
    if n < 0:
        raise RuntimeError('iwvi_gp_fullcov_ws_doubles: bad descriptor')
    return int(n)


def gp_fullcov_fwd(d, S, N, aux, X, save, mean, eps, chol_jitter, cov, sample, info, ws=None):
    _count(1)
    L.check(L.load().iwvi_gp_fullcov_fwd(C.byref(d), int(S), int(N), _ptr(aux), _ptr(X), _ptr(save), _ptr(mean),
                                         _ptr(eps), float(chol_jitter), _ptr(cov), _ptr(sample), _ptr(info), _ptr(ws),
                                         _stream()), 'iwvi_gp_fullcov_fwd')


def gp_fullcov_bwd(d, S, N, aux, X, save, eps, chol_jitter, d_sample, d_cov, save2, dX_knn, part, ws=None):
    _count(1)
    L.check(L.load().iwvi_gp_fullcov_bwd(C.byref(d), int(S), int(N), _ptr(aux), _ptr(X), _ptr(save), _ptr(eps),
                                         float(chol_jitter), _ptr(d_sample), _ptr(d_cov), _ptr(save2), _ptr(dX_knn),
                                         _ptr(part), _ptr(ws), _stream()), 'iwvi_gp_fullcov_bwd')


def gauss_kl_fwd(M, R, q_mu, q_sqrt, kl):
    _count(1)
    L.check(L.load().iwvi_gauss_kl_fwd(int(M), int(R), _ptr(q_mu), _ptr(q_sqrt), _ptr(kl), _stream()), 'iwvi_gauss_kl_fwd')


def gauss_kl_bwd(M, R, q_mu, q_sqrt, dkl, dq_mu, dq_sqrt):
    _count(1)
    L.check(L.load().iwvi_gauss_kl_bwd(int(M), int(R), _ptr(q_mu), _ptr(q_sqrt), _ptr(dkl), _ptr(dq_mu), _ptr(dq_sqrt),
                                       _stream()), 'iwvi_gauss_kl_bwd')


def lv_fwd(d, F, enc_in, params, eps, samples, kl, mu, sigma):
    _count(1)
    L.check(L.load().iwvi_lv_fwd(C.byref(d), _ptr(F), _ptr(enc_in), _ptr(params), _ptr(eps), _ptr(samples), _ptr(kl),
                                 _ptr(mu), _ptr(sigma), _stream()), 'iwvi_lv_fwd')


def lv_bwd(d, F, enc_in, params, eps, mu, sigma, d_samples, d_kl, d_mu, d_sigma, d_params, dF, ws):
    _count(1 if d.prior else 2)
    L.check(L.load().iwvi_lv_bwd(C.byref(d), _ptr(F), _ptr(enc_in), _ptr(params), _ptr(eps), _ptr(mu), _ptr(sigma),
                                 _ptr(d_samples), _ptr(d_kl), _ptr(d_mu), _ptr(d_sigma), _ptr(d_params), _ptr(dF),
                                 _ptr(ws), _stream()), 'iwvi_lv_bwd')


def iwelbo_fwd(d, fmean, fvar, Y, lik_var, kl_local, elbo_data, logp, w, ws):
    _count(2)
    L.check(L.load().iwvi_iwelbo_fwd(C.byref(d), _ptr(fmean), _ptr(fvar), _ptr(Y), _ptr(lik_var), _ptr(kl_local),
                                     _ptr(elbo_data), _ptr(logp), _ptr(w), _ptr(ws), _stream()), 'iwvi_iwelbo_fwd')


def iwelbo_bwd(d, fmean, fvar, Y, lik_var, w, d_elbo, dmean, dvar, dkl_local, dlik, ws):
    _count(2)
    L.check(L.load().iwvi_iwelbo_bwd(C.byref(d), _ptr(fmean), _ptr(fvar), _ptr(Y), _ptr(lik_var), _ptr(w),
                                     _ptr(d_elbo), _ptr(dmean), _ptr(dvar), _ptr(dkl_local), _ptr(dlik), _ptr(ws),
                                     _stream()), 'iwvi_iwelbo_bwd')


def normal_fill(out, n_points, C_, first_point, seed):
    _count(1)
    L.check(L.load().iwvi_normal_fill(_ptr(out), int(n_points), int(C_), int(first_point),
                                      int(seed) & 0xFFFFFFFFFFFFFFFF, _stream()), 'iwvi_normal_fill')


def normal_fill_counter(out, n_points, C_, first_point, seed_base, layer, step_add, state):
    _count(1)
    L.check(L.load().iwvi_normal_fill_counter(_ptr(out), int(n_points), int(C_), int(first_point),
                                              int(seed_base) & 0xFFFFFFFFFFFFFFFF, int(layer), int(step_add), _ptr(state),
                                              _stream()), 'iwvi_normal_fill_counter')


def adam_step_counter(x, grad_elbo, m, v, mask, theta_pos, n, n_pos, lr_dev, beta1, beta2, eps, state):
    _count(2)
    L.check(L.load().iwvi_adam_step_counter(_ptr(x), _ptr(grad_elbo), _ptr(m), _ptr(v), _ptr(mask), _ptr(theta_pos), int(n),
                                            int(n_pos), _ptr(lr_dev), float(beta1), float(beta2), float(eps), _ptr(state),
                                            _stream()), 'iwvi_adam_step_counter')


def adam_step_counter_part(x, grad_elbo, m, v, mask, theta_pos, n, n_pos, lr_dev, beta1, beta2, eps, state, advance):
    _count(2 if advance else 1)
    L.check(L.load().iwvi_adam_step_counter_part(_ptr(x), _ptr(grad_elbo), _ptr(m), _ptr(v), _ptr(mask), _ptr(theta_pos),
                                                 int(n), int(n_pos), _ptr(lr_dev), float(beta1), float(beta2), float(eps),
                                                 _ptr(state), int(bool(advance)), _stream()), 'iwvi_adam_step_counter_part')


def positive_fwd(x, theta, n):
    _count(1)
    L.check(L.load().iwvi_positive_fwd(_ptr(x), _ptr(theta), int(n), _stream()), 'iwvi_positive_fwd')


def adam_step(x, grad_elbo, m, v, mask, theta_pos, n, n_pos, lr, beta1, beta2, eps, t):
    _count(1)
    L.check(L.load().iwvi_adam_step(_ptr(x), _ptr(grad_elbo), _ptr(m), _ptr(v), _ptr(mask), _ptr(theta_pos), int(n),
                                    int(n_pos), float(lr), float(beta1), float(beta2), float(eps), int(t), _stream()),
            'iwvi_adam_step')


def probe_dmma(out, blocks, warps, iters):
    """Measurement utility: register-only DMMA issue loop (csrc/probe.cu); 2*256*8*iters*warps*blocks flops."""
    L.check(L.load().iwvi_probe_dmma(_ptr(out), int(blocks), int(warps), int(iters), _stream()), 'iwvi_probe_dmma')
    return 2.0 * 256 * 8 * iters * warps * blocks


def batch_gather(X, Y, idx, B, Dx, Dy, Xb, Yb, XYb):
    _count(1)
    L.check(L.load().iwvi_batch_gather(_ptr(X), _ptr(Y), _ptr(idx), int(B), int(Dx), int(Dy), _ptr(Xb), _ptr(Yb),
                                       _ptr(XYb), _stream()), 'iwvi_batch_gather')


def dp_push(g, index, n, seg_off, bucket_len, seg, rank, world, recv_ptrs, flag_ptrs, epoch, counters):
    """recv_ptrs / flag_ptrs: ctypes arrays (c_uint64 * world) of device addresses (host memory)."""
    _count(1)
    L.check(L.load().iwvi_dp_push(_ptr(g), _ptr(index), int(n), int(seg_off), int(bucket_len), int(seg), int(rank),
                                  int(world), C.cast(recv_ptrs, C.c_void_p), C.cast(flag_ptrs, C.c_void_p), _ptr(epoch),
                                  _ptr(counters), _stream()), 'iwvi_dp_push')


def dp_reduce(g, index, n, seg_off, bucket_len, seg, world, recv_local, flags_local, epoch, counters):
    """recv_local / flags_local: integer device addresses of this rank's own receive buffer and flag array."""
    _count(2)
    L.check(L.load().iwvi_dp_reduce(_ptr(g), _ptr(index), int(n), int(seg_off), int(bucket_len), int(seg), int(world),
                                    C.c_void_p(int(recv_local)), C.c_void_p(int(flags_local)), _ptr(epoch),
                                    _ptr(counters), _stream()), 'iwvi_dp_reduce')
