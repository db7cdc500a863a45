"""
TEST INFRASTRUCTURE ONLY -- numpy restatement of the IW-ELBO path *in the stage decomposition the CUDA
kernels use* (prologue / rows / reduce / LV / IW-ELBO, forward and hand-derived backward), so every CUDA
stage can be compared with a CPU array of the same meaning, and the hand-derived adjoints
 are verified against torch.autograd on oracle/iwvi_oracle.py before any kernel
relies on them (tests/test_staged.py).

Stage <-> reference mapping:
  gp_prologue_fwd : temp_workaround.py:39,48 (Kuu + jitter, Cholesky) and :167-188 -> GPflow gauss_kl
  gp_rows_fwd     : temp_workaround.py:44-91 (Kuf, TRSM, fvar, fmean, LTA, sample), :142-145 (Mok mixing),
                    layers.py:46-48 (mean function)
  lv_fwd          : layers.py:72-105, :137-152
  iwelbo_fwd      : models.py:133-150 (IW) / :66-86 (VI)
Backward stages are the adjoints of these (the reference gets them from tf.gradients).
"""
import math

import numpy as np
from scipy.linalg import cholesky, solve_triangular

KERN_IDS = {'RBF': 0, 'Matern12': 1, 'Matern32': 2, 'Matern52': 3}


def k_of_r2(kind, r2, variance):
    """Returns (K, dK/dr2). Matern family clamps r2 at 1e-40 with zero gradient where clamped."""
    if kind == 'RBF':
        K = variance * np.exp(-0.5 * r2)
        return K, -0.5 * K
    clamped = r2 < 1e-40
    r = np.sqrt(np.maximum(r2, 1e-40))
    if kind == 'Matern52':
        s5 = math.sqrt(5.0)
        e = np.exp(-s5 * r)
        K = variance * (1 + s5 * r + 5.0 / 3.0 * r * r) * e
        dK = -(5.0 / 6.0) * variance * (1 + s5 * r) * e
    elif kind == 'Matern32':
        s3 = math.sqrt(3.0)
        e = np.exp(-s3 * r)
        K = variance * (1 + s3 * r) * e
        dK = -1.5 * variance * e
    elif kind == 'Matern12':
        e = np.exp(-r)
        K = variance * e
        dK = -variance * e / (2 * r)
    else:
        raise ValueError(kind)
    return K, np.where(clamped, 0.0, dK)


def sqdist(Xs, Zs):
    """Expanded form on already length-scaled inputs: |z|^2 + |x|^2 - 2 z.x  -> [M, T]"""
    return (Zs ** 2).sum(1)[:, None] + (Xs ** 2).sum(1)[None, :] - 2.0 * Zs @ Xs.T


# ------------------------------------------------------------------ GP layer ----------------------

def gp_prologue_fwd(kind, Z, ls, variance, q_mu, q_sqrt, jitter):
    M, R = q_mu.shape
    Zs = Z / ls
    Kuu, _ = k_of_r2(kind, sqdist(Zs, Zs), variance)
    Kuu = Kuu + jitter * np.eye(M)
    Lm = cholesky(Kuu, lower=True)
    Lq = np.tril(q_sqrt)
    diag = np.diagonal(Lq, axis1=1, axis2=2)
    kl = 0.5 * ((q_mu ** 2).sum() - M * R - np.log(diag ** 2).sum() + (Lq ** 2).sum())
    return Lm, kl


def gp_rows_fwd(kind, X, Z, ls, variance, Lm, q_mu, q_sqrt, W, mf, mfA, mfb, eps):
    """X [T,D]. Returns dict with outputs (sample/mean/var [T,P]) and saved tensors (A [M,T], U [R,M,T],
    gvar/gmean/gsample [T,R])."""
    Zs, Xs = Z / ls, X / ls
    Kuf, _ = k_of_r2(kind, sqdist(Xs, Zs), variance)
    A = solve_triangular(Lm, Kuf, lower=True)
    fvar0 = variance - (A ** 2).sum(0)
    gmean = A.T @ q_mu
    Lq = np.tril(q_sqrt)
    U = np.einsum('rab,an->rbn', Lq, A)            # U_r = Lq_r^T A
    gvar = fvar0[:, None] + (U ** 2).sum(1).T
    gsample = None if eps is None else gmean + eps * np.sqrt(gvar)
    if mf == 'Zero':
        m = 0.0
    elif mf == 'Identity':
        m = X
    else:
        m = X @ mfA + mfb
    if W is not None:
        sample = None if gsample is None else gsample @ W.T + m
        mean = gmean @ W.T + m
        var = gvar @ (W ** 2).T
    else:
        sample = None if gsample is None else gsample + m
        mean = gmean + m
        var = gvar
    return dict(sample=sample, mean=mean, var=var, A=A, U=U, gvar=gvar, gmean=gmean, gsample=gsample, Kuf=Kuf)


def gp_rows_bwd(kind, X, Z, ls, variance, Lm, q_mu, q_sqrt, W, mf, mfA, mfb, eps, saved, ds, dm, dv):
    """Adjoint of gp_rows_fwd. ds/dm/dv: cotangents of sample/mean/var [T,P] (None = zero).
    Returns dict: dX, dZ, dls, dvariance (rows part), dq_mu, dq_sqrt (tril), dLm (tril), dW, dmfA, dmfb."""
    T, D = X.shape
    M, R = q_mu.shape
    A, U, gvar, gmean, gsample = saved['A'], saved['U'], saved['gvar'], saved['gmean'], saved['gsample']
    P = R if W is None else W.shape[0]
    z = lambda a: np.zeros((T, P)) if a is None else a
    ds_, dm_, dv_ = z(ds), z(dm), z(dv)
    Wm = np.eye(R) if W is None else W
    gs_bar = ds_ @ Wm                                  # [T,R]
    gmean_bar = gs_bar + dm_ @ Wm
    gvar_bar = dv_ @ (Wm ** 2)
    if eps is not None:
        gvar_bar = gvar_bar + gs_bar * eps / (2.0 * np.sqrt(gvar))
    out = {}
    if W is not None:
        dW = dm_.T @ gmean + 2.0 * Wm * (dv_.T @ gvar)
        if gsample is not None:
            dW = dW + ds_.T @ gsample
        out['dW'] = dW
    dmf = ds_ + dm_
    dX = np.zeros((T, D))
    if mf == 'Identity':
        dX += dmf
    elif mf == 'Linear':
        dX += dmf @ mfA.T
        out['dmfA'] = X.T @ dmf
        out['dmfb'] = dmf.sum(0)
    dvariance = gvar_bar.sum()
    Lq = np.tril(q_sqrt)
    V = U * gvar_bar.T[:, None, :]                     # [R,M,T]  U_r * gvar_bar_r
    Abar = q_mu @ gmean_bar.T - 2.0 * A * gvar_bar.sum(1)[None, :] + 2.0 * np.einsum('rab,rbn->an', Lq, V)
    out['dq_mu'] = A @ gmean_bar
    out['dq_sqrt'] = np.tril(2.0 * np.einsum('an,rbn->rab', A, V))
    Bbar = solve_triangular(Lm, Abar, lower=True, trans='T')
    out['dLm'] = -np.tril(Bbar @ A.T)
    # gram backward through Kuf
    Zs, Xs = Z / ls, X / ls
    Kuf, dK = k_of_r2(kind, sqdist(Xs, Zs), variance)
    G = Bbar * dK                                      # [M,T]
    dvariance += (Bbar * Kuf).sum() / variance
    diff = Xs[None, :, :] - Zs[:, None, :]             # [M,T,D]  x~ - z~
    dX += 2.0 * np.einsum('mn,mnd->nd', G, diff) / ls
    out['dZ'] = -2.0 * np.einsum('mn,mnd->md', G, diff) / ls
    out['dls'] = -2.0 * np.einsum('mn,mnd->d', G, diff ** 2) / ls
    out['dX'] = dX
    out['dvariance'] = dvariance
    out['Bbar'] = Bbar
    out['gmean_bar'] = gmean_bar
    out['gvar_bar'] = gvar_bar
    return out


def gp_prologue_bwd(kind, Z, ls, variance, q_mu, q_sqrt, jitter, Lm, dLm, dkl):
    """Adjoint of gp_prologue_fwd: dLm (lower) and the scalar cotangent dkl -> dZ, dls, dvariance, dq_mu, dq_sqrt."""
    M, R = q_mu.shape
    Pm = np.tril(Lm.T @ np.tril(dLm))
    Pm[np.diag_indices(M)] *= 0.5
    S = solve_triangular(Lm, Pm, lower=True, trans='T')              # Lm^-T P
    S = solve_triangular(Lm, S.T, lower=True, trans='T').T           # (Lm^-T (Lm^-T P)^T)^T = Lm^-T P Lm^-1
    Kbar = 0.5 * (S + S.T)
    Zs = Z / ls
    r2 = sqdist(Zs, Zs)
    Kuu, dK = k_of_r2(kind, r2, variance)
    G = Kbar * dK
    np.fill_diagonal(G, 0.0)   # d r2_ii / d anything == 0
    dvariance = (Kbar * Kuu).sum() / variance
    diff = Zs[None, :, :] - Zs[:, None, :]             # [i,j,d] = z~_j - z~_i   (x role = j, z role = i)
    dZ = 2.0 * np.einsum('ij,ijd->jd', G, diff) / ls - 2.0 * np.einsum('ij,ijd->id', G, diff) / ls
    dls = -2.0 * np.einsum('ij,ijd->d', G, diff ** 2) / ls
    Lq = np.tril(q_sqrt)
    dq_mu = dkl * q_mu
    dq_sqrt = dkl * (Lq - np.stack([np.diag(1.0 / np.diag(Lq[r])) for r in range(R)]))
    return dict(dZ=dZ, dls=dls, dvariance=dvariance, dq_mu=dq_mu, dq_sqrt=dq_sqrt, Kbar=Kbar)


# ------------------------------------------------------------------ LV layer ----------------------

def softplus(x):
    return np.logaddexp(0.0, x)


def encoder_fwd(Ws, bs, H, latent_dim):
    acts = [H]
    n = len(Ws)
    for i, (W, b) in enumerate(zip(Ws, bs)):
        h = acts[-1]
        a = h @ W + b
        if i < n - 1:
            a = np.tanh(a)
        if W.shape[0] == W.shape[1]:
            a = a + h
        acts.append(a)
    out = acts[-1]
    mu, raw = out[:, :latent_dim], out[:, latent_dim:]
    sigma = softplus(raw - 3.0)
    return mu, sigma, acts


def encoder_bwd(Ws, bs, acts, latent_dim, dmu, dsigma):
    n = len(Ws)
    raw = acts[-1][:, latent_dim:]
    dout = np.concatenate([dmu, dsigma / (1.0 + np.exp(-(raw - 3.0)))], 1)
    dWs, dbs = [None] * n, [None] * n
    for i in reversed(range(n)):
        h = acts[i]
        W = Ws[i]
        skip = W.shape[0] == W.shape[1]
        dh = dout.copy() if skip else 0.0
        if i < n - 1:
            t = acts[i + 1] - (h if skip else 0.0)     # tanh output
            dpre = dout * (1.0 - t * t)
        else:
            dpre = dout
        dWs[i] = h.T @ dpre
        dbs[i] = dpre.sum(0)
        dout = dpre @ W.T + dh
    return dWs, dbs


def lv_fwd(Ws, bs, latent_dim, F, enc_in, eps, Kt, sampled):
    """F [Be,Df] and enc_in [Be,Dxy] are broadcast Kt times (point index = n*Kt + k); eps [Be*Kt, Lw].
    sampled: IW per-sample log q/p (layers.py:100) else closed-form KL (layers.py:103)."""
    Be = enc_in.shape[0]
    mu, sigma, acts = encoder_fwd(Ws, bs, enc_in, latent_dim)
    mu_t = np.repeat(mu, Kt, 0); sig_t = np.repeat(sigma, Kt, 0)
    Wlat = mu_t + eps * sig_t
    samples = np.concatenate([np.repeat(F, Kt, 0), Wlat], 1)
    if sampled:
        kl = -0.5 * ((Wlat - mu_t) / sig_t) ** 2 - np.log(sig_t) + 0.5 * Wlat ** 2
    else:
        kl = 0.5 * mu_t ** 2 + 0.5 * (sig_t ** 2 - 1.0 - np.log(sig_t ** 2))
    return dict(samples=samples, kl=kl, mu=mu, sigma=sigma, acts=acts, Wlat=Wlat)


def lv_bwd(Ws, bs, latent_dim, F, enc_in, eps, Kt, sampled, saved, d_samples, d_kl):
    Be = enc_in.shape[0]
    Df = F.shape[1]
    mu, sigma, Wlat = saved['mu'], saved['sigma'], saved['Wlat']
    sig_t = np.repeat(sigma, Kt, 0); mu_t = np.repeat(mu, Kt, 0)
    Wbar = d_samples[:, Df:].copy()
    if sampled:
        Wbar = Wbar + d_kl * Wlat
        mubar_t = Wbar
        sigbar_t = Wbar * eps - d_kl / sig_t
    else:
        mubar_t = Wbar + d_kl * mu_t
        sigbar_t = Wbar * eps + d_kl * (sig_t - 1.0 / sig_t)
    mubar = mubar_t.reshape(Be, Kt, -1).sum(1)
    sigbar = sigbar_t.reshape(Be, Kt, -1).sum(1)
    dWs, dbs = encoder_bwd(Ws, bs, saved['acts'], latent_dim, mubar, sigbar)
    dF = d_samples[:, :Df].reshape(Be, Kt, Df).sum(1)
    return dict(dWs=dWs, dbs=dbs, dF=dF)


# ------------------------------------------------------------------ IW-ELBO ----------------------

def iwelbo_fwd(fmean, fvar, Y, lik_var, kl_local, K, scale, iw=True, data_major=True):
    """fmean/fvar [T,Dy]; Y [B,Dy]; kl_local [T,Lw] or None. data_major: point = n*K + k (IW layout,
    models.py:113); else point = k*B + n (VI layout, models.py:50)."""
    B = Y.shape[0]
    Yt = np.repeat(Y, K, 0) if data_major else np.tile(Y, (K, 1))
    ve = -0.5 * math.log(2 * math.pi) - 0.5 * np.log(lik_var) - 0.5 * ((Yt - fmean) ** 2 + fvar) / lik_var
    L = ve.sum(1)
    if kl_local is not None:
        L = L - kl_local.sum(1)
    L_NK = L.reshape(B, K) if data_major else L.reshape(K, B).T
    if iw:
        mx = L_NK.max(1, keepdims=True)
        e = np.exp(L_NK - mx)
        s = e.sum(1, keepdims=True)
        logp = (mx + np.log(s))[:, 0] - math.log(K)
        w = e / s
    else:
        logp = L_NK.mean(1)
        w = np.full_like(L_NK, 1.0 / K)
    return dict(elbo_data=scale * logp.sum(), logp=logp, w=w, L_NK=L_NK)


def iwelbo_bwd(fmean, fvar, Y, lik_var, kl_local, K, scale, saved, d_elbo, data_major=True):
    B = Y.shape[0]
    g_nk = d_elbo * scale * saved['w']                 # [B,K]
    g = g_nk.reshape(-1) if data_major else g_nk.T.reshape(-1)
    Yt = np.repeat(Y, K, 0) if data_major else np.tile(Y, (K, 1))
    dmean = g[:, None] * (Yt - fmean) / lik_var
    dvar = -g[:, None] / (2 * lik_var) * np.ones_like(fvar)
    dlik = (g[:, None] * (-0.5 / lik_var + 0.5 * ((Yt - fmean) ** 2 + fvar) / lik_var ** 2)).sum()
    dkl = None if kl_local is None else -g[:, None] * np.ones_like(kl_local)
    return dict(dmean=dmean, dvar=dvar, dlik=dlik, dkl=dkl)


# ------------------------------------------------------------------ whole model -------------------

def iw_elbo_and_grads(spec, X, Y, eps, iw=True):
    """Full IW-ELBO forward + hand-written backward on a spec (same dict as iwvi_oracle.build_from_spec).
    X [B,Dx], Y [B,1], eps per layer with IW index order [B,K,C].  Returns (elbo, grads-by-leaf-name)."""
    K = spec['num_samples']
    B = X.shape[0]
    T = B * K
    scale = spec['num_data'] / B
    layers = spec['layers']
    tape = []
    F = None         # current [T, D] samples; None means "X broadcast"
    kl_local = []
    kl_global = []
    for i, ls_ in enumerate(layers):
        e = None if eps[i] is None else np.asarray(eps[i]).reshape(T, -1)
        if ls_['type'] == 'lv':
            if F is None:
                Fin, enc_in, Kt = X, np.concatenate([X, Y], 1), K
            else:
                Fin, enc_in, Kt = F, np.repeat(np.concatenate([X, Y], 1), K, 0), 1
            o = lv_fwd(ls_['Ws'], ls_['bs'], ls_['latent_dim'], Fin, enc_in, e, Kt, sampled=iw)
            tape.append(('lv', i, Fin, enc_in, e, Kt, o))
            F = o['samples']
            kl_local.append(o['kl'])
        else:
            if F is None:
                F = np.repeat(X, K, 0)
            lsc = np.broadcast_to(np.asarray(ls_['lengthscales'], dtype=np.float64), (F.shape[1],)).copy()
            var = float(ls_['variance'])
            jit = ls_.get('jitter', 1e-6)
            Lm, kl = gp_prologue_fwd(ls_['kern'], ls_['Z'], lsc, var, ls_['q_mu'], ls_['q_sqrt'], jit)
            o = gp_rows_fwd(ls_['kern'], F, ls_['Z'], lsc, var, Lm, ls_['q_mu'], ls_['q_sqrt'], ls_.get('W'),
                            ls_['mf'], ls_.get('mf_A'), ls_.get('mf_b'), e)
            tape.append(('gp', i, F, lsc, var, jit, Lm, e, o))
            kl_global.append(kl)
            last = o
            F = o['sample']
    klc = np.concatenate(kl_local, 1) if kl_local else None
    top = iwelbo_fwd(last['mean'], last['var'], Y, float(spec['lik_variance']), klc, K, scale, iw=iw)
    elbo = top['elbo_data'] - sum(kl_global)
    # ---- backward
    grads = {}
    tb = iwelbo_bwd(last['mean'], last['var'], Y, float(spec['lik_variance']), klc, K, scale, top, 1.0)
    grads['likelihood.variance'] = np.array(tb['dlik'])
    ds, dm, dv = None, tb['dmean'], tb['dvar']
    kl_off = klc.shape[1] if klc is not None else 0
    for entry in reversed(tape):
        if entry[0] == 'gp':
            _, i, Fin, lsc, var, jit, Lm, e, o = entry
            ls_ = layers[i]
            p = 'layers.%d.' % i
            b = gp_rows_bwd(ls_['kern'], Fin, ls_['Z'], lsc, var, Lm, ls_['q_mu'], ls_['q_sqrt'], ls_.get('W'),
                            ls_['mf'], ls_.get('mf_A'), ls_.get('mf_b'), e, o, ds, dm, dv)
            pb = gp_prologue_bwd(ls_['kern'], ls_['Z'], lsc, var, ls_['q_mu'], ls_['q_sqrt'], jit, Lm, b['dLm'], -1.0)
            grads[p + 'Z'] = b['dZ'] + pb['dZ']
            dls = b['dls'] + pb['dls']
            grads[p + 'kern.lengthscales'] = dls if np.ndim(ls_['lengthscales']) else np.array(dls.sum())
            grads[p + 'kern.variance'] = np.array(b['dvariance'] + pb['dvariance'])
            grads[p + 'q_mu'] = b['dq_mu'] + pb['dq_mu']
            grads[p + 'q_sqrt'] = b['dq_sqrt'] + pb['dq_sqrt']
            if ls_.get('W') is not None:
                grads[p + 'kern.W'] = b['dW']
            if ls_['mf'] == 'Linear':
                grads[p + 'mf.A'] = b['dmfA']; grads[p + 'mf.b'] = b['dmfb']
            ds, dm, dv = b['dX'], None, None
        else:
            _, i, Fin, enc_in, e, Kt, o = entry
            ls_ = layers[i]
            p = 'layers.%d.' % i
            Lw = ls_['latent_dim']
            kl_off -= Lw
            d_kl = tb['dkl'][:, kl_off:kl_off + Lw]
            b = lv_bwd(ls_['Ws'], ls_['bs'], Lw, Fin, enc_in, e, Kt, iw, o, ds, d_kl)
            for j in range(len(ls_['Ws'])):
                grads[p + 'encoder.Ws.%d' % j] = b['dWs'][j]
                grads[p + 'encoder.bs.%d' % j] = b['dbs'][j]
            ds = b['dF'] if Kt == 1 else None
    return elbo, grads
