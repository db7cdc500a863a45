"""GPU parity of the operator-level boundary (dgps_with_iwvi_b200/temp_workaround.py and the layers' `propagate`
protocol -- the reference's temp_workaround.py:118-188 and layers.py:35-50,72-105,137-152) against the CPU oracle,
values and torch-autograd gradients.  Tolerance rtol 1e-8 relative to the largest entry (float64)."""
import numpy as np
import pytest
import torch

from oracle import iwvi_oracle as O
from oracle import svgp_closed_form as SV
from oracle import synthetic as S

pytestmark = pytest.mark.gpu
RTOL = 1e-8
T64 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)


def close(name, got, want, rtol=RTOL):
    got = got.detach().cpu().numpy() if torch.is_tensor(got) else np.asarray(got)
    want = want.detach().cpu().numpy() if torch.is_tensor(want) else np.asarray(want)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    # an all-zero reference (e.g. log q - log p under the prior branch, where q == p) is compared absolutely
    err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-6)
    assert np.isfinite(got).all() and err < rtol, '%s: %.3e' % (name, err)


def _layer_pair(conf, D, M, kern='RBF', final_mf='Zero', which=1, seed=3):
    """(product GPLayer, oracle GPLayer, spec entry) for layer `which` of a synthetic spec."""
    from dgps_with_iwvi_b200.build_models import model_from_spec
    X, Y = S.make_data(40, D, seed=seed)
    spec = S.make_spec(X, conf, M, 3, seed=seed, perturb=0.3, inner_q_sqrt_scale=0.3, kern=kern, final_mf=final_mf)
    model = model_from_spec(spec, X, Y)
    omodel, leaves = O.build_from_spec(spec, requires_grad=True)
    return model.layers[which], omodel.layers[which], leaves, 'layers.%d.' % which


@pytest.mark.parametrize('conf,which,kern', [('L1_G3', 1, 'RBF'), ('L1_G3', 2, 'Matern52'), ('G2', 0, 'Matern32')])
def test_gplayer_propagate_values_and_autograd(conf, which, kern):
    D = 4
    layer, olayer, leaves, pre = _layer_pair(conf, D, 37, kern=kern, which=which, final_mf='Linear')
    Din = olayer.Z.shape[1]
    rng = np.random.default_rng(1)
    Sn, N = 5, 9
    F = rng.standard_normal((Sn, N, Din))
    R = olayer.q_mu.shape[1]
    eps = rng.standard_normal((Sn, N, R))
    Fo = T64(F).requires_grad_(True)
    so, mo, vo, klo = olayer.propagate(Fo, eps=T64(eps))
    Fg = T64(F).cuda().requires_grad_(True)
    for _, p in layer.named_parameters():
        p.unconstrained.requires_grad_(True)
    s, m, v, kl = layer.propagate(Fg, eps=T64(eps).cuda())
    close('sample', s, so); close('mean', m, mo); close('var', v, vo); close('kl', kl, klo)
    cs, cm, cv = [T64(rng.standard_normal(so.shape)) for _ in range(3)]
    (so * cs).sum().add((mo * cm).sum()).add((vo * cv).sum()).sub(klo).backward()
    (s * cs.cuda()).sum().add((m * cm.cuda()).sum()).add((v * cv.cuda()).sum()).sub(kl).backward()
    close('dF', Fg.grad, Fo.grad)
    mix = hasattr(layer.kern, 'W')
    base = layer.kern.kernel if mix else layer.kern
    feat = layer.feature.feat if hasattr(layer.feature, 'feat') else layer.feature
    # q_sqrt / lengthscales / variance carry transforms: compare the constrained-space gradient through the chain
    close('dZ', feat.Z.unconstrained.grad, leaves[pre + 'Z'].grad)
    close('dq_mu', layer.q_mu.unconstrained.grad, leaves[pre + 'q_mu'].grad)
    close('dq_sqrt', layer.q_sqrt.unconstrained.grad, torch.tril(leaves[pre + 'q_sqrt'].grad))
    sig = lambda p: torch.sigmoid(p.unconstrained.detach())
    close('dls', base.lengthscales.unconstrained.grad / sig(base.lengthscales), leaves[pre + 'kern.lengthscales'].grad)
    close('dvariance', base.variance.unconstrained.grad / sig(base.variance), leaves[pre + 'kern.variance'].grad)
    if mix:
        close('dW', layer.kern.W.unconstrained.grad, leaves[pre + 'kern.W'].grad)
    if layer.mean_function.kind == 'Linear':
        close('dA', layer.mean_function.A.unconstrained.grad, leaves[pre + 'mf.A'].grad)
        close('db', layer.mean_function.b.unconstrained.grad, leaves[pre + 'mf.b'].grad)


def test_gauss_kl_operator():
    from dgps_with_iwvi_b200 import temp_workaround as tw
    rng = np.random.default_rng(2)
    M, R = 70, 3
    q_mu = rng.standard_normal((M, R))
    q_sqrt = np.tril(rng.standard_normal((R, M, M))) * 0.1 + np.eye(M)[None]
    a, b = T64(q_mu).requires_grad_(True), T64(q_sqrt).requires_grad_(True)
    ko = O.gauss_kl(a, b)
    ko.backward()
    ag, bg = T64(q_mu).cuda().requires_grad_(True), T64(q_sqrt).cuda().requires_grad_(True)
    k = tw.gauss_kl(ag, bg)
    (2.5 * k).backward()
    close('kl', k, ko); close('dq_mu', ag.grad, 2.5 * a.grad); close('dq_sqrt', bg.grad, 2.5 * torch.tril(b.grad))
    # K given (unwhitened prior N(0, K), temp_workaround.py:188 -> gpflow gauss_kl(q_mu, q_sqrt, K)): the textbook form
    A = rng.standard_normal((M, M)); Kp = A @ A.T / M + np.eye(M)
    ag2, bg2 = T64(q_mu).cuda().requires_grad_(True), T64(q_sqrt).cuda().requires_grad_(True)
    k2 = tw.gauss_kl(ag2, bg2, K=T64(Kp).cuda())
    want = 0.0
    Ki = np.linalg.inv(Kp)
    for r in range(R):
        Sr = np.tril(q_sqrt[r]) @ np.tril(q_sqrt[r]).T
        want += 0.5 * (np.trace(Ki @ Sr) + q_mu[:, r] @ Ki @ q_mu[:, r] - M + np.linalg.slogdet(Kp)[1]
                       - np.linalg.slogdet(Sr)[1])
    assert abs(k2.item() - want) < 1e-9 * abs(want)
    k2.backward()
    close('dq_mu (K)', ag2.grad, T64(Ki @ q_mu), rtol=1e-7)
    # q_sqrt = None (:174-184): minus the log density of q_mu under N(0, K + jitter I)
    k3 = tw.gauss_kl(T64(q_mu).cuda(), None, K=T64(Kp).cuda())
    close('nlp (K)', k3, O.gauss_kl(T64(q_mu), None, K=T64(Kp)))


@pytest.mark.parametrize('sampled,amortised', [(True, True), (False, True), (True, False)])
def test_latent_variable_layer_propagate(sampled, amortised):
    from dgps_with_iwvi_b200.build_models import model_from_spec
    D = 3
    X, Y = S.make_data(30, D, seed=5)
    spec = S.make_spec(X, 'L2_G2', 10, 3, seed=5)
    layer = model_from_spec(spec, X, Y).layers[0]
    omodel, leaves = O.build_from_spec(spec, requires_grad=True)
    olayer = omodel.layers[0]
    rng = np.random.default_rng(6)
    N, K = 7, 4
    F = rng.standard_normal((N, K, D)); XY = rng.standard_normal((N, K, D + 1)); eps = rng.standard_normal((N, K, 2))
    Fo = T64(F).requires_grad_(True)
    outs_o = olayer.propagate(Fo, T64(XY) if amortised else None, sampled, eps=T64(eps))
    Fg = T64(F).cuda().requires_grad_(True)
    for _, p in layer.named_parameters():
        p.unconstrained.requires_grad_(True)
    outs = layer.propagate(Fg, inference_amorization_inputs=T64(XY).cuda() if amortised else None,
                           is_sampled_local_regularizer=sampled, eps=T64(eps).cuda())
    cots = [T64(rng.standard_normal(o.shape)) for o in outs_o]
    for n, a, b in zip(['samples', 'mean', 'cov', 'kl'], outs, outs_o):
        close(n, a, b)
    sum((o * c).sum() for o, c in zip(outs_o, cots)).backward()
    sum((o * c.cuda()).sum() for o, c in zip(outs, cots)).backward()
    close('dF', Fg.grad, Fo.grad)
    if amortised:
        for j, (W, b) in enumerate(zip(layer.encoder.Ws, layer.encoder.bs)):
            close('dW%d' % j, W.unconstrained.grad, leaves['layers.0.encoder.Ws.%d' % j].grad)
            close('db%d' % j, b.unconstrained.grad, leaves['layers.0.encoder.bs.%d' % j].grad)
        mu, sg = layer.encoder(T64(XY).cuda())
        mo, so = olayer.encoder(T64(XY))
        close('enc mu', mu, mo); close('enc sigma', sg, so)


def test_full_cov_branch_and_predict_f_full_cov():
    """reference tests/test_gp_layer.py:15-54: single Matern52 GPLayer with Linear mean function == SVGP closed form
    (mean and FULL covariance); and the 3-D full_cov=True branch of the conditional (temp_workaround.py:55-57,82-83)."""
    from dgps_with_iwvi_b200 import temp_workaround as tw
    from dgps_with_iwvi_b200.build_models import model_from_spec
    rng = np.random.default_rng(8)
    N, D, M, Ns = 60, 2, 33, 150      # 150 test points: the covariance spans 3 x 3 blocks of 64, the last one ragged
    X, Y = S.make_data(N, D, seed=8)
    spec = S.make_spec(X, '', M, 1, seed=8, perturb=0.4, kern='Matern52', final_mf='Linear', lik_variance=0.1)
    g = spec['layers'][0]
    m = model_from_spec(spec, X, Y, mode='vi')
    Xs = rng.standard_normal((Ns, D))
    mean, cov = m.predict_f_full_cov(Xs)
    _, ref_mean, ref_cov = SV.svgp(g['kern'], g['variance'], g['lengthscales'], g['Z'], g['q_mu'], g['q_sqrt'], X, Y, Xs,
                                   float(spec['lik_variance']), mf_A=g.get('mf_A'), mf_b=g.get('mf_b'))
    # the closed form goes through an explicit Kuu solve: agreement is limited by its conditioning, not by the GPU path
    close('mean', mean, ref_mean, 1e-7); close('cov', cov, ref_cov, 1e-7)
    # 3-D branch against the oracle
    layer = m.layers[0]
    omodel, _ = O.build_from_spec(spec)
    F = rng.standard_normal((3, 11, D))
    _, mo, vo = O.independent_multisample_sample_conditional(T64(F), omodel.layers[0].Z, omodel.layers[0].kern,
                                                            omodel.layers[0].q_mu, full_cov=True,
                                                            q_sqrt=omodel.layers[0].q_sqrt, white=True)
    _, mg, vg = tw.multisample_sample_conditional(T64(F).cuda(), layer.feature, layer.kern, layer.q_mu, full_cov=True,
                                                  q_sqrt=layer.q_sqrt, white=True, sample=False)
    close('mean3', mg, mo); close('cov3', vg, vo)
    assert vg.shape == (3, 1, 11, 11)
    # more than 64 points per group: covariance in 64 x 64 blocks; the joint draw is built for up to 256 points
    F2 = rng.standard_normal((2, 97, D))
    _, mo2, vo2 = O.independent_multisample_sample_conditional(T64(F2), omodel.layers[0].Z, omodel.layers[0].kern,
                                                              omodel.layers[0].q_mu, full_cov=True,
                                                              q_sqrt=omodel.layers[0].q_sqrt, white=True)
    _, mg2, vg2 = tw.multisample_sample_conditional(T64(F2).cuda(), layer.feature, layer.kern, layer.q_mu, full_cov=True,
                                                    q_sqrt=layer.q_sqrt, white=True, sample=False)
    close('mean3b', mg2, mo2); close('cov3b', vg2, vo2)
    assert torch.equal(vg2, vg2.transpose(-1, -2))
    F3 = rng.standard_normal((1, 257, D))
    with pytest.raises(NotImplementedError):
        tw.multisample_sample_conditional(T64(F3).cuda(), layer.feature, layer.kern, layer.q_mu, full_cov=True,
                                          q_sqrt=layer.q_sqrt, white=True)


@pytest.mark.parametrize('form', ['none', 'diag'])
def test_conditional_q_sqrt_none_and_diagonal(form):
    """temp_workaround.py:71-73: q_sqrt=None (SGHMC, no variance contribution) and the 2-D diagonal q_sqrt [M, R];
    gauss_kl's matching branches (:174-188)."""
    from dgps_with_iwvi_b200 import temp_workaround as tw
    from dgps_with_iwvi_b200.build_models import model_from_spec
    D, M = 3, 29
    X, Y = S.make_data(40, D, seed=7)
    spec = S.make_spec(X, 'G2', M, 3, seed=7, perturb=0.3, inner_q_sqrt_scale=0.3, kern='Matern32')
    layer = model_from_spec(spec, X, Y).layers[0]
    omodel, _ = O.build_from_spec(spec)
    ol = omodel.layers[0]
    rng = np.random.default_rng(3)
    F = rng.standard_normal((4, 6, D)); eps = rng.standard_normal((4, 6, 2))
    qd = np.abs(rng.standard_normal((M, 2))) + 0.2
    Fo, Fg = T64(F).requires_grad_(True), T64(F).cuda().requires_grad_(True)
    qo = None if form == 'none' else T64(qd).requires_grad_(True)
    qg = None if form == 'none' else T64(qd).cuda().requires_grad_(True)
    so, mo, vo = O.multisample_sample_conditional(Fo, ol.Z, ol.kern, ol.q_mu, q_sqrt=qo, white=True, eps=T64(eps))
    s, m, v = tw.multisample_sample_conditional(Fg, layer.feature, layer.kern, layer.q_mu, q_sqrt=qg, white=True,
                                                eps=T64(eps).cuda(), mean_function=None)
    close('sample', s, so); close('mean', m, mo); close('var', v, vo)
    c = T64(rng.standard_normal(so.shape))
    ((so * c).sum() + (vo * c).sum()).backward()
    ((s * c.cuda()).sum() + (v * c.cuda()).sum()).backward()
    close('dF', Fg.grad, Fo.grad)
    if form == 'diag':
        close('dq_sqrt', qg.grad, qo.grad)
    # the KL wrapper's branches
    ko = O.gauss_kl(ol.q_mu, None if form == 'none' else T64(np.stack([np.diag(qd[:, r]) for r in range(2)])))
    kg = tw.gauss_kl(layer.q_mu, None if form == 'none' else T64(qd).cuda())
    close('kl', kg, ko)


@pytest.mark.parametrize('S_,N,M,kern', [(5, 50, 100, 'RBF'), (3, 64, 37, 'Matern52'), (7, 8, 64, 'Matern32'),
                                         (4, 1, 29, 'RBF'), (2, 20, 130, 'Matern12'),
                                         # beyond one 64-point block: batched blocked Cholesky over global memory
                                         (3, 65, 40, 'RBF'), (2, 150, 100, 'Matern52'), (2, 256, 70, 'RBF')])
def test_full_cov_joint_draw(S_, N, M, kern):
    """iwvi_gp_fullcov_fwd: covariance over the inner axis [S, R, N, N] (temp_workaround.py:55-57,82-83) and the joint
    draw the reference intends at :92-96 (noise in its [S, R, N, 1] order), against the oracle; groups straddle the
    64-point chunks of the saved panels, N is not a multiple of 8, M is not a multiple of 64."""
    from dgps_with_iwvi_b200 import temp_workaround as tw
    from dgps_with_iwvi_b200.build_models import model_from_spec
    D = 3
    X, Y = S.make_data(max(M, 40), D, seed=11)
    spec = S.make_spec(X, 'G2', M, 3, seed=11, perturb=0.3, inner_q_sqrt_scale=0.3, kern=kern)
    layer = model_from_spec(spec, X, Y).layers[0]
    omodel, _ = O.build_from_spec(spec)
    ol = omodel.layers[0]
    okern = ol.kern.kernel if isinstance(ol.kern, O.Mok) else ol.kern
    gkern = layer.kern.kernel if hasattr(layer.kern, 'W') else layer.kern
    R = ol.q_mu.shape[1]
    rng = np.random.default_rng(5)
    # (large groups: points spread out, so that the N x N covariance of a smooth kernel stays well conditioned -- the
    #  reference factorises it without jitter, temp_workaround.py:95)
    F = rng.standard_normal((S_, N, D)) * (4.0 if N > 64 else 1.0); z = rng.standard_normal((S_, R, N))
    so, mo, vo = O.independent_multisample_sample_conditional(T64(F), ol.Z, okern, ol.q_mu, full_cov=True,
                                                              q_sqrt=ol.q_sqrt, white=True, eps_joint=T64(z))
    s, m, v = tw.independent_multisample_sample_conditional(T64(F).cuda(), layer.feature, gkern, layer.q_mu,
                                                            full_cov=True, q_sqrt=layer.q_sqrt, white=True,
                                                            eps=T64(z).cuda())
    assert v.shape == (S_, R, N, N) and s.shape == (S_, N, R)
    # Matern12 is not smooth at r = 0: r = sqrt(max(r2, 1e-40)) turns the rounding noise of the expanded-form r2 on the
    # Kuu diagonal into ~1e-8 noise, in the reference as much as here (see tests/test_gpu_stages.py)
    tol = 1e-6 if kern == 'Matern12' else RTOL
    close('mean', m, mo, tol); close('cov', v, vo, tol); close('sample', s, so, tol)
    # exact symmetry of the DMMA Gram products
    assert torch.equal(v, v.transpose(-1, -2))


@pytest.mark.parametrize('conf,which,kern,S_,N', [('G2', 0, 'RBF', 3, 20), ('L1_G3', 2, 'Matern52', 2, 50),
                                                  ('G3', 0, 'Matern32', 2, 64), ('G2', 1, 'RBF', 5, 1),
                                                  ('G2', 0, 'RBF', 2, 100), ('L1_G3', 2, 'Matern52', 3, 129),
                                                  ('G2', 0, 'Matern32', 1, 256)])
def test_full_cov_joint_draw_autograd(conf, which, kern, S_, N):
    """Gradients through the covariance over the inner axis and the joint draw (iwvi_gp_fullcov_bwd + the per-point
    backward kernels) against torch autograd on the oracle (Cholesky adjoint included): cotangents on the sample, the
    mean and the full covariance; gradients of the inputs and of every parameter.  Final layers (one output, Linear mean
    function) are taken as they are; of the Mok layers the shared base kernel and the latent q(u) are used (R > 1)."""
    from dgps_with_iwvi_b200 import temp_workaround as tw
    layer, ol, leaves, pre = _layer_pair(conf, 4, 37, kern=kern, which=which, final_mf='Linear')
    mok = isinstance(ol.kern, O.Mok)
    okern = ol.kern.kernel if mok else ol.kern
    gkern = layer.kern.kernel if mok else layer.kern
    gmf = None if mok else layer.mean_function
    Din, R = ol.Z.shape[1], ol.q_mu.shape[1]
    rng = np.random.default_rng(4)
    F = rng.standard_normal((S_, N, Din)) * (4.0 if N > 64 else 1.0); z = rng.standard_normal((S_, R, N))
    Fo = T64(F).requires_grad_(True)
    so, mo, vo = O.independent_multisample_sample_conditional(Fo, ol.Z, okern, ol.q_mu, full_cov=True, q_sqrt=ol.q_sqrt,
                                                              white=True, eps_joint=T64(z))
    if not mok:
        mf = ol.mean_function(Fo)
        so, mo = so + mf, mo + mf
    Fg = T64(F).cuda().requires_grad_(True)
    for _, p in layer.named_parameters():
        p.unconstrained.requires_grad_(True)
    s, m, v = tw.independent_multisample_sample_conditional(Fg, layer.feature, gkern, layer.q_mu, full_cov=True,
                                                            q_sqrt=layer.q_sqrt, white=True, eps=T64(z).cuda(),
                                                            mean_function=gmf)
    close('sample', s, so); close('mean', m, mo); close('cov', v, vo)
    cs, cm = T64(rng.standard_normal(so.shape)), T64(rng.standard_normal(mo.shape))
    cv = T64(rng.standard_normal(vo.shape))          # deliberately not symmetric
    (so * cs).sum().add((mo * cm).sum()).add((vo * cv).sum()).backward()
    (s * cs.cuda()).sum().add((m * cm.cuda()).sum()).add((v * cv.cuda()).sum()).backward()
    close('dF', Fg.grad, Fo.grad)
    feat = layer.feature.feat if hasattr(layer.feature, 'feat') else layer.feature
    close('dZ', feat.Z.unconstrained.grad, leaves[pre + 'Z'].grad)
    close('dq_mu', layer.q_mu.unconstrained.grad, leaves[pre + 'q_mu'].grad)
    close('dq_sqrt', layer.q_sqrt.unconstrained.grad, torch.tril(leaves[pre + 'q_sqrt'].grad))
    sig = lambda p: torch.sigmoid(p.unconstrained.detach())
    close('dls', gkern.lengthscales.unconstrained.grad / sig(gkern.lengthscales), leaves[pre + 'kern.lengthscales'].grad)
    close('dvariance', gkern.variance.unconstrained.grad / sig(gkern.variance), leaves[pre + 'kern.variance'].grad)
    if not mok:
        close('dA', layer.mean_function.A.unconstrained.grad, leaves[pre + 'mf.A'].grad)
        close('db', layer.mean_function.b.unconstrained.grad, leaves[pre + 'mf.b'].grad)


@pytest.mark.parametrize('conf,which,kern,form', [('L1_G3', 1, 'RBF', 'full'), ('G2', 0, 'Matern52', 'full'),
                                                  ('L1_G3', 2, 'RBF', 'diag'), ('G2', 0, 'RBF', 'none')])
def test_conditional_unwhitened(conf, which, kern, form):
    """white=False (reference temp_workaround.py:63-65: A <- Lm^-T A before the mean and the q_sqrt projection), values
    and every gradient -- including the extra dependence on Z / lengthscales / variance through Lm -- against the oracle."""
    from dgps_with_iwvi_b200 import temp_workaround as tw
    layer, ol, leaves, pre = _layer_pair(conf, 4, 45, kern=kern, which=which)
    Din = ol.Z.shape[1]
    rng = np.random.default_rng(9)
    Sn, N = 3, 11
    M, R = ol.q_mu.shape
    F = rng.standard_normal((Sn, N, Din)); eps = rng.standard_normal((Sn, N, R))
    if form == 'full':
        qo, qg = ol.q_sqrt, layer.q_sqrt
    elif form == 'diag':
        qd = np.abs(rng.standard_normal((M, R))) + 0.2
        qo, qg = T64(qd).requires_grad_(True), T64(qd).cuda().requires_grad_(True)
    else:
        qo = qg = None
    for _, p in layer.named_parameters():
        p.unconstrained.requires_grad_(True)
    Fo = T64(F).requires_grad_(True)
    so, mo, vo = O.multisample_sample_conditional(Fo, ol.Z, ol.kern, ol.q_mu, q_sqrt=qo, white=False, eps=T64(eps))
    Fg = T64(F).cuda().requires_grad_(True)
    s, m, v = tw.multisample_sample_conditional(Fg, layer.feature, layer.kern, layer.q_mu, q_sqrt=qg, white=False,
                                                eps=T64(eps).cuda(), jitter=layer.jitter)
    close('sample', s, so); close('mean', m, mo); close('var', v, vo)
    cs, cm, cv = [T64(rng.standard_normal(so.shape)) for _ in range(3)]
    (so * cs).sum().add((mo * cm).sum()).add((vo * cv).sum()).backward()
    (s * cs.cuda()).sum().add((m * cm.cuda()).sum()).add((v * cv.cuda()).sum()).backward()
    mix = hasattr(layer.kern, 'W')
    base = layer.kern.kernel if mix else layer.kern
    feat = layer.feature.feat if hasattr(layer.feature, 'feat') else layer.feature
    sig = lambda p: torch.sigmoid(p.unconstrained.detach())
    close('dF', Fg.grad, Fo.grad)
    close('dZ', feat.Z.unconstrained.grad, leaves[pre + 'Z'].grad)
    close('dq_mu', layer.q_mu.unconstrained.grad, leaves[pre + 'q_mu'].grad)
    close('dls', base.lengthscales.unconstrained.grad / sig(base.lengthscales), leaves[pre + 'kern.lengthscales'].grad)
    close('dvariance', base.variance.unconstrained.grad / sig(base.variance), leaves[pre + 'kern.variance'].grad)
    if form == 'full':
        close('dq_sqrt', layer.q_sqrt.unconstrained.grad, torch.tril(leaves[pre + 'q_sqrt'].grad))
    elif form == 'diag':
        close('dq_sqrt', qg.grad, qo.grad)


def test_model_propagate_matches_oracle_layer_loop():
    """DGP_VI.propagate (reference models.py:30-46): (samples, means, covs, kls, kl_types), one entry per layer, on the
    IW layout [N, K, .] with the sampled local regulariser and on the prediction layout (no amortisation inputs),
    against oracle.DGP.propagate; gradient of a scalar of the outputs w.r.t. a first-layer parameter through autograd."""
    from dgps_with_iwvi_b200.build_models import model_from_spec
    from dgps_with_iwvi_b200.layers import RegularizerType
    N, D, M, K = 9, 3, 20, 4
    X, Y = S.make_data(40, D, seed=13)
    spec = S.make_spec(X, 'L1_G3_G2', M, K, seed=13, perturb=0.3, inner_q_sqrt_scale=0.3)
    model = model_from_spec(spec, X, Y)
    omodel, leaves = O.build_from_spec(spec, requires_grad=True)
    eps = S.make_noise(spec, (N, K), seed=14, final_noise=True)
    Xt = np.repeat(X[:N, None, :], K, 1); XY = np.concatenate([Xt, np.repeat(Y[:N, None, :], K, 1)], -1)
    outs_o = omodel.propagate(T64(Xt), [T64(e) for e in eps], inference_amorization_inputs=T64(XY),
                              is_sampled_local_regularizer=True)
    model.requires_grad_()
    outs = model.propagate(Xt, inference_amorization_inputs=T64(XY).cuda(), is_sampled_local_regularizer=True,
                           eps=[T64(e).cuda() for e in eps])
    assert len(outs) == 5 and [len(o) for o in outs] == [4] * 5
    for name, got, want in zip(['samples', 'means', 'covs', 'kls'], outs[:4], outs_o[:4]):
        for i, (a, b) in enumerate(zip(got, want)):
            close('%s[%d]' % (name, i), a, b)
    assert outs[4] == [RegularizerType.LOCAL] + [RegularizerType.GLOBAL] * 3
    assert [t == O.LOCAL for t in outs_o[4]] == [t is RegularizerType.LOCAL for t in outs[4]]
    c = T64(np.random.default_rng(15).standard_normal(outs_o[1][-1].shape))
    (outs_o[1][-1] * c).sum().add(outs_o[0][1].sum()).backward()
    (outs[1][-1] * c.cuda()).sum().add(outs[0][1].sum()).backward()
    close('dW0', model.layers[0].encoder.Ws[0].unconstrained.grad, leaves['layers.0.encoder.Ws.0'].grad)
    close('dZ1', model.layers[1].feature.feat.Z.unconstrained.grad, leaves['layers.1.Z'].grad)
    # prediction layout [S, N, .]: LV layers sample from the prior (layers.py:73-81), closed-form KL
    Sn = 5
    eps_p = S.make_noise(spec, (Sn, N), seed=16, final_noise=True)
    Xp = np.repeat(X[None, :N, :], Sn, 0)
    with torch.no_grad():
        po = omodel.propagate(T64(Xp), [T64(e) for e in eps_p])
        pg = model.propagate(Xp, eps=[T64(e).cuda() for e in eps_p])
    for name, got, want in zip(['samples', 'means', 'covs', 'kls'], pg[:4], po[:4]):
        for i, (a, b) in enumerate(zip(got, want)):
            close('predict %s[%d]' % (name, i), a, b)


@pytest.mark.parametrize('act', ['relu', 'sigmoid', 'softplus', 'elu', 'identity'])
def test_encoder_activation_functions(act):
    """Encoder(activation_func=...) (reference layers.py:109,122,144): the non-default non-linearities of the fused
    encoder kernel -- layer-level values and gradients, and the whole-model IW-ELBO + gradients through the engine."""
    from dgps_with_iwvi_b200.build_models import model_from_spec
    import helpers as H
    D, N, K = 3, 12, 4
    X, Y = S.make_data(30, D, seed=5)
    spec = S.make_spec(X, 'L2_G2', 10, K, seed=5, perturb=0.3, inner_q_sqrt_scale=0.3)
    spec['layers'][0]['activation'] = act
    model = model_from_spec(spec, X, Y)
    layer = model.layers[0]
    assert layer.encoder.activation_func == act
    omodel, leaves = O.build_from_spec(spec, requires_grad=True)
    olayer = omodel.layers[0]
    rng = np.random.default_rng(6)
    F = rng.standard_normal((N, K, D)); XY = rng.standard_normal((N, K, D + 1)); eps = rng.standard_normal((N, K, 2))
    Fo = T64(F).requires_grad_(True)
    outs_o = olayer.propagate(Fo, T64(XY), True, eps=T64(eps))
    Fg = T64(F).cuda().requires_grad_(True)
    for _, p in layer.named_parameters():
        p.unconstrained.requires_grad_(True)
    outs = layer.propagate(Fg, inference_amorization_inputs=T64(XY).cuda(), is_sampled_local_regularizer=True,
                           eps=T64(eps).cuda())
    cots = [T64(rng.standard_normal(o.shape)) for o in outs_o]
    for n, a, b in zip(['samples', 'mean', 'cov', 'kl'], outs, outs_o):
        close(n, a, b)
    sum((o * c).sum() for o, c in zip(outs_o, cots)).backward()
    sum((o * c.cuda()).sum() for o, c in zip(outs, cots)).backward()
    for j, (W, b) in enumerate(zip(layer.encoder.Ws, layer.encoder.bs)):
        close('dW%d' % j, W.unconstrained.grad, leaves['layers.0.encoder.Ws.%d' % j].grad)
        close('db%d' % j, b.unconstrained.grad, leaves['layers.0.encoder.bs.%d' % j].grad)
    # whole model through the engine
    m2 = model_from_spec(spec, X, Y)
    eps_m = S.make_noise(spec, (N, K), seed=7)
    e_ref, g_ref = O.iw_elbo_and_grads(spec, X[:N], Y[:N], eps_m)
    e, g = m2.compute_log_likelihood_and_grads(X[:N], Y[:N], eps_m)
    assert abs(e - e_ref.item()) < RTOL * abs(e_ref.item())
    H.assert_grads_close(g, {k: v.numpy() for k, v in g_ref.items()}, RTOL, act)
    with pytest.raises(NotImplementedError):
        from dgps_with_iwvi_b200.layers import Encoder
        Encoder(1, 3, [4], activation_func='swish')


@pytest.mark.parametrize('ard', [True, False])
def test_kernel_active_dims(ard):
    """gpflow kernels restricted by `active_dims` (Kern._slice): a Mok layer whose base kernel acts on 3 of its 5 input
    columns, ARD and shared lengthscale -- the operator-level conditional (values, gradients) and the whole-model
    IW-ELBO + gradients through the engine, against the oracle evaluating the sliced kernel."""
    from dgps_with_iwvi_b200 import temp_workaround as tw
    from dgps_with_iwvi_b200.build_models import model_from_spec
    import helpers as H
    D, N, K, M = 4, 14, 3, 17
    X, Y = S.make_data(40, D, seed=9)
    spec = S.make_spec(X, 'L1_G3', M, K, seed=9, perturb=0.3, inner_q_sqrt_scale=0.3)
    g1 = spec['layers'][1]                                   # input width D + 1 = 5
    g1['active_dims'] = [0, 2, 4]
    g1['lengthscales'] = g1['lengthscales'][[0, 2, 4]] if ard else np.array(1.7)
    model = model_from_spec(spec, X, Y)
    layer = model.layers[1]
    assert layer.kern.kernel.active_dims == [0, 2, 4] and layer.kern.kernel.ARD == ard
    omodel, leaves = O.build_from_spec(spec, requires_grad=True)
    ol = omodel.layers[1]
    rng = np.random.default_rng(2)
    F = rng.standard_normal((N, K, D + 1)); eps = rng.standard_normal((N, K, 3))
    Fo = T64(F).requires_grad_(True)
    so, mo, vo = O.multisample_sample_conditional(Fo, ol.Z, ol.kern, ol.q_mu, q_sqrt=ol.q_sqrt, white=True, eps=T64(eps))
    Fg = T64(F).cuda().requires_grad_(True)
    model.requires_grad_()
    s, m, v = tw.multisample_sample_conditional(Fg, layer.feature, layer.kern, layer.q_mu, q_sqrt=layer.q_sqrt, white=True,
                                                eps=T64(eps).cuda(), jitter=layer.jitter)
    close('sample', s, so); close('mean', m, mo); close('var', v, vo)
    c = T64(rng.standard_normal(so.shape))
    ((so * c).sum() + (vo * c).sum()).backward()
    ((s * c.cuda()).sum() + (v * c.cuda()).sum()).backward()
    close('dF', Fg.grad, Fo.grad)
    base = layer.kern.kernel
    sig = torch.sigmoid(base.lengthscales.unconstrained.detach())
    close('dls', (base.lengthscales.unconstrained.grad / sig).reshape(leaves['layers.1.kern.lengthscales'].shape),
          leaves['layers.1.kern.lengthscales'].grad)
    close('dZ', layer.feature.feat.Z.unconstrained.grad, leaves['layers.1.Z'].grad)
    assert (layer.feature.feat.Z.unconstrained.grad[:, [1, 3]] == 0).all()      # inactive columns: exact zeros
    m2 = model_from_spec(spec, X, Y)
    eps_m = S.make_noise(spec, (N, K), seed=3)
    e_ref, g_ref = O.iw_elbo_and_grads(spec, X[:N], Y[:N], eps_m)
    e, g = m2.compute_log_likelihood_and_grads(X[:N], Y[:N], eps_m)
    assert abs(e - e_ref.item()) < RTOL * abs(e_ref.item())
    H.assert_grads_close(g, {k: v.numpy() for k, v in g_ref.items()}, RTOL, 'active_dims')
