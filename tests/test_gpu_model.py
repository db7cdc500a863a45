"""Whole-model GPU parity through the reference-shaped API (DGP_IWVI / DGP_VI -> engine -> C ABI -> CUDA) against
(a) the committed golden vectors (generated from the oracle by tests/golden/make_golden.py) and (b) the oracle run
live on the same seeded inputs.  Tolerance: ELBO and gradients within rtol 1e-8 (north_star; float64)."""
import numpy as np
import pytest
import torch

import helpers as H
from oracle import iwvi_oracle as O
from oracle import synthetic as S

pytestmark = pytest.mark.gpu
RTOL = 1e-8


def _model(spec, X, Y, mode='iw'):
    from dgps_with_iwvi_b200.build_models import model_from_spec
    return model_from_spec(spec, X, Y, mode=mode)


@pytest.mark.parametrize('name', ['iw_c1_demo_shape', 'iw_L1_G5', 'iw_L1_G5_G5', 'iw_matern52_linear'])
def test_golden_iw_elbo_and_grads(name):
    spec, X, Y, eps, elbo, grads, extra = H.load_golden(name)
    m = _model(spec, X, Y)
    got_elbo, got = m.compute_log_likelihood_and_grads(X, Y, eps)
    assert abs(got_elbo - elbo) < RTOL * abs(elbo), (got_elbo, elbo)
    H.assert_grads_close(got, grads, RTOL, name)
    # forward-only entry point gives the same number
    assert abs(m.compute_log_likelihood(X, Y, eps) - elbo) < RTOL * abs(elbo)
    # the VI bound of the same model and noise (sample-major layout, models.py:50-53)
    K, N = spec['num_samples'], len(X)
    mv = _model(spec, X, Y, mode='vi')
    eps_vi = [None if e is None else np.transpose(e, (1, 0, 2)).reshape(K * N, -1) for e in eps]
    v = mv.compute_log_likelihood(X, Y, eps_vi)
    assert abs(v - extra['vi_elbo']) < RTOL * abs(extra['vi_elbo'])


@pytest.mark.parametrize('conf,N,D,M,K,kern', [('L1_G3_G2', 50, 5, 64, 6, 'RBF'), ('G3_L1_G2', 31, 3, 20, 5, 'RBF'),
                                               ('G2', 40, 2, 30, 4, 'Matern32'), ('L2_G3', 33, 4, 130, 9, 'Matern52'),
                                               ('L1_G5', 64, 8, 256, 50, 'RBF')])
def test_live_oracle_iw(conf, N, D, M, K, kern):
    X, Y = S.make_data(N, D, seed=11)
    spec = S.make_spec(X, conf, M, K, seed=11, perturb=0.3, inner_q_sqrt_scale=0.3, kern=kern)
    eps = S.make_noise(spec, (N, K), seed=12)
    e_ref, g_ref = O.iw_elbo_and_grads(spec, X, Y, eps, reference_style=True)
    m = _model(spec, X, Y)
    e, g = m.compute_log_likelihood_and_grads(X, Y, eps)
    assert abs(e - e_ref.item()) < RTOL * abs(e_ref.item())
    H.assert_grads_close(g, {k: v.numpy() for k, v in g_ref.items()}, RTOL, conf)


@pytest.mark.parametrize('conf,kern', [('L1_G3', 'RBF'), ('G2', 'Matern52')])
def test_live_oracle_vi_grads(conf, kern):
    N, D, M, K = 37, 3, 25, 4
    X, Y = S.make_data(N, D, seed=21)
    spec = S.make_spec(X, conf, M, K, seed=21, perturb=0.3, inner_q_sqrt_scale=0.3, kern=kern)
    eps = S.make_noise(spec, (K * N,), seed=22)
    e_ref, g_ref = O.vi_elbo_and_grads(spec, X, Y, eps)
    m = _model(spec, X, Y, mode='vi')
    e, g = m.compute_log_likelihood_and_grads(X, Y, eps)
    assert abs(e - e_ref.item()) < RTOL * abs(e_ref.item())
    H.assert_grads_close(g, {k: v.numpy() for k, v in g_ref.items()}, RTOL, conf)


def test_predict_matches_oracle():
    N, D, M, K, Sn = 20, 3, 25, 4, 6
    X, Y = S.make_data(N, D, seed=31)
    spec = S.make_spec(X, 'L1_G3', M, K, seed=31, perturb=0.3, inner_q_sqrt_scale=0.3)
    model, _ = O.build_from_spec(spec)
    eps = S.make_noise(spec, (Sn, N), seed=32)
    T = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float64)
    m_ref, v_ref = model.predict_f_multisample(T(X), Sn, [T(e) for e in eps])
    m = _model(spec, X, Y)
    mm, vv = m.predict_f_multisample(X, Sn, [None if e is None else e.reshape(Sn * N, -1) for e in eps])
    np.testing.assert_allclose(mm, m_ref.numpy(), rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(vv, v_ref.numpy(), rtol=1e-8, atol=1e-12)
    ys = m.predict_y_samples(X, Sn)
    assert ys.shape == (Sn, N, 1) and np.isfinite(ys).all()
    # models.py:99-107 with the likelihood noise injected: y = mean + z sqrt(var + likelihood variance), on the device
    eps_y = np.random.default_rng(33).standard_normal((Sn, N, 1))
    ys_ref = model.predict_y_samples(T(X), Sn, [T(e) for e in eps], T(eps_y))
    ys2 = m.predict_y_samples(X, Sn, [None if e is None else e.reshape(Sn * N, -1) for e in eps], eps_y)
    np.testing.assert_allclose(ys2, ys_ref.numpy(), rtol=1e-8, atol=1e-10)


def test_parameter_assignment_and_frozen_flags():
    """tests/test_gp_layer.py:46-47 style assignment reaches the device buffers; set_trainable reaches the mask."""
    from dgps_with_iwvi_b200.engine import FlatParams
    N, D, M, K = 16, 2, 10, 3
    X, Y = S.make_data(N, D, seed=41)
    spec = S.make_spec(X, 'L1_G2', M, K, seed=41)
    eps = S.make_noise(spec, (N, K), seed=42)
    m = _model(spec, X, Y)
    e0 = m.compute_log_likelihood(X, Y, eps)
    new_q_mu = np.random.default_rng(0).standard_normal((M, 1))
    m.layers[-1].q_mu = new_q_mu
    spec['layers'][-1]['q_mu'] = new_q_mu
    e1 = m.compute_log_likelihood(X, Y, eps)
    e_ref = O.iw_elbo_and_grads(spec, X, Y, eps)[0].item()
    assert e0 != e1 and abs(e1 - e_ref) < RTOL * abs(e_ref)
    np.testing.assert_array_equal(m.layers[-1].q_mu.read_value(), new_q_mu)
    m.layers[1].kern.W.set_trainable(False)
    flat = FlatParams.of(m)
    flat.refresh_mask()
    _, o, sz, _, _ = flat.entries[id(m.layers[1].kern.W)]
    assert flat.mask[o:o + sz].sum().item() == 0 and flat.mask.sum().item() > 0


def test_full_size_properties_c2():
    """BASELINE config 2 at full size (N=10k, D=8, L1_G5, M=100, K=20, B=512): properties that need no oracle run.
    (1) the ELBO equals scale * sum_n (logsumexp_k L_nk - log K) - sum KL recomputed from the per-point outputs;
    (2) softmax weights sum to one per row; (3) q_sqrt's strict upper triangle gets exactly zero gradient;
    (4) the same call twice is bit-identical (fixed-order reductions); (5) K=1 IW == VI with the same noise."""
    from dgps_with_iwvi_b200.engine import FlatParams
    c = S.CONFIGS['c2']
    X, Y = S.make_data(c['N'], c['D'], seed=0)
    spec = S.make_spec(X, c['configuration'], c['M'], c['K'], lik_variance=c['lik_variance'], seed=0)
    B, K = c['B'], c['K']
    m = _model(spec, X, Y)
    Xb, Yb = X[:B], Y[:B]
    e1, g1 = m.compute_log_likelihood_and_grads(Xb, Yb)
    eng = m.engine(B, K)
    w = eng.w.cpu().numpy()
    np.testing.assert_allclose(w.sum(1), 1.0, rtol=0, atol=1e-13)
    logp = eng.logp.cpu().numpy()
    kl = eng.kls[:eng.n_gp].cpu().numpy().sum()
    assert abs(e1 - (c['N'] / B * logp.sum() - kl)) < 1e-10 * abs(e1)
    for k, v in g1.items():
        assert np.isfinite(v).all(), k
        if k.endswith('q_sqrt'):
            assert np.all(np.triu(v, 1) == 0), k
    m._evals -= 1
    e2, g2 = m.compute_log_likelihood_and_grads(Xb, Yb)
    assert e1 == e2
    for k in g1:
        assert np.array_equal(g1[k], g2[k]), k
    # K = 1: importance weighting is vacuous, IW == VI on the same noise
    spec1 = dict(spec, num_samples=1)
    eps = S.make_noise(spec1, (B, 1), seed=5)
    a = _model(spec1, X, Y, 'iw').compute_log_likelihood(Xb, Yb, eps)
    # the VI bound uses the closed-form local KL, the IW bound its one-sample estimate: compare through the parts
    mv = _model(spec1, X, Y, 'vi')
    b = mv.compute_log_likelihood(Xb, Yb, [None if e is None else e.reshape(B, -1) for e in eps])
    engi, engv = m.engine(B, K), mv.engine(B, 1)
    assert np.isfinite(a) and np.isfinite(b) and abs(a - b) < 0.05 * abs(a)


def test_data_parallel_two_gpus_matches_single():
    """2 ranks over NCCL (row shards + one all-reduce of the flat bucket) == one rank on the whole minibatch."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29533', os.path.join(here, 'dp_check.py')],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_trainer_cuda_graph_matches_eager():
    """The captured CUDA graph of the training step (device-side step counter for noise seeds and Adam's bias correction)
    reproduces the eager step bit for bit over several steps, including fresh noise on every replay."""
    from dgps_with_iwvi_b200.build_models import build_model
    from dgps_with_iwvi_b200.engine import FlatParams
    from dgps_with_iwvi_b200.training import Trainer
    N, D, B, K = 600, 4, 48, 5
    X, Y = S.make_data(N, D, seed=3)
    out = []
    for use_graph in (False, True):
        model = build_model(X, Y, 'L1_G3', M=70, num_IW_samples=K, minibatch_size=B, mode='IWAE', seed=1)
        tr = Trainer(model, B, lr=1e-2, seed=5, use_graph=use_graph)
        losses = []
        for i in range(6):
            idx = (np.arange(B) + 7 * i) % N
            losses.append(tr.step(X[idx], Y[idx]))
        tr.engine.check_info()
        out.append((losses, FlatParams.of(model).x.clone(), int(tr.state[0].item())))
    (l0, x0, t0), (l1, x1, t1) = out
    assert t0 == t1 == 6
    assert len(set(l1)) == 6 and l0 == l1, (l0, l1)
    assert torch.equal(x0, x1)


def test_reference_iteration_natgrad_then_adam():
    """Row f3: one reference-style iteration (NatGrad on the last layer's q(u), then Adam on the rest, two minibatches,
    fresh noise each) against the oracle: the counter-based noise is regenerated with the numpy Philox restatement."""
    from dgps_with_iwvi_b200.engine import FlatParams, layer_seed
    from dgps_with_iwvi_b200.training import ReferenceIterationTrainer
    from oracle import natgrad_oracle as NO
    from oracle import philox_np
    N, D, M, K, B = 90, 3, 21, 4, 30
    X, Y = S.make_data(N, D, seed=61)
    spec = S.make_spec(X, 'L1_G2', M, K, seed=61, perturb=0.3, inner_q_sqrt_scale=0.3)
    m = _model(spec, X, Y)
    tr = ReferenceIterationTrainer(m, B, lr=1e-2, gamma=0.3, seed=17)
    X1, Y1, X2, Y2 = X[:B], Y[:B], X[B:2 * B], Y[B:2 * B]
    e_ng, e_adam = tr.iteration(X1, Y1, X2, Y2)
    tr.engine.check_info()

    def noise(step):
        out = []
        for li, ls in enumerate(spec['layers']):
            C = ls['latent_dim'] if ls['type'] == 'lv' else (ls['q_mu'].shape[1] if li < len(spec['layers']) - 1 else 0)
            out.append(None if C == 0 else philox_np.normal(B * K, C, 0, layer_seed(17, step, li)).reshape(B, K, C))
        return out
    # NatGrad half
    e1, g1 = O.iw_elbo_and_grads(spec, X1, Y1, noise(ReferenceIterationTrainer.NG_STEP_BASE))
    assert abs(e_ng.item() - e1.item()) < RTOL * abs(e1.item())
    last = len(spec['layers']) - 1
    mu_new, L_new = NO.natgrad_step(spec['layers'][last]['q_mu'], spec['layers'][last]['q_sqrt'],
                                    g1['layers.%d.q_mu' % last], g1['layers.%d.q_sqrt' % last], 0.3)
    np.testing.assert_allclose(m.layers[last].q_mu.read_value(), mu_new.numpy(), rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(np.tril(m.layers[last].q_sqrt.read_value()), L_new.numpy(), rtol=1e-8, atol=1e-10)
    # Adam half: evaluated at the updated q(u), on the second minibatch, with the noise of optimiser step 1
    spec2 = dict(spec, layers=[dict(l) for l in spec['layers']])
    spec2['layers'][last]['q_mu'], spec2['layers'][last]['q_sqrt'] = mu_new.numpy(), L_new.numpy()
    e2, g2 = O.iw_elbo_and_grads(spec2, X2, Y2, noise(1))
    assert abs(e_adam.item() - e2.item()) < RTOL * abs(e2.item())
    got = {k: v.cpu().numpy() for k, v in FlatParams.of(m).grads_by_name().items()}
    H.assert_grads_close(got, {k: v.numpy() for k, v in g2.items()}, RTOL, 'adam half')
    # the NatGrad-owned parameters are frozen for Adam; everything else moved
    z0 = spec['layers'][1]['Z']
    assert np.abs(m.layers[1].feature.feat.Z.read_value() - z0).max() > 1e-4
    np.testing.assert_allclose(m.layers[last].q_mu.read_value(), mu_new.numpy(), rtol=1e-8, atol=1e-10)


def test_full_size_properties_c3():
    """BASELINE config 3 at full per-GPU size (L1_G5_G5, D=16, M=256, K=50, B=512: 25 600 points per layer), properties
    that need no oracle run: bit-identical repeat (fixed-order reductions, side-stream overlap included), ELBO
    reassembled from the per-row outputs, exact zeros above q_sqrt's diagonal, finite gradients, and the training step
    (CUDA graph) moving the bound."""
    from dgps_with_iwvi_b200.build_models import build_model
    from dgps_with_iwvi_b200.training import Trainer
    c = S.CONFIGS['c3']
    N = 20000
    X, Y = S.make_data(N, c['D'], seed=0)
    m = build_model(X, Y, c['configuration'], M=c['M'], num_IW_samples=c['K'], minibatch_size=c['B'],
                    likelihood_variance=c['lik_variance'], mode='IWAE', seed=0)
    B, K = c['B'], c['K']
    Xb, Yb = X[:B], Y[:B]
    e1, g1 = m.compute_log_likelihood_and_grads(Xb, Yb)
    eng = m.engine(B, K)
    np.testing.assert_allclose(eng.w.cpu().numpy().sum(1), 1.0, rtol=0, atol=1e-13)
    kl = eng.kls[:eng.n_gp].cpu().numpy().sum()
    assert abs(e1 - (N / B * eng.logp.cpu().numpy().sum() - kl)) < 1e-10 * abs(e1)
    for k, v in g1.items():
        assert np.isfinite(v).all(), k
        if k.endswith('q_sqrt'):
            assert np.all(np.triu(v, 1) == 0), k
    m._evals -= 1
    e2, g2 = m.compute_log_likelihood_and_grads(Xb, Yb)
    assert e1 == e2 and all(np.array_equal(g1[k], g2[k]) for k in g1)
    # the forward pass and the per-point half of the backward pass run the full waves of tiles and the remainder as two
    # chains on two streams: same kernels on the same points -- the ELBO (forward only) is bit-identical to the single-chain
    # pass, per-point outputs too; the parameter gradients group the per-CTA partials differently (148 + 104 slots
    # instead of 148), so they agree to summation order
    from dgps_with_iwvi_b200.engine import Engine, FlatParams
    flat = FlatParams.of(m)
    assert eng.split == 296 * 64 and eng.split_bwd == 296 * 64
    eng.elbo_and_grads(Xb, Yb, None, seed=11, step=2)
    g_split = flat.g.clone()
    dX_split = eng.recs[1]['dX'].clone()
    one = Engine(m, B, K, 'iw', split_waves=False)
    assert one.split is None and one.split_bwd is None
    one.elbo_and_grads(Xb, Yb, None, seed=11, step=2)
    assert flat.g[flat.n].item() == g_split[flat.n].item()
    assert torch.equal(one.recs[1]['dX'], dX_split)
    for name, off, size, _, _ in flat.entries.values():
        a, b = g_split[off:off + size], flat.g[off:off + size]
        assert (a - b).abs().max().item() <= 1e-12 * max(b.abs().max().item(), 1e-300), name
    del one
    tr = Trainer(m, B, lr=1e-2, seed=3)
    losses = [tr.step(X[i * B:(i + 1) * B], Y[i * B:(i + 1) * B]) for i in range(12)]
    tr.engine.check_info()
    assert np.isfinite(losses).all() and np.mean(losses[-3:]) > np.mean(losses[:3])


def test_full_size_properties_c4_row_shards_add_up():
    """BASELINE config 4 at full size (L1_G5, D=8, M=512, K=256, B=4096: 1 048 576 points per layer, the largest shapes
    the build serves), size-independent properties: bit-identical repeat; the gradient bucket (last slot = ELBO) of the
    full minibatch equals the SUM of the buckets of its two row shards evaluated as ranks 0 / 1 of a world of 2 (what
    the single all-reduce of the data-parallel step adds up, models.py:144-150 with scale = num_data / B_global), and the
    per-row log-weights of the shards are those of the full batch -- noise is keyed by the global point index."""
    from dgps_with_iwvi_b200.build_models import build_model
    from dgps_with_iwvi_b200.engine import FlatParams
    c = S.CONFIGS['c4']
    X, Y = S.make_data(c['N'], c['D'], seed=0)
    m = build_model(X, Y, c['configuration'], M=c['M'], num_IW_samples=c['K'], minibatch_size=c['B'],
                    likelihood_variance=c['lik_variance'], mode='IWAE', seed=0)
    B, K = c['B'], c['K']
    Xb, Yb = X[:B], Y[:B]
    flat = FlatParams.of(m)
    eng = m.engine(B, K)
    eng.elbo_and_grads(Xb, Yb, None, seed=3, step=1, row0=0)
    eng.check_info()
    g_full, logp_full = flat.g.clone(), eng.logp.clone()
    assert torch.isfinite(g_full).all()
    eng.elbo_and_grads(Xb, Yb, None, seed=3, step=1, row0=0)
    assert torch.equal(flat.g, g_full) and torch.equal(eng.logp, logp_full)
    del eng
    m._engines.clear()
    torch.cuda.empty_cache()
    g_sum, logp = torch.zeros_like(g_full), []
    h = B // 2
    for r in range(2):
        e = m.engine(h, K, None, 2, r)
        e.elbo_and_grads(Xb[r * h:(r + 1) * h], Yb[r * h:(r + 1) * h], None, seed=3, step=1, row0=r * h)
        e.check_info()
        g_sum += flat.g
        logp.append(e.logp.clone())
    logp = torch.cat(logp)
    assert (logp - logp_full).abs().max().item() <= 1e-12 * logp_full.abs().max().item()
    for name, off, size, _, _ in flat.entries.values():
        a, b = g_sum[off:off + size], g_full[off:off + size]
        scale = b.abs().max().item()
        assert (a - b).abs().max().item() <= 1e-9 * max(scale, 1e-300), name
    assert abs(g_sum[flat.n].item() - g_full[flat.n].item()) <= 1e-11 * abs(g_full[flat.n].item())


def test_predict_two_chains_equals_single_chain():
    """Prediction (models.py:93-98, no saved panels) over enough points for the forward pass to split into two point
    chains: identical to the single-chain pass, and the ragged last tile (T not a multiple of the tile width) is served
    by the remainder chain."""
    from dgps_with_iwvi_b200.build_models import build_model
    from dgps_with_iwvi_b200.engine import Engine
    c = S.CONFIGS['c2']
    X, Y = S.make_data(2000, c['D'], seed=4)
    m = build_model(X, Y, c['configuration'], M=c['M'], num_IW_samples=c['K'], minibatch_size=c['B'],
                    likelihood_variance=c['lik_variance'], mode='IWAE', seed=0)
    N, Sn = 1003, 21                                    # 21 063 points
    two = Engine(m, N, Sn, 'predict')
    one = Engine(m, N, Sn, 'predict', split_waves=False)
    assert two.split is not None and one.split is None and (N * Sn) % 32
    outs = []
    for eng in (two, one):
        eng.set_batch(X[:N])
        eng.draw_noise(None, seed=5, step=1)
        mean, var = eng.forward()
        eng.check_info()
        outs.append((mean.clone(), var.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.isfinite(outs[0][0]).all() and (outs[0][1] > 0).all()
