"""The optimiser kernels that sit inside the timed training step (iwvi_adam_step / iwvi_adam_step_counter /
iwvi_positive_fwd, csrc/lv_elbo.cu) against oracle/adam_oracle.py -- the numpy restatement of
tf.train.AdamOptimizer(exponential_decay(lr, global_step, 1000, rate, staircase=True)) on GPflow's unconstrained
variables (reference experiments/build_models.py:289-300), itself pinned against torch.optim.Adam in
tests/test_oracle.py.  Tolerance: 1e-13 relative on the unconstrained buffer and both moment buffers after every
step (the update is a handful of flops per entry; nothing accumulates)."""
import numpy as np
import pytest
import torch

from oracle import adam_oracle as AO
from oracle import synthetic as S

pytestmark = pytest.mark.gpu


def _trainer(use_graph, lr, decay):
    from dgps_with_iwvi_b200.build_models import build_model
    from dgps_with_iwvi_b200.training import Trainer
    N, D, B, K = 400, 3, 32, 4
    X, Y = S.make_data(N, D, seed=8)
    model = build_model(X, Y, 'L1_G2', M=24, num_IW_samples=K, minibatch_size=B, mode='IWAE', seed=2)
    tr = Trainer(model, B, lr=lr, lr_decay=decay, seed=9, use_graph=use_graph)
    return X, Y, model, tr, B


@pytest.mark.parametrize('use_graph', [False, True])
def test_adam_kernel_matches_numpy_adam_across_decay_boundary(use_graph):
    """Six steps starting at global_step 997 -> 1003: the learning rate drops by `decay` AT step 1000 (the reference
    increments global_step before running the optimiser, build_models.py:297-300); steps 3.. replay the captured CUDA
    graph with the step count and learning rate read from device memory.  Frozen entries (kernel variance of the Mok
    layer, its W and Linear mean function: build_models.py:209,213,225-227) must not move and keep zero moments."""
    lr, decay = 1e-2, 0.5          # an exaggerated decay so that a one-step-late schedule cannot hide in the tolerance
    X, Y, model, tr, B = _trainer(use_graph, lr, decay)
    flat = tr.flat
    t0 = 997
    tr.t = t0
    tr.state[0] = t0
    assert flat.n_pos >= 3 and 0 < flat.mask.sum().item() < flat.n
    x = flat.x.cpu().numpy().copy()
    m, v = np.zeros_like(x), np.zeros_like(x)
    mask = flat.mask.cpu().numpy()
    lrs = []
    for i in range(6):
        idx = (np.arange(B) + 11 * i) % len(X)
        tr.step(X[idx], Y[idx])
        t = t0 + i + 1
        g = flat.g[:flat.n].cpu().numpy()          # dELBO/d(constrained) of THIS step (left in the bucket)
        lr_t = AO.staircase_decay(lr, t, 1000, decay)
        lrs.append(lr_t)
        x, m, v = AO.adam_step(x, g, m, v, t, lr_t, flat.n_pos, mask)
        np.testing.assert_allclose(flat.x.cpu().numpy(), x, rtol=1e-13, atol=1e-15, err_msg='x after step %d' % t)
        np.testing.assert_allclose(tr.m.cpu().numpy(), m, rtol=1e-13, atol=1e-300, err_msg='m after step %d' % t)
        np.testing.assert_allclose(tr.v.cpu().numpy(), v, rtol=1e-13, atol=1e-300, err_msg='v after step %d' % t)
        # the constrained copies the next forward pass reads: theta = softplus(x) + 1e-6 (gpflow Log1pe)
        np.testing.assert_allclose(flat.theta_pos[:flat.n_pos].cpu().numpy(), AO.positive_forward(x[:flat.n_pos]),
                                   rtol=1e-14, atol=0)
        assert int(tr.state[0].item()) == t
    assert lrs[:2] == [lr, lr] and lrs[2:] == [lr * decay] * 4
    frozen = mask == 0
    assert np.array_equal(flat.x.cpu().numpy()[frozen], x[frozen]) and not m[frozen].any() and not v[frozen].any()
    tr.engine.check_info()


def test_set_trainable_after_capture_is_honoured():
    """set_trainable(False) after the step graph has been captured: the mask is rebuilt, the stale graphs are dropped
    and the newly frozen parameter stops moving (reference: GPflow rebuilds var_list on compile)."""
    X, Y, model, tr, B = _trainer(True, 1e-2, 0.98)
    for i in range(4):
        tr.step(X[:B], Y[:B])
    assert tr._graphs is not None
    Z = model.layers[1].feature.feat.Z
    z0 = Z.read_value().copy()
    Z.set_trainable(False)
    tr.step(X[:B], Y[:B])
    tr.step(X[:B], Y[:B])
    assert np.array_equal(Z.read_value(), z0)
    q = model.layers[1].q_mu.read_value().copy()
    tr.step(X[:B], Y[:B])
    assert not np.array_equal(model.layers[1].q_mu.read_value(), q)


def test_positive_fwd_and_plain_adam_entry_points():
    """iwvi_positive_fwd and the host-scalar iwvi_adam_step (step count / learning rate passed by value)."""
    from dgps_with_iwvi_b200 import capi
    rng = np.random.default_rng(0)
    n, n_pos = 1001, 333
    x = rng.standard_normal(n) * 3
    x[:4] = [-40.0, 40.0, 0.0, -745.0]
    g = rng.standard_normal(n)
    mask = (rng.random(n) < 0.8).astype(np.float64)
    dev = torch.device('cuda')
    xd, gd, md, vd = (torch.as_tensor(a, device=dev) for a in (x.copy(), g, np.zeros(n), np.zeros(n)))
    maskd = torch.as_tensor(mask, device=dev)
    th = torch.zeros(n_pos, dtype=torch.float64, device=dev)
    capi.positive_fwd(xd, th, n_pos)
    np.testing.assert_allclose(th.cpu().numpy(), AO.positive_forward(x[:n_pos]), rtol=1e-14, atol=0)
    m, v = np.zeros(n), np.zeros(n)
    for t in (1, 2, 3):
        capi.adam_step(xd, gd, md, vd, maskd, th, n, n_pos, 3e-3, 0.9, 0.999, 1e-8, t)
        x, m, v = AO.adam_step(x, g, m, v, t, 3e-3, n_pos, mask)
        np.testing.assert_allclose(xd.cpu().numpy(), x, rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(th.cpu().numpy(), AO.positive_forward(x[:n_pos]), rtol=1e-14, atol=0)


def test_step_pipelined_matches_step():
    """Trainer.step_pipelined (staged H2D on a copy stream, ELBO handed back one call later) trains exactly like
    Trainer.step: same ELBO sequence, same parameters after six steps (eager steps, capture, replays)."""
    outs = []
    for pipelined in (False, True):
        X, Y, model, tr, B = _trainer(True, 1e-2, 0.98)
        Xh, Yh = torch.as_tensor(X).pin_memory(), torch.as_tensor(Y).pin_memory()
        elbos = []
        for i in range(6):
            xs, ys = Xh[i * B:(i + 1) * B].clone().pin_memory(), Yh[i * B:(i + 1) * B].clone().pin_memory()
            if pipelined:
                v = tr.step_pipelined(xs, ys)
                if v is not None:
                    elbos.append(v)
            else:
                elbos.append(tr.step(xs, ys))
        if pipelined:
            elbos.append(tr.flush())
        outs.append((elbos, tr.flat.x.cpu().numpy().copy()))
    assert len(outs[0][0]) == len(outs[1][0]) == 6
    np.testing.assert_array_equal(np.array(outs[0][0]), np.array(outs[1][0]))
    np.testing.assert_array_equal(outs[0][1], outs[1][1])


def test_pipelined_graph_trainer_matches_plain_trainer_at_c3_size():
    """BASELINE's headline shape (c3: L1_G5_G5, M=256, K=50, 512 rows): the pipelined, graph-captured Trainer (segment-wise
    Adam, next-step Cholesky factorisations on high-priority streams behind the backward pass, two point chains, the
    first layer's reductions in two ordered halves) trains BIT-identically to the plain eager Trainer over seven steps.
    Any cross-stream hazard at full size shows up here as a differing parameter (the small-shape tests cannot see one:
    their kernels are too short to overlap)."""
    import bench
    from dgps_with_iwvi_b200.build_models import build_model
    from dgps_with_iwvi_b200.engine import FlatParams
    from dgps_with_iwvi_b200.training import Trainer
    cfg = bench.CONFIGS['c3']
    X, Y = bench.make_data(20000, cfg['D'], seed=0)
    B = cfg['B']

    def run(pipeline, graph, steps=7):
        model = build_model(X, Y, cfg['configuration'], M=cfg['M'], num_IW_samples=cfg['K'], minibatch_size=B,
                            likelihood_variance=cfg['lik_variance'], mode='IWAE', seed=0)
        tr = Trainer(model, B, lr=5e-3, seed=3, use_graph=graph, pipeline=pipeline)
        losses = []
        for i in range(steps):
            idx = (torch.arange(i * B, (i + 1) * B, device=model.X.device) * 7919) % len(X)
            losses.append(tr.step_device(model.X[idx], model.Y[idx]).clone())
        torch.cuda.synchronize()
        tr.engine.check_info()
        return FlatParams.of(model).x.clone(), torch.cat(losses).cpu().numpy()

    x0, l0 = run(False, False)
    for _ in range(2):
        x1, l1 = run(True, True)
        assert np.isfinite(l1).all()
        np.testing.assert_array_equal(l0, l1)
        assert torch.equal(x0, x1), int((x0 != x1).sum())
