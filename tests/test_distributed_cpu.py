"""Data-parallel host logic on CPU, world_size 2 over gloo (the GPU path uses the same code with NCCL):
row sharding of the global minibatch, GPU-count-invariant noise indexing, and the single all-reduce of the flat
float64 gradient bucket whose last slot carries the ELBO -- with the global KL entering exactly once."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

N, D, M, K, BG, CONF = 96, 3, 12, 4, 16, 'L1_G3_G2'


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_objective(spec, Xb, Yb, eps_b, B_global, world):
    """This rank's term of the sharded objective, as engine.Engine forms it: (num_data / B_global) * sum_n logp_n
    over the local rows minus (1/world) * sum_l KL_l; gradients by autograd on the oracle."""
    from oracle import iwvi_oracle as O
    model, leaves = O.build_from_spec(spec, requires_grad=True)
    T = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float64)
    _, parts = model.iw_likelihood(T(Xb), T(Yb), [T(e) for e in eps_b], reference_style=True, return_parts=True)
    kl_glob = sum(kl for kl in parts['kls'] if kl.dim() == 0)
    obj = parts['logp'].sum() * (float(spec['num_data']) / B_global) - kl_glob / world
    names = list(leaves)
    grads = torch.autograd.grad(obj, [leaves[n] for n in names], allow_unused=True)
    return obj.detach(), {n: (torch.zeros_like(leaves[n]) if g is None else g) for n, g in zip(names, grads)}


def _worker(rank, world, port, out):
    import helpers as H
    from dgps_with_iwvi_b200.build_models import model_from_spec
    from dgps_with_iwvi_b200.engine import FlatParams, layer_seed
    from dgps_with_iwvi_b200.models import Minibatch
    from dgps_with_iwvi_b200.training import shard_rows
    from oracle import philox_np
    from oracle import synthetic as S
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        X, Y = S.make_data(N, D, seed=3)
        spec = S.make_spec(X, CONF, M, K, seed=3, perturb=0.3, inner_q_sqrt_scale=0.3)
        idx = Minibatch(N, BG, seed=0).next()                 # the same stream on every rank
        r0, r1 = shard_rows(BG, world, rank)
        mine = idx[r0:r1]
        # (1) shards partition the global minibatch
        got = [torch.zeros(r1 - r0, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(got, torch.as_tensor(mine))
        assert np.array_equal(torch.cat(got).numpy(), idx)
        # (2) noise keyed by the global point index: element (point p, column c) of rank r equals element
        #     (r0*K + p, c) of the single-process draw (engine.draw_noise passes first_point = row0 * K)
        eps_full, eps_mine = [], []
        for li, ls in enumerate(spec['layers']):
            C_ = ls['latent_dim'] if ls['type'] == 'lv' else (ls['q_mu'].shape[1] if li < len(spec['layers']) - 1 else 0)
            if C_ == 0:
                eps_full.append(None); eps_mine.append(None); continue
            seed = layer_seed(7, 1, li)
            full = philox_np.normal(BG * K, C_, 0, seed)
            part = philox_np.normal((r1 - r0) * K, C_, r0 * K, seed)
            assert np.array_equal(part, full[r0 * K:r1 * K])
            eps_full.append(full.reshape(BG, K, C_)); eps_mine.append(part.reshape(r1 - r0, K, C_))
        # (3) one SUM all-reduce of the flat bucket reproduces the full-batch ELBO and gradients
        model = model_from_spec(spec, X, Y)
        flat = FlatParams.of(model)
        obj, grads = _shard_objective(spec, X[mine], Y[mine], eps_mine, BG, world)
        by_canon = {H.canon(flat.entries[id(p)][0]): p for p in flat.params}
        assert set(by_canon) == set(grads), set(by_canon) ^ set(grads)
        flat.g.zero_()
        for n, g in grads.items():
            flat.gview(by_canon[n]).copy_(g.reshape(flat.gview(by_canon[n]).shape))
        flat.loss_slot.copy_(obj.reshape(1))
        # the Trainer's exchange: packed bucket (trainable entries, lower triangles of q_sqrt, ELBO slot).  Poison what
        # must NOT travel: the strict upper triangle of q_sqrt's gradient and a frozen parameter's slot.
        from dgps_with_iwvi_b200.training import GradBucket
        frozen = model.layers[1].kern.W
        frozen.set_trainable(False)
        gb = GradBucket(flat, always_reduce=[])
        dense = sum(p.size for p in flat.params)
        tri = sum(l['q_mu'].shape[1] * M * (M - 1) // 2 for l in spec['layers'] if l['type'] == 'gp')
        assert gb.index.numel() == dense - tri - frozen.size + 1, (gb.index.numel(), dense, tri)
        qs = flat.gview(model.layers[1].q_sqrt)
        qs += torch.triu(torch.full_like(qs, 1e30), 1)
        w_before = (flat.gview(frozen) + 0).clone()
        gb.allreduce()
        assert torch.equal(flat.gview(frozen), w_before)            # untouched, not summed
        qs -= torch.triu(torch.full_like(qs, 1e30), 1)
        if rank == 0:
            from oracle import iwvi_oracle as O
            e_ref, g_ref = O.iw_elbo_and_grads(spec, X[idx], Y[idx], eps_full, reference_style=True)
            assert abs(flat.loss_slot.item() - e_ref.item()) < 1e-11 * abs(e_ref.item())
            H.assert_grads_close({n: v.numpy() for n, v in flat.grads_by_name().items()},
                                 {k: v.numpy() for k, v in g_ref.items() if k != 'layers.1.kern.W'}, 1e-10, 'dp2')
        out.put((rank, 'ok'))
    except Exception as e:   # noqa: BLE001 -- reported to the parent
        import traceback
        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_sharded_bucket_allreduce():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == 'ok', 'rank %d:\n%s' % (rank, msg)


def test_shard_rows_and_lr_schedule():
    from dgps_with_iwvi_b200.training import shard_rows, staircase_decay
    assert [shard_rows(4096, 8, r) for r in (0, 7)] == [(0, 512), (3584, 4096)]
    try:
        shard_rows(10, 4, 0)
        assert False
    except ValueError:
        pass
    assert staircase_decay(5e-3, 999) == 5e-3 and abs(staircase_decay(5e-3, 2000) - 5e-3 * 0.98 ** 2) < 1e-18
