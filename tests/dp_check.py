"""Launched by torchrun (one process per GPU): data-parallel IW-ELBO + gradients on row shards, ONE NCCL all-reduce of
the flat bucket, compared on rank 0 with the same global minibatch evaluated by a single-process plan.  Also checks
that counter-based noise makes the result independent of the number of GPUs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
    from dgps_with_iwvi_b200.build_models import build_model
    from dgps_with_iwvi_b200.engine import FlatParams
    from dgps_with_iwvi_b200.training import shard_rows
    N, D, M, K, Bg = 4096, 6, 96, 7, 64 * world
    rng = np.random.default_rng(0)
    X = rng.standard_normal((N, D)); Y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.standard_normal((N, 1))
    model = build_model(X, Y, 'L1_G4_G3', M=M, num_IW_samples=K, minibatch_size=Bg, mode='IWAE', seed=3)
    for i, l in enumerate(model.layers):        # non-degenerate variational parameters
        if hasattr(l, 'q_mu'):
            l.q_mu = 0.3 * np.random.default_rng(10 + i).standard_normal(l.q_mu.shape)
            l.q_sqrt = l.q_sqrt.read_value() * (1e5 if i < len(model.layers) - 1 else 1.0) * 0.4
    idx = np.random.default_rng(5).permutation(N)[:Bg]
    r0, r1 = shard_rows(Bg, world, rank)
    flat = FlatParams.of(model)
    eng = model.engine(r1 - r0, K, None, world, rank)
    eng.elbo_and_grads(X[idx[r0:r1]], Y[idx[r0:r1]], None, seed=11, step=1, row0=r0)
    dist.all_reduce(flat.g, op=dist.ReduceOp.SUM)
    g_dp = flat.g.clone()
    ok = True
    if rank == 0:
        eng1 = model.engine(Bg, K, None, 1, 0)
        eng1.elbo_and_grads(X[idx], Y[idx], None, seed=11, step=1, row0=0)
        g_1 = flat.g.clone()
        scale = g_1.abs().max().item()
        err = (g_dp - g_1).abs().max().item() / scale
        rel_elbo = abs(g_dp[flat.n].item() - g_1[flat.n].item()) / abs(g_1[flat.n].item())
        print('dp%d vs single: bucket max err / max = %.3e, elbo rel err = %.3e' % (world, err, rel_elbo), flush=True)
        ok = err < 1e-10 and rel_elbo < 1e-12
    dist.barrier()
    # ---- the Trainer's step: segmented all-reduce overlapped with the backward pass, captured into ONE CUDA graph (steps
    #      3+), against a single-process Trainer fed the same global minibatches; and the eager one-bucket exchange
    from dgps_with_iwvi_b200.training import Trainer
    import warnings
    B = (r1 - r0)
    finals = {}
    for mode in ('overlap_graph', 'one_bucket_two_graphs'):
        m2 = build_model(X, Y, 'L1_G4_G3', M=M, num_IW_samples=K, minibatch_size=Bg, mode='IWAE', seed=3)
        with warnings.catch_warnings():
            warnings.simplefilter('error')            # a refused NCCL capture must not pass silently here
            tr = Trainer(m2, B, lr=1e-2, seed=4, overlap_comm=(mode == 'overlap_graph'), graph_comm=(mode == 'overlap_graph'))
            assert tr.overlap_comm == (mode == 'overlap_graph')
            for i in range(6):
                gi_ = np.random.default_rng(100 + i).permutation(N)[:Bg]
                loss = tr.step_device(torch.as_tensor(X[gi_[r0:r1]]).cuda(), torch.as_tensor(Y[gi_[r0:r1]]).cuda())
        torch.cuda.synchronize()
        assert tr._graphs is not None and (tr._graphs[1] is None) == (mode == 'overlap_graph')
        finals[mode] = (FlatParams.of(m2).x.clone(), float(loss.item()))
        tr.engine.check_info()
        # replicas must agree bit for bit (the peer-memory exchange sums the ranks' slots in rank order on every rank; NCCL's
        # all-reduce is bitwise reproducible across ranks as well)
        xs = [torch.empty_like(finals[mode][0]) for _ in range(world)]
        dist.all_gather(xs, finals[mode][0])
        same = all(torch.equal(xs[0], xi) for xi in xs)
        if rank == 0:
            print('Trainer dp%d [%s]: exchange = %s, replicas bit-identical = %s'
                  % (world, mode, 'peer memory (iwvi_dp_push / iwvi_dp_reduce)' if tr.gbucket.p2p is not None
                     else 'NCCL (%s)' % tr.gbucket.p2p_error, same), flush=True)
        ok = ok and same
    if rank == 0:
        m1 = build_model(X, Y, 'L1_G4_G3', M=M, num_IW_samples=K, minibatch_size=Bg, mode='IWAE', seed=3)
        tr1 = Trainer(m1, Bg, lr=1e-2, seed=4, distributed=False)
        for i in range(6):
            gi_ = np.random.default_rng(100 + i).permutation(N)[:Bg]
            loss1 = tr1.step_device(torch.as_tensor(X[gi_]).cuda(), torch.as_tensor(Y[gi_]).cuda())
        x1 = FlatParams.of(m1).x
        for mode, (x2, l2) in finals.items():
            err = (x2 - x1).abs().max().item() / x1.abs().max().item()
            lerr = abs(l2 - float(loss1.item())) / abs(float(loss1.item()))
            print('Trainer dp%d [%s] vs single after 6 steps: x max err / max = %.3e, elbo rel err = %.3e'
                  % (world, mode, err, lerr), flush=True)
            ok = ok and err < 1e-8 and lerr < 1e-9
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
