#!/usr/bin/env python
"""Data-parallel correctness on real GPUs (run under torchrun, one rank per GPU):
R ranks x local batch b from identical weights == one process x batch R*b (same injected noise):
gradients after the bucketed NCCL all-reduce and weights after the Adam step must agree."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import m1b200  # noqa: E402,F401
from m1b200.model.distribute import init_from_env  # noqa: E402
from oracle import m1_oracle as O  # noqa: E402  (test infrastructure: noise + synthetic data only)
import test_model_gpu as T  # noqa: E402


def main():
    rank, local, world = init_from_env()
    dims, b = (8, 32, 32), 2
    arch = T.TINY
    # global batch and global noise, identical on every rank
    x, y = O.synthetic_batch(world * b, dims, seed=5)
    cfg = O.default_config(dropout_rate=0.5, dropout_mode='monte-carlo', strides=T.STRIDES, kernel_sizes=T.KERNELS,
                           dense_skip=True, deep_supervision=True, probabilistic=True, prob_latent_dims=(3, 2, 1, 0),
                           **arch)
    ps = O.ParamStore(dtype=torch.float32, seed=3)
    noise = O.Noise(7, torch.float32)
    with torch.no_grad():
        O.train_loss(ps, cfg, x, y, noise)            # materialises weights + the global noise tensors
    T._perturb(ps)
    weights = {n: t.numpy() for n, t in ps.p.items()}

    def make():
        m, *_ = T._build(arch, dims, b, 'fp32', True, True, True)
        m.set_weights(weights)
        return m

    # single-process reference on the full batch (every rank computes it; cheap): gradients, then weights
    ref = make()
    ref.set_noise(noise.t)
    ref.train_step(x, y, apply_update=False)
    torch.cuda.synchronize()
    g_ref = torch.cat([t.flatten() for t in ref.gradients().values()]).double()
    ref.set_noise(noise.t)
    ref.train_step(x, y)
    torch.cuda.synchronize()
    w_ref = ref.get_weights()

    m = make()
    m.distribute(bucket_bytes=256 << 10)
    sl = slice(rank * b, (rank + 1) * b)
    loc = {k: v[sl] for k, v in noise.t.items()}
    m.set_noise(loc)
    m.train_step(x[sl], y[sl], apply_update=False)
    torch.cuda.synchronize()
    g = torch.cat([t.flatten() for t in m.gradients().values()]).double()
    rel = float((g - g_ref).norm() / g_ref.norm())
    cos = float((g @ g_ref) / (g.norm() * g_ref.norm()))
    m.set_noise(loc)
    m.train_step(x[sl], y[sl])
    torch.cuda.synchronize()
    w = m.get_weights()
    worst = max(float(abs(torch.from_numpy(w[n]) - torch.from_numpy(w_ref[n])).max()) for n in w)
    # all ranks hold the same weights
    flat = m.params.w.clone()
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    spread = float((hi - lo).abs().max())
    print(f'rank {rank}/{world}: buckets {len(m.grad_sync.buckets)} fired-in-backward '
          f'{len(m.grad_sync.fired)}  all-reduced grad vs single-process: rel-l2 {rel:.2e} cos {cos:.8f}  '
          f'max|w_dp - w_single| = {worst:.3e}  replica spread = {spread:.3e}', flush=True)
    assert spread == 0.0, 'replicas diverged'
    assert rel < 1e-3 and cos > 0.999999, (rel, cos)
    assert worst <= 2.1e-3, worst   # Adam turns rounding noise of ~0 gradients into +-lr (see tests)
    dist.barrier()
    if rank == 0:
        print('DP OK')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
