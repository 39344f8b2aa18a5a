#!/usr/bin/env python
"""CPU study (oracle only, no GPU): which bf16 storage points cost how much accuracy. Runs the oracle's training
step with subsets of the bf16-storage emulation (oracle.m1_oracle.emulate_bf16_storage) against its fp32 run on
the configuration of the GPU parity test and prints softmax error, loss errors and gradient cosines.
  conv = raw convolution outputs that feed a norm / an elementwise kernel (could stay fp32 at +2 B/element)
  act  = every other stored activation (feeds a tensor-core convolution: has to be bf16)"""
import contextlib
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import m1_oracle as O  # noqa: E402

STRIDES = ((1, 1, 1), (1, 2, 2), (1, 2, 2), (2, 2, 2), (2, 2, 2))
KERNELS = ((1, 3, 3), (1, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3))


def step(cfg, x, y, emu):
    ps = O.ParamStore(dtype=torch.float32, seed=3, requires_grad=True)
    with torch.no_grad():
        O.train_loss(ps, cfg, x, y, O.Noise(0, torch.float32))
    g = torch.Generator().manual_seed(17)
    with torch.no_grad():
        for n, t in ps.p.items():
            if ps.kind[n] in ('gamma', 'beta', 'se_bias'):
                t.add_(0.2 * torch.randn(t.shape, generator=g).to(t.dtype))
    with (emu if emu is not None else contextlib.nullcontext()):
        r = O.train_loss(ps, cfg, x, y, O.Noise(5, torch.float32), alpha=(0.75, 0.25), gamma=2.0, kl_weight=10.0)
    (r['detection_loss'] + 10.0 * r['KL_loss']).backward()
    return ps, r


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.default_config(num_classes=2, dropout_rate=0.5, dropout_mode='monte-carlo', strides=STRIDES,
                           kernel_sizes=KERNELS, dense_skip=True, deep_supervision=True, probabilistic=True,
                           prob_latent_dims=(3, 2, 1, 0), filters=(32, 64, 128, 192, 256), se_reduction=(8,) * 5)
    x, y = O.synthetic_batch(2, (8, 32, 32), probabilistic=True, seed=11)
    x = x.bfloat16().float()
    ps_f, r_f = step(cfg, x, y, None)
    b = torch.cat([(t.grad if t.grad is not None else torch.zeros_like(t)).double().flatten() for t in ps_f.p.values()])
    variants = [
        ('product: values + gradients of conv and act, bf16 weights', dict(fwd=('conv', 'act'), bwd=('conv', 'act'), weights=True)),
        ('pre-norm conv outputs and their gradients in fp32', dict(fwd=('act',), bwd=('act',), weights=True)),
        ('all activation gradients in fp32', dict(fwd=('conv', 'act'), bwd=(), weights=True)),
        ('conv outputs fp32 + all gradients fp32', dict(fwd=('act',), bwd=(), weights=True)),
        ('only gradients rounded (values fp32)', dict(fwd=(), bwd=('conv', 'act'), weights=False)),
        ('only bf16 weights', dict(fwd=(), bwd=(), weights=True)),
        ('fp32 weights, everything else as the product', dict(fwd=('conv', 'act'), bwd=('conv', 'act'), weights=False)),
        ('fp16 values + fp16 weights, bf16 gradients', dict(fwd=('conv', 'act'), bwd=('conv', 'act'), weights=True,
                                                            value_dtype=torch.float16)),
    ]
    print('%-58s %-28s %-9s %-9s %-7s %-7s %-7s' % ('variant', 'softmax err mean/p99/max', 'focal rel', 'KL rel',
                                                    'cos', 'prior', 'post'))
    for name, kw in variants:
        ps, r = step(cfg, x, y, O.emulate_bf16_storage(**kw))
        e = (r['detection'].detach() - r_f['detection'].detach()).abs().flatten()
        p99 = e.kthvalue(int(0.99 * e.numel())).values.item()
        a = torch.cat([(t.grad if t.grad is not None else torch.zeros_like(t)).double().flatten() for t in ps.p.values()])
        cos = (a @ b).item() / (a.norm().item() * b.norm().item())
        per = {'prior': [], 'posterior': []}
        for n, t in ps.p.items():
            t2 = ps_f.p[n]
            if t.grad is None or t2.grad is None or n.endswith('bias') or n.split('/')[0] not in per:
                continue
            g1, g2 = t.grad.double().flatten(), t2.grad.double().flatten()
            if g1.norm() > 0 and g2.norm() > 0:
                per[n.split('/')[0]].append((g1 @ g2).item() / (g1.norm().item() * g2.norm().item()))
        print('%-58s %.1e / %.1e / %.1e   %.1e   %.1e   %.4f  %.4f  %.4f' % (
            name, e.mean().item(), p99, e.max().item(),
            abs(r['detection_loss'].item() - r_f['detection_loss'].item()) / abs(r_f['detection_loss'].item()),
            abs(r['KL'].item() - r_f['KL'].item()) / abs(r_f['KL'].item()), cos,
            statistics.median(per['prior']), statistics.median(per['posterior'])), flush=True)


if __name__ == '__main__':
    main()
