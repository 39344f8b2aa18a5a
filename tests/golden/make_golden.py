#!/usr/bin/env python
"""Generates tests/golden/*.npz — ORACLE-GENERATED golden vectors (fp64 CPU oracle, fixed seeds), NOT
TensorFlow-generated: TF 2.5 cannot run in this image (see oracle/m1_oracle.py header). They pin the
oracle against regressions and give the CUDA path fixed targets that do not depend on the oracle code
at test time.   Usage: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import m1_oracle as O  # noqa: E402

STRIDES = ((1, 1, 1), (1, 2, 2), (1, 2, 2), (2, 2, 2), (2, 2, 2))
KERNELS = ((1, 3, 3), (1, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3))
TINY = dict(filters=(8, 16, 24, 32, 48), se_reduction=(4, 4, 4, 4, 4))


def perturb(ps, seed=17):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, t in ps.p.items():
            if ps.kind[n] in ('gamma', 'beta', 'se_bias'):
                t.add_(0.2 * torch.randn(t.shape, generator=g).to(t.dtype))


def conv_cases():
    g = torch.Generator().manual_seed(0)
    out = {}
    for i, (k, s) in enumerate([((1, 3, 3), (1, 1, 1)), ((3, 3, 3), (2, 2, 2)), ((3, 3, 3), (1, 2, 2))]):
        x = torch.randn((1, 4, 6, 8, 3), generator=g, dtype=torch.float64)
        w = torch.randn((*k, 3, 5), generator=g, dtype=torch.float64) * 0.3
        b = torch.randn((5,), generator=g, dtype=torch.float64) * 0.1
        y = O.conv3d_same(x, w, b, s)
        wt = torch.randn((*k, 4, 5), generator=g, dtype=torch.float64) * 0.3     # ConvT: (k, Cout=4, Cin=5)
        yt = O.conv3d_transpose_same(y, wt, None, s)
        out.update({f'c{i}_x': x.numpy(), f'c{i}_w': w.numpy(), f'c{i}_b': b.numpy(), f'c{i}_y': y.numpy(),
                    f'c{i}_wt': wt.numpy(), f'c{i}_yt': yt.numpy(), f'c{i}_k': np.array(k), f'c{i}_s': np.array(s)})
    return out


def model_case(probabilistic):
    cfg = O.default_config(num_classes=2, dropout_rate=0.5, dropout_mode='monte-carlo', strides=STRIDES,
                           kernel_sizes=KERNELS, dense_skip=True, deep_supervision=True,
                           probabilistic=probabilistic, prob_latent_dims=(3, 2, 1, 0), **TINY)
    dims = (8, 32, 32)
    x, y = O.synthetic_batch(1, dims, probabilistic=probabilistic, seed=21)
    ps = O.ParamStore(dtype=torch.float64, seed=4, requires_grad=True)
    with torch.no_grad():
        O.train_loss(ps, cfg, x.double(), y.double(), O.Noise(0))
    perturb(ps)
    noise = O.Noise(6)
    r = O.train_loss(ps, cfg, x.double(), y.double(), noise, alpha=(0.75, 0.25), gamma=2.0, kl_weight=10.0)
    data = r['detection_loss'] + (10.0 * r['KL_loss'] if probabilistic else 0.0)
    data.backward()
    pick = sorted(ps.p)[::17]            # a deterministic subset of gradients (norms + one full small tensor)
    out = dict(detection=r['detection'].detach().numpy().astype(np.float32),
               focal=np.array(r['detection_loss'].item()), loss=np.array(r['loss'].item()),
               kl=np.array(r['KL'].item() if probabilistic else 0.0),
               grad_names=np.array(pick),
               grad_norms=np.array([ps.p[n].grad.norm().item() if ps.p[n].grad is not None else 0.0 for n in pick]),
               nparams=np.array(ps.num_params()))
    small = [n for n in sorted(ps.p) if n.endswith('norm1/gamma')][:3]
    for n in small:
        out['g:' + n] = ps.p[n].grad.numpy()
    return out


if __name__ == '__main__':
    np.savez_compressed(os.path.join(HERE, 'conv_same.npz'), **conv_cases())
    np.savez_compressed(os.path.join(HERE, 'm1_prob_tiny.npz'), **model_case(True))
    np.savez_compressed(os.path.join(HERE, 'm1_det_tiny.npz'), **model_case(False))
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
