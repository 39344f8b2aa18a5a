#!/usr/bin/env python
"""Isolated timing of the bandwidth-bound kernels (K4/K5/K6/K8/K9) at the benchmark shapes: achieved
GB/s against the algorithmic bytes (tensors read + written once per pass that needs them)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import m1b200  # noqa: E402,F401
from m1b200 import _lib, ops  # noqa: E402

B = 8
dev = 'cuda'
bt = torch.bfloat16


def timeit(fn, flushbuf, iters=5):
    ts = []
    for _ in range(iters + 2):
        flushbuf.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts[2:]))


def main():
    ctx = _lib.Context.get(0)
    flushbuf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    shapes = {'res0x32': ((20, 160, 160), 32), 'res1x64': ((20, 80, 80), 64), 'res2x128': ((20, 40, 40), 128),
              'res3x256': ((10, 20, 20), 256), 'res0x16': ((20, 160, 160), 16)}
    only = sys.argv[1:] or list(shapes)
    for name in only:
        dhw, C = shapes[name]
        shp = (B,) + dhw + (C,)
        nel = int(np.prod(shp))
        x = torch.randn(shp, device=dev).to(bt)
        x4 = torch.randn(shp, device=dev).to(bt)
        dy = torch.randn(shp, device=dev).to(bt)
        y = torch.empty_like(x)
        dx = torch.empty_like(x)
        dx4 = torch.empty_like(x)
        gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        dgam, dbet = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        st = torch.empty((B, C, 2), device=dev)
        st4 = torch.empty((B, C, 2), device=dev)
        ops.inorm_stats(ctx, x, st)
        ops.inorm_stats(ctx, x4, st4)
        gate = torch.rand((B, C), device=dev)
        pool = torch.empty((B, C), device=dev)
        red = torch.empty((B, C, 5), device=dev)
        dgate = torch.empty((B, C), device=dev)
        dpool = torch.zeros((B, C), device=dev)
        drop = ops.make_dropout(0.5, None, 1, 2, mask=torch.zeros(nel // 8, dtype=torch.uint8, device=dev))
        gb = nel * 2 / 1e9          # one bf16 pass over the tensor
        rows = [
            ('inorm_stats', 1, lambda: ops.inorm_stats(ctx, x, st)),
            ('inorm_act_fwd', 2, lambda: ops.inorm_act_fwd(ctx, x, st, gam, bet, 0.1, y)),
            ('inorm_act_bwd', 5, lambda: ops.inorm_act_bwd(ctx, dy, x, st, gam, bet, 0.1, dx, False, dgam, dbet)),
            ('se_squeeze', 1, lambda: ops.se_squeeze(ctx, x, st, gam, bet, pool)),
            ('se_gate_fwd', 3, lambda: ops.se_gate_fwd(ctx, x, x4, st, st4, gam, bet, gam, bet, gate, drop, y)),
            ('se_gate_bwd_reduce', 3, lambda: ops.se_gate_bwd_reduce(ctx, dy, x, x4, st, st4, gam, bet, gam, bet, gate,
                                                                    drop, red, dgate)),
            ('se_gate_bwd_apply', 5, lambda: ops.se_gate_bwd_apply(ctx, dy, x, x4, st, st4, gam, bet, gam, bet, gate,
                                                                  drop, red, dpool, dx, dx4, dgam, dbet, dgam, dbet)),
        ]
        for kname, passes, fn in rows:
            t = timeit(fn, flushbuf)
            print('%-9s %-20s %7.3f ms  %7.1f GB/s (%d tensor passes, %.2f GB)' % (
                name, kname, t, passes * gb / (t / 1e3), passes, passes * gb), flush=True)
    n = 64_000_000
    w, g, m, v, vh = (torch.randn(n, device=dev) for _ in range(5))
    l2 = torch.zeros(1, device=dev)
    t = timeit(lambda: ops.adam_amsgrad(ctx, w, g, m, v, vh, 1e-3, 0.9, 0.999, 1e-7, 1e-4, 1.0, l2), flushbuf)
    print('adam_amsgrad 64M params   %7.3f ms  %7.1f GB/s' % (t, n * 36 / 1e9 / (t / 1e3)))


if __name__ == '__main__':
    main()
