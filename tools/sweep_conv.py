#!/usr/bin/env python
"""Every distinct convolution launch shape of one M1 training step (shape-only trace, tools/list_launches.py),
timed in isolation on the GPU for each variant of the tcgen05 engines:

  forward launch        tune[0] = 1 (one TMA box per tap), 2 (halo tile, if plannable), 3 (multi-tile, experimental)
  weight gradient       default tilings vs SHIFT mode (tune[1] = 2, if plannable)


CUDA events, L2 flushed between iterations; prints ms per launch, launches per step and the per-step total of
the best variant. Usage: sweep_conv.py [--iters N] [--batch B] [--min-ms X] [--what fwd,wgrad]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import m1b200  # noqa: E402,F401
from m1b200 import _lib, ops  # noqa: E402
import list_launches as LL  # noqa: E402


def timeit(fn, iters, flushbuf):
    ts = []
    for _ in range(iters + 2):
        flushbuf.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts[2:]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=3)
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--what', default='fwd,wgrad')
    ap.add_argument('--min-gflop', type=float, default=0.0, help='skip launches below this many GFLOP')
    args = ap.parse_args()
    what = args.what.split(',')
    ctx = _lib.Context.get(0)
    dev = 'cuda'
    flushbuf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    bt = torch.bfloat16
    # distinct launch shapes and how often each occurs in a step
    shapes = {}
    for r in LL.trace(args.batch):
        if r['out_fp32']:
            continue
        key = (r['transposed'], r['in_dhw'], r['out_dhw'], tuple(r['kernel']), tuple(r['stride']), tuple(r['pad']),
               tuple(r['src_c']), tuple(r['out_c']))
        ent = shapes.setdefault(key, [r, 0])
        ent[1] += 1
    tot = {'fwd': [0.0, 0.0], 'wgrad': [0.0, 0.0]}          # [baseline variant, best variant] ms per step
    print('%-40s %3s | %-36s | %s' % ('launch', 'n', 'forward ms: per-tap / halo / multi', 'wgrad ms: default / shift'))
    for key, (r, count) in sorted(shapes.items(), key=lambda kv: -kv[1][0]['flops'] * kv[1][1]):
        if r['flops'] / 1e9 < args.min_gflop:
            continue
        nb = r['batch']
        xs = [torch.randn((nb,) + r['in_dhw'] + (c,), device=dev).to(bt) for c in r['src_c']]
        outs = [torch.empty((nb,) + r['out_dhw'] + (c,), device=dev, dtype=bt) for c in r['out_c']]
        cin = sum(r['src_c'])
        k = tuple(r['kernel'])
        if r['transposed']:
            ws = [torch.randn(k + (r['out_c'][0], cin), device=dev) * 0.05]
        else:
            ws = [torch.randn(k + (cin, co), device=dev) * 0.05 for co in r['out_c']]
        name = ('+'.join(n.split('/', 1)[-1] for n in r['names']) + ('(T)' if r['transposed'] else ''))[:40]
        line = '%-40s %3d | ' % (name, count)
        if 'fwd' in what:
            d = LL.fwd_desc(r)
            cols = []
            if ops.conv3d_tc_supported(d):
                packed = ops.conv3d_pack_weights(ctx, d, ws)
                for var in (1, 2, 3):
                    d.tune[0] = var
                    if var == 2 and not ops.conv3d_halo_engine(d):
                        cols.append(None)
                        continue
                    cols.append(timeit(lambda: ops.conv3d(ctx, d, xs, ws, None, outs, packed), args.iters, flushbuf))
                ok = [c for c in cols if c is not None]
                tot['fwd'][0] += cols[0] * count
                tot['fwd'][1] += min(ok) * count
            line += '%-36s | ' % ' / '.join('%7.3f' % c if c is not None else '   -   ' for c in cols)
        if 'wgrad' in what and not r['transposed']:
            d = LL.fwd_desc(r)
            cols = []
            if ops.conv3d_wgrad_tc_supported(d):
                douts = [torch.randn_like(o) for o in outs]
                dws = [torch.zeros_like(w) for w in ws]
                for mode in (0, 2):
                    d.tune[0], d.tune[1], d.tune[2], d.tune[3] = (64, 0, 3, 1) if mode == 0 else (192, 2, 3, 1)
                    if mode == 2:
                        info = ops.conv3d_plan_info(d, 2)
                        if not info or not info[12]:
                            cols.append(None)
                            continue
                    cols.append(timeit(lambda: ops.conv3d_wgrad(ctx, d, xs, douts, dws, None), args.iters, flushbuf))
                ok = [c for c in cols if c is not None]
                tot['wgrad'][0] += cols[0] * count
                tot['wgrad'][1] += min(ok) * count
            line += ' / '.join('%7.3f' % c if c is not None else '   -   ' for c in cols)
        print(line, flush=True)
    for k_, (base, best) in tot.items():
        if base:
            print('%s per step: baseline variant %.2f ms, best variant per shape %.2f ms' % (k_, base, best))


if __name__ == '__main__':
    main()
