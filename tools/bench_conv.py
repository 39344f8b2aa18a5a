#!/usr/bin/env python
"""Isolated timing of the dominant convolution launches (CUDA events, L2 flushed between iterations).
Usage: bench_conv.py [shape-name ...] [--iters N] [--what fwd,dgrad,wgrad]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import m1b200  # noqa: E402,F401
from m1b200 import _lib, ops  # noqa: E402

B = 8
SHAPES = {
    # name: (dhw, src channels, out channels, kernel, stride, transposed)
    'sersp2': ((20, 40, 40), [128, 128, 128, 128], [32, 128], (3, 3, 3), (1, 1, 1), False),
    'x_n128': ((20, 40, 40), [128, 128, 128, 128], [128], (3, 3, 3), (1, 1, 1), False),
    'x_n192': ((20, 40, 40), [128, 128, 128, 128], [64, 128], (3, 3, 3), (1, 1, 1), False),
    'x_n256': ((20, 40, 40), [128, 128, 128, 128], [256], (3, 3, 3), (1, 1, 1), False),
    'sersd2': ((20, 40, 40), [128, 128, 128], [32, 128], (3, 3, 3), (1, 1, 1), False),
    'sersp3': ((10, 20, 20), [256, 256, 256], [64, 256], (3, 3, 3), (1, 1, 1), False),
    'sersp1': ((20, 80, 80), [64] * 5, [16, 64], (1, 3, 3), (1, 1, 1), False),
    'sersp0': ((20, 160, 160), [32] * 6, [16, 32], (1, 3, 3), (1, 1, 1), False),
    'sersd0': ((20, 160, 160), [32] * 5, [16, 32], (1, 3, 3), (1, 1, 1), False),
    'sersd1': ((20, 80, 80), [64] * 4, [16, 64], (1, 3, 3), (1, 1, 1), False),
    'conve0': ((20, 160, 160), [16], [32], (1, 3, 3), (1, 1, 1), False),
    'conv2_r0': ((20, 160, 160), [16], [16], (3, 3, 3), (1, 1, 1), False),
    'conv2_r1': ((20, 80, 80), [16], [16], (3, 3, 3), (1, 1, 1), False),
    'conv2_r2': ((20, 40, 40), [32], [32], (3, 3, 3), (1, 1, 1), False),
    'conv3_r0': ((20, 160, 160), [16], [32], (1, 1, 1), (1, 1, 1), False),
    'serse2': ((20, 80, 80), [64], [32, 128], (3, 3, 3), (1, 2, 2), False),
    'convtd1': ((20, 40, 40), [128], [64], (3, 3, 3), (1, 2, 2), True),
    'convtd0': ((20, 80, 80), [64], [32], (1, 3, 3), (1, 2, 2), True),
    'convtd2': ((10, 20, 20), [256], [128], (3, 3, 3), (2, 2, 2), True),
    'att2_c1': ((20, 40, 40), [128], [128], (1, 1, 1), (1, 1, 1), False),
    'conv3_r2': ((20, 40, 40), [32], [128], (1, 1, 1), (1, 1, 1), False),
    'conv3_r1': ((20, 80, 80), [16], [64], (1, 1, 1), (1, 1, 1), False),
}


def flush(buf):
    buf.zero_()


def timeit(fn, iters, flushbuf):
    ts = []
    for _ in range(iters + 2):
        flush(flushbuf)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts[2:]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('names', nargs='*', default=list(SHAPES))
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--what', default='fwd,dgrad,wgrad')
    ap.add_argument('--batch', type=int, default=B)
    ap.add_argument('--variant', type=int, default=0, help='m1_conv_desc.tune[0]: 0 heuristic, 1 per-tap, 2 halo')
    args = ap.parse_args()
    what = args.what.split(',')
    ctx = _lib.Context.get(0)
    dev = 'cuda'
    flushbuf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    bt = torch.bfloat16
    for name in args.names:
        dhw, cins, couts, k, s, tr = SHAPES[name]
        nb = args.batch
        cin = sum(cins)
        taps = int(np.prod(k))
        if tr:
            out_dhw = tuple(d * st for d, st in zip(dhw, s))
            pad = [ops.same_pads(out_dhw[i], k[i], s[i])[1] for i in range(3)]
            mode, wstr = _lib.CONV_TRANSPOSED, [(couts[0] * cin, 1, cin)]
            wshape = [k + (couts[0], cin)]
            vox = nb * int(np.prod(dhw))
        else:
            geo = [ops.same_pads(dhw[i], k[i], s[i]) for i in range(3)]
            out_dhw = tuple(g[0] for g in geo)
            pad = [g[1] for g in geo]
            mode, wstr = _lib.CONV_FWD, [(cin * co, co, 1) for co in couts]
            wshape = [k + (cin, co) for co in couts]
            vox = nb * int(np.prod(out_dhw))
        flops = 2.0 * vox * taps * cin * sum(couts)
        xs = [torch.randn((nb,) + dhw + (c,), device=dev).to(bt) for c in cins]
        ws = [torch.randn(sh, device=dev) * 0.05 for sh in wshape]
        bs = [torch.zeros(co, device=dev) for co in couts]
        outs = [torch.empty((nb,) + out_dhw + (co,), device=dev, dtype=bt) for co in couts]
        douts = [torch.randn_like(o) for o in outs]
        line = '%-9s %5.1f GFLOP ' % (name, flops / 1e9)
        if 'fwd' in what:
            d = ops.conv_desc(mode, nb, dhw, out_dhw, k, s, pad, cins, couts, wstr, act_dtype=_lib.BF16,
                              engine=_lib.ENGINE_TCGEN05)
            d.tune[0] = args.variant
            packed = ops.conv3d_pack_weights(ctx, d, ws)
            t = timeit(lambda: ops.conv3d(ctx, d, xs, ws, bs, outs, packed), args.iters, flushbuf)
            line += '| fwd %7.3f ms %6.1f TF ' % (t, flops / t / 1e9)
        if 'dgrad' in what:
            tt = 0.0
            offs = np.cumsum([0] + cins)[:-1]
            dxs = [torch.empty_like(x) for x in xs]
            fused = (not tr) and len(couts) > 1 and all(c % 16 == 0 for c in couts)
            if fused:      # the engine's K-fused launch: K runs over [dy_j ...], per-(produced, gathered) weights
                dd = ops.conv_desc(_lib.CONV_TRANSPOSED, nb, out_dhw, dhw, k, s, pad, couts, cins,
                                   [(cin * co, 1, co) for co in couts], act_dtype=_lib.BF16,
                                   engine=_lib.ENGINE_TCGEN05, accumulate=[True] + [False] * (len(cins) - 1),
                                   w_by_src=True)
                dd.tune[0] = args.variant
                wv = [ws[j].view(-1)[int(o) * couts[j]:] for o in offs for j in range(len(couts))]
                pk = ops.conv3d_pack_weights(ctx, dd, wv)
                tt += timeit(lambda: ops.conv3d(ctx, dd, douts, wv, None, dxs, pk), args.iters, flushbuf)
            for j, co in enumerate([] if fused else couts):
                if tr:
                    dd = ops.conv_desc(_lib.CONV_FWD, nb, out_dhw, dhw, k, s, pad, [co], cins,
                                       [(co * cin, cin, 1)] * len(cins), act_dtype=_lib.BF16,
                                       engine=_lib.ENGINE_TCGEN05, accumulate=j > 0)
                    wv = [ws[0].view(-1)[int(o):] for o in offs]
                else:
                    dd = ops.conv_desc(_lib.CONV_TRANSPOSED, nb, out_dhw, dhw, k, s, pad, [co], cins,
                                       [(cin * co, 1, co)] * len(cins), act_dtype=_lib.BF16,
                                       engine=_lib.ENGINE_TCGEN05, accumulate=j > 0)
                    wv = [ws[j].view(-1)[int(o) * co:] for o in offs]
                dd.tune[0] = args.variant
                pk = ops.conv3d_pack_weights(ctx, dd, wv)
                tt += timeit(lambda: ops.conv3d(ctx, dd, [douts[j]], wv, None, dxs, pk), args.iters, flushbuf)
            line += '| dgrad %7.3f ms %6.1f TF ' % (tt, flops / tt / 1e9)
        if 'wgrad' in what:
            if tr:
                dws = [torch.zeros_like(ws[0])]
                dw = ops.conv_desc(_lib.CONV_FWD, nb, out_dhw, dhw, k, s, pad, couts, cins, [(couts[0] * cin, cin, 1)],
                                   act_dtype=_lib.BF16, engine=_lib.ENGINE_TCGEN05)
                t = timeit(lambda: ops.conv3d_wgrad(ctx, dw, douts, xs[:1], [dws[0].view(-1)], None), args.iters,
                           flushbuf)
            else:
                dws = [torch.zeros_like(w) for w in ws]
                dw = ops.conv_desc(mode, nb, dhw, out_dhw, k, s, pad, cins, couts, wstr, act_dtype=_lib.BF16,
                                   engine=_lib.ENGINE_TCGEN05)
                t = timeit(lambda: ops.conv3d_wgrad(ctx, dw, xs, douts, dws, None), args.iters, flushbuf)
            line += '| wgrad %7.3f ms %6.1f TF' % (t, flops / t / 1e9)
        print(line, flush=True)


if __name__ == '__main__':
    main()
