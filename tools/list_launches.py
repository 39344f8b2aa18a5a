#!/usr/bin/env python
"""Every convolution launch of one M1 training step (4-pass probabilistic graph), from a shape-only trace -
no GPU needed. For each forward launch: gathered/produced channels, grid, kernel, stride, algorithmic GFLOP and
which tcgen05 engines plan it (per-tap / halo / weight gradient / SHIFT weight gradient)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import m1b200  # noqa: E402,F401
from m1b200 import _lib, ops  # noqa: E402
from m1b200.model import unets  # noqa: E402
from m1b200.model.unets.engine import Engine  # noqa: E402
from m1b200.model.unets.params import ParamTable  # noqa: E402

README_CFG = dict(filters=(32, 64, 128, 256, 512),
                  strides=((1, 1, 1), (1, 2, 2), (1, 2, 2), (2, 2, 2), (2, 2, 2)),
                  kernel_sizes=((1, 3, 3), (1, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3)),
                  se_reduction=(8, 8, 8, 8, 8), att_sub_samp=((1, 1, 1),) * 4)


def trace(batch=8, dims=(20, 160, 160), dead=False, share=True):
    """list of launch records (Engine.trace_log) of one training step of the cfg-2 model. dead=True also lists the
    prior net's sersd0 + logits, which the reference builds but whose output only feeds an empty slice (Q3): the
    product does not execute them (SURVEY 8(d) counts them: 837.65 GMAC forward; live graph: 806.58 GMAC; executed
    with the stem + serse1 trunk shared between the passes that read the same input: 797.61 GMAC)."""
    m = unets.networks.M1(dims, 4, 2, dropout_mode='monte-carlo', dense_skip=True, deep_supervision=True,
                          probabilistic=True, prob_latent_dims=(3, 2, 1, 0), summary=False, build=False,
                          **README_CFG)
    eng = Engine(ParamTable(), 'bf16', device=None)
    eng.share_trunk = share      # False: every pass computes its own stem + serse1 (the reference graph)
    m._graph(eng, batch=batch, training=True, trace=True, dead=dead)
    return eng.trace_log


def fwd_desc(r):
    """m1_conv_desc of the forward launch of a trace record"""
    cin = sum(r['src_c'])
    if r['transposed']:
        co = r['out_c'][0]
        mode, wstr = _lib.CONV_TRANSPOSED, [(co * cin, 1, cin)]
    else:
        mode, wstr = _lib.CONV_FWD, [(cin * co, co, 1) for co in r['out_c']]
    return ops.conv_desc(mode, r['batch'], r['in_dhw'], r['out_dhw'], r['kernel'], r['stride'], r['pad'],
                         r['src_c'], r['out_c'], wstr, act_dtype=_lib.BF16,
                         out_dtype=_lib.F32 if r['out_fp32'] else _lib.BF16)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=8)
    args = ap.parse_args()
    log = trace(args.batch)
    tot = 0.0
    print('%-34s %-22s %-12s %-14s %-9s %-9s %8s  engines' % ('layer(s)', 'gathered', 'produced', 'grid', 'kernel',
                                                               'stride', 'GFLOP'))
    for r in log:
        d = fwd_desc(r)
        eng = []
        if ops.conv3d_tc_supported(d):
            eng.append('tap')
            d.tune[0] = 0
            if ops.conv3d_plan_info(d, 1):
                eng.append('halo*' if ops.conv3d_halo_engine(d) else 'halo')
        if not r['transposed'] and ops.conv3d_wgrad_tc_supported(d):
            eng.append('wgrad')
            d.tune[1] = 2
            info = ops.conv3d_plan_info(d, 2)
            if info and info[12]:
                eng.append('shift')
        tot += r['flops']
        print('%-34s %-22s %-12s %-14s %-9s %-9s %8.1f  %s' % (
            '+'.join(n.split('/', 1)[-1] for n in r['names'])[:34] + ('(T)' if r['transposed'] else ''),
            str(r['src_c'])[:22], str(r['out_c']), 'x'.join(map(str, r['out_dhw'])), 'x'.join(map(str, r['kernel'])),
            'x'.join(map(str, r['stride'])), r['flops'] / 1e9, ','.join(eng) or 'cuda-core'))
    print('%d forward launches, %.1f GFLOP forward (x3 with both gradients) for batch %d' % (len(log), tot / 1e9,
                                                                                           args.batch))


if __name__ == '__main__':
    main()
