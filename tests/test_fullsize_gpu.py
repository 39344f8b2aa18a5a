"""Parity at BASELINE (cfg-2) shapes: every distinct convolution launch family of the 20x160x160 training step is
run ONCE at full spatial size (batch 1) through the production path - Engine.conv -> m1_conv3d (variant chosen by
the one-off autotuning), Engine._conv_bwd -> m1_conv3d_wgrad + the data-gradient launches - in the benchmarked
precision (fp16 values and weights, bf16 gradients) and compared with autograd of the oracle convolution on the
same rounded operands; the halo / SHIFT variants are additionally forced at the 160x160 grid; and one whole-model
training step at 20x160x160 (batch 1) is compared with the fp32 oracle at the north-star bounds."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import m1_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'
H, B = torch.float16, torch.bfloat16


def _rel(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float()
    assert torch.isfinite(got).all(), "non-finite output (unwritten voxels?)"
    return (got - ref).abs().max().item() / max(1e-6, ref.abs().max().item())


def _layer(ctx, dhw, src_c, src_lc, layers, k, s, transposed, pad_out=None, seed=0):
    """One Engine.conv launch family at batch 1: forward, both gradients. src_c / src_lc: physical / logical
    channels of the gathered tensors; layers: [(name, cout)]. Returns the worst relative errors (fwd, dgrad, wgrad)."""
    from m1b200.model.unets.engine import Engine
    from m1b200.model.unets.params import ParamTable
    g = torch.Generator().manual_seed(seed)
    params = ParamTable()
    tr = Engine(params, 'fp16', device=None)
    tr.conv([tr.input((1, *dhw, c), needs_grad=True, lc=lc) for c, lc in zip(src_c, src_lc)], layers, k, s,
            transposed=transposed, pad_out=pad_out)
    params.finalize()
    params.allocate(torch.device(DEV))
    lcin = sum(src_lc)
    taps = int(np.prod(k))
    weights = {}
    for name, co in layers:
        shape = (*k, co, lcin) if transposed else (*k, lcin, co)
        weights[name + '/kernel'] = (torch.randn(shape, generator=g) / (lcin * taps / (int(np.prod(s)) if transposed else 1)) ** 0.5
                                     ).to(H).float().numpy()
        weights[name + '/bias'] = (torch.randn((co,), generator=g) * 0.1).numpy()
    params.load_state_dict(weights)
    eng = Engine(params, 'fp16', device=torch.device(DEV))
    eng.begin(record=True)
    xs = []
    for c, lc in zip(src_c, src_lc):
        t = torch.zeros((1, *dhw, c))
        t[..., :lc] = torch.randn((1, *dhw, lc), generator=g)
        xs.append(t.to(H))
    acts = [eng.input(t.to(DEV), needs_grad=True, lc=lc) for t, lc in zip(xs, src_lc)]
    outs = eng.conv(acts, layers, k, s, transposed=transposed, pad_out=pad_out)
    torch.cuda.synchronize()
    # oracle: fp32 autograd on the same (fp16-rounded) operands and bf16-rounded output gradients
    xr = [t[..., :lc].float().requires_grad_() for t, lc in zip(xs, src_lc)]
    xcat = torch.cat(xr, -1)
    errs = {'fwd': 0.0, 'dgrad': 0.0, 'wgrad': 0.0}
    dys, wr = [], []
    for (name, co), o in zip(layers, outs):
        w = torch.from_numpy(weights[name + '/kernel']).requires_grad_()
        b = torch.from_numpy(weights[name + '/bias'])
        y = (O.conv3d_transpose_same if transposed else O.conv3d_same)(xcat, w, b, s)
        assert o.t.shape[:-1] == y.shape[:-1] and o.lc == co
        errs['fwd'] = max(errs['fwd'], _rel(o.t[..., :co], y))
        if o.c > co:
            assert o.t[..., co:].abs().max().item() == 0.0, "padded output channels must be exactly zero"
        dy = torch.randn(y.shape, generator=g).to(B)
        y.backward(dy.float(), retain_graph=True)
        dyp = torch.zeros(o.t.shape, dtype=B)
        dyp[..., :co] = dy
        o.g = dyp.to(DEV)
        dys.append(dy)
        wr.append(w)
    params.g.zero_()
    eng.backward()
    torch.cuda.synchronize()
    for a, x_, lc in zip(acts, xr, src_lc):
        assert a.g.dtype == B
        errs['dgrad'] = max(errs['dgrad'], _rel(a.g[..., :lc], x_.grad))
    gd = params.grad_dict()
    for (name, co), w in zip(layers, wr):
        errs['wgrad'] = max(errs['wgrad'], _rel(gd[name + '/kernel'], w.grad))
    tuned = {k_[0]: v for k_, v in eng.tuned.items()}
    print(layers[0][0], 'errors', errs, 'autotuned', tuned)
    return errs


# name, grid, physical / logical gathered channels, layers, kernel, stride, transposed, pad_out  (cfg-2)
FAMILIES = [
    ('sersp2', (20, 40, 40), [128] * 4, [128] * 4, [('l/conv1', 32), ('l/conv4', 128)], (3, 3, 3), (1, 1, 1), False, [True, False]),
    ('sersp0', (20, 160, 160), [32] * 6, [32] * 6, [('l/conv1', 8), ('l/conv4', 32)], (1, 3, 3), (1, 1, 1), False, [True, False]),
    ('sersp3', (10, 20, 20), [256] * 3, [256] * 3, [('l/conv1', 64), ('l/conv4', 256)], (3, 3, 3), (1, 1, 1), False, [True, False]),
    ('sersd1', (20, 80, 80), [64] * 4, [64] * 4, [('l/conv1', 16), ('l/conv4', 64)], (1, 3, 3), (1, 1, 1), False, [True, False]),
    ('serse3', (20, 40, 40), [128], [128], [('l/conv1', 64), ('l/conv4', 256)], (3, 3, 3), (2, 2, 2), False, [True, False]),
    ('serse1', (20, 160, 160), [32], [32], [('l/conv1', 16), ('l/conv4', 64)], (1, 3, 3), (1, 2, 2), False, [True, False]),
    ('conv2_r0', (20, 160, 160), [16], [8], [('l/conv2', 8)], (3, 3, 3), (1, 1, 1), False, [True]),
    ('conv3_r0', (20, 160, 160), [16], [8], [('l/conv3', 32)], (1, 1, 1), (1, 1, 1), False, None),
    ('conve0', (20, 160, 160), [16], [3], [('l/conve0', 32)], (1, 3, 3), (1, 1, 1), False, None),
    ('convtd1', (20, 40, 40), [128], [128], [('l/convtd1', 64)], (3, 3, 3), (1, 2, 2), True, None),
    ('convtd0', (20, 80, 80), [64], [64], [('l/convtd0', 32)], (1, 3, 3), (1, 2, 2), True, None),
    ('dec_hi3', (5, 10, 10), [16, 512], [3, 512], [('l/dec_hi3', 256)], (3, 3, 3), (2, 2, 2), True, None),
    ('dec_hi1', (20, 40, 40), [16, 128], [1, 128], [('l/dec_hi1', 64)], (3, 3, 3), (1, 2, 2), True, None),
]


@pytest.mark.parametrize("name,dhw,src_c,src_lc,layers,k,s,transposed,pad_out", FAMILIES, ids=[f[0] for f in FAMILIES])
def test_launch_family_at_cfg2_shape(ctx, name, dhw, src_c, src_lc, layers, k, s, transposed, pad_out):
    errs = _layer(ctx, dhw, src_c, src_lc, layers, k, s, transposed, pad_out, seed=__import__('zlib').crc32(name.encode()) % 1000)
    assert errs['fwd'] < 2e-3, errs          # exact products, fp32 accumulation, fp16 output rounding (2^-11)
    assert errs['dgrad'] < 1.5e-2, errs      # bf16 output rounding (2^-8)
    assert errs['wgrad'] < 4e-3, errs        # bf16 twin of the activations (2^-9 per element), fp32 accumulation


def test_forced_variants_at_160x160(ctx):
    """sersp0 at the full 160x160 grid with the variants FORCED: halo forward (G stacked sub-tiles), halo K-fused
    data gradient, SHIFT-mode weight gradient at W = 160 (row pitch 176)."""
    from m1b200 import ops, _lib
    g = torch.Generator().manual_seed(77)
    dhw, cins, couts, k = (20, 160, 160), [32] * 6, [16, 32], (1, 3, 3)
    cin = sum(cins)
    xs = [torch.randn((1, *dhw, c), generator=g).to(H) for c in cins]
    ws = [(torch.randn((*k, cin, co), generator=g) / (cin * 9) ** 0.5).to(H).float() for co in couts]
    xr = [x.float().requires_grad_() for x in xs]
    wr = [w.clone().requires_grad_() for w in ws]
    dys = []
    ys = []
    for w in wr:
        y = O.conv3d_same(torch.cat(xr, -1), w, None, (1, 1, 1))
        dy = torch.randn(y.shape, generator=g).to(B)
        y.backward(dy.float(), retain_graph=True)
        ys.append(y)
        dys.append(dy)
    pad = [ops.same_pads(dhw[i], k[i], 1)[1] for i in range(3)]
    xd = [x.to(DEV) for x in xs]
    wd = [w.to(DEV).contiguous() for w in ws]
    # halo forward
    d = ops.conv_desc(_lib.CONV_FWD, 1, dhw, dhw, k, (1, 1, 1), pad, cins, couts, [(cin * co, co, 1) for co in couts],
                      act_dtype=_lib.F16, engine=_lib.ENGINE_TCGEN05)
    d.tune[0] = 2
    assert _lib.lib().m1_conv3d_halo_engine(ctypes.byref(d)) == 1
    info = ops.conv3d_plan_info(d, 1)
    assert info[3] >= 2, "expected G-stacked sub-tiles at 160x160: %r" % (info,)
    outs = [torch.full((1, *dhw, co), float('nan'), device=DEV, dtype=H) for co in couts]
    ops.conv3d(ctx, d, xd, wd, None, outs, ops.conv3d_pack_weights(ctx, d, wd))
    torch.cuda.synchronize()
    for o, y in zip(outs, ys):
        assert _rel(o, y) < 2e-3
    # halo K-fused data gradient (bf16 dY x bf16 weight pack)
    offs = [sum(cins[:i]) for i in range(len(cins))]
    wv = [wd[j].view(-1)[o * couts[j]:] for o in offs for j in range(len(couts))]
    dd = ops.conv_desc(_lib.CONV_TRANSPOSED, 1, dhw, dhw, k, (1, 1, 1), pad, couts, cins,
                       [(cin * co, 1, co) for co in couts], act_dtype=_lib.BF16, engine=_lib.ENGINE_TCGEN05,
                       w_by_src=True)
    dd.tune[0] = 2
    assert _lib.lib().m1_conv3d_halo_engine(ctypes.byref(dd)) == 1
    bufs = [torch.full(x.shape, float('nan'), device=DEV, dtype=B) for x in xs]
    ops.conv3d(ctx, dd, [t.to(DEV) for t in dys], wv, None, bufs, ops.conv3d_pack_weights(ctx, dd, wv))
    torch.cuda.synchronize()
    for b, x_ in zip(bufs, xr):
        assert _rel(b, x_.grad) < 1.5e-2
    # SHIFT-mode weight gradient
    # SHIFT-mode weight gradient on the bf16 twins of the activations
    dw_ = ops.conv_desc(_lib.CONV_FWD, 1, dhw, dhw, k, (1, 1, 1), pad, cins, couts, [(cin * co, co, 1) for co in couts],
                        act_dtype=_lib.BF16, out_dtype=_lib.BF16, engine=_lib.ENGINE_TCGEN05)
    dw_.tune[0], dw_.tune[1], dw_.tune[2], dw_.tune[3] = 192, 2, 3, 1
    info = ops.conv3d_plan_info(dw_, 2)
    assert info and info[12] == 1, "SHIFT mode refused W = 160: %r" % (info,)
    dws = [torch.zeros(w.shape, device=DEV) for w in ws]
    ops.conv3d_wgrad(ctx, dw_, [x.to(B) for x in xd], [t.to(DEV) for t in dys], dws, None)
    torch.cuda.synchronize()
    for dwt, w in zip(dws, wr):
        assert _rel(dwt, w.grad) < 4e-3          # bf16 twin of the activations: 2^-9 relative per element


def test_whole_model_step_at_20x160x160(ctx):
    """One training step of the full cfg-2 model (README filters, probabilistic + dense_skip + deep_supervision,
    monte-carlo dropout) on ONE 20x160x160 volume in the benchmarked precision against the fp32 oracle: identical
    fp32 inputs, weights, injected dropout uniforms and latent noise. North-star bounds: per-voxel softmax 2e-2 max
    abs, focal and KL 1e-3 relative; gradient cosine >= 0.98."""
    from m1b200.model import losses, optimizers, unets
    S = ((1, 1, 1), (1, 2, 2), (1, 2, 2), (2, 2, 2), (2, 2, 2))
    K = ((1, 3, 3), (1, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3))
    F = (32, 64, 128, 256, 512)
    dims = (20, 160, 160)
    torch.set_num_threads(max(1, __import__('os').cpu_count() or 1))
    model = unets.networks.M1(dims, 4, 2, dropout_rate=0.5, dropout_mode='monte-carlo', filters=F, strides=S,
                              kernel_sizes=K, se_reduction=(8,) * 5, att_sub_samp=((1, 1, 1),) * 4, dense_skip=True,
                              deep_supervision=True, probabilistic=True, prob_latent_dims=(3, 2, 1, 0), summary=False,
                              precision='fp16', seed=0)
    model.compile(optimizer=optimizers.Adam(1e-3, amsgrad=True),
                  loss=[losses.Focal(alpha=[0.75, 0.25], gamma=2.0).loss, losses.EvidenceLowerBound().loss],
                  loss_weights=[1.0, 10.0])
    cfg = O.default_config(num_classes=2, dropout_rate=0.5, dropout_mode='monte-carlo', strides=S, kernel_sizes=K,
                           dense_skip=True, deep_supervision=True, probabilistic=True, prob_latent_dims=(3, 2, 1, 0),
                           filters=F, se_reduction=(8,) * 5)
    x, y = O.synthetic_batch(1, dims, probabilistic=True, seed=11)
    ps = O.ParamStore(dtype=torch.float32, seed=3, requires_grad=True)
    with torch.no_grad():
        O.train_loss(ps, cfg, x, y, O.Noise(0, torch.float32))
    g = torch.Generator().manual_seed(17)
    with torch.no_grad():
        for n, t in ps.p.items():
            if ps.kind[n] in ('gamma', 'beta', 'se_bias'):
                t.add_(0.2 * torch.randn(t.shape, generator=g).to(t.dtype))
    noise = O.Noise(5, torch.float32)
    r = O.train_loss(ps, cfg, x, y, noise, alpha=(0.75, 0.25), gamma=2.0, kl_weight=10.0)
    (r['detection_loss'] + 10.0 * r['KL_loss']).backward()
    model.set_weights({n: t.detach().float().numpy() for n, t in ps.p.items()})
    model.set_noise(noise.t)
    out = model.train_step(x, y, apply_update=False)
    torch.cuda.synchronize()
    e = (out['detection'].double().cpu() - r['detection'].detach().double()).abs().flatten()
    fl, fl_ref = out['focal'].item(), r['detection_loss'].item()
    kl, kl_ref = out['kl'].item(), r['KL'].item()
    grads = model.gradients()
    a = torch.cat([grads[n].double().flatten() for n in ps.p])
    b = torch.cat([(t.grad if t.grad is not None else torch.zeros_like(t)).double().flatten() for t in ps.p.values()])
    cos = (a @ b).item() / (a.norm().item() * b.norm().item())
    print(f'cfg-2 volume, fp16: softmax abs err mean {e.mean().item():.2e} max {e.max().item():.2e} | focal rel '
          f'{abs(fl - fl_ref) / abs(fl_ref):.2e} | KL rel {abs(kl - kl_ref) / abs(kl_ref):.2e} | grad cosine {cos:.4f}')
    assert e.max().item() < 2e-2, e.max().item()
    assert abs(fl - fl_ref) < 1e-3 * abs(fl_ref), (fl, fl_ref)
    assert abs(kl - kl_ref) < 1e-3 * abs(kl_ref), (kl, kl_ref)
    assert cos >= 0.98, cos
