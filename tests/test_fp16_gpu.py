"""precision='fp16': fp16 activation VALUES and tensor-core weights, bf16 activation GRADIENTS (M1_GRAD_DTYPE),
fp32 accumulation. Each kernel family against the CPU oracle on operands rounded the way the kernels see them,
the bf16 twins that feed the tensor-core weight gradients (one operand format per MMA), the determinism of the
forward reductions, and the whole model against the fp32 oracle at the north-star bounds
(softmax 2e-2 max abs, focal and KL 1e-3 relative, gradient cosine >= 0.98) - un-relaxed."""
import ctypes

import pytest
import torch

from oracle import m1_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'
H, B = torch.float16, torch.bfloat16


def _h(t):
    return t.to(H).double()


def _b(t):
    return t.to(B).double()


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _close(got, ref, tol, what=''):
    got = got.detach().double().cpu()
    ref = ref.detach().double()
    scale = max(1.0, ref.abs().max().item())
    err = (got - ref).abs().max().item() / scale
    assert torch.isfinite(got).all(), what
    assert err < tol, f'{what}: rel err {err:.3e} (tol {tol})'


# ---- K4 / K5 / K6 on fp16 values + bf16 gradients -------------------------------------------------------------
@pytest.mark.parametrize("C,slope", [(8, 0.1), (64, 0.1), (12, 1.0)])
def test_inorm_act_fp16(ctx, C, slope):
    from m1b200 import ops
    g = _gen(1)
    shape = (2, 3, 10, 12, C)
    x = _h(torch.randn(shape, generator=g) * 2 + 0.5)
    gamma = (torch.rand(C, generator=g) + 0.5).double()
    beta = (torch.randn(C, generator=g) * 0.3).double()
    dy = _b(torch.randn(shape, generator=g))
    xr, gr, br = x.clone().requires_grad_(), gamma.clone().requires_grad_(), beta.clone().requires_grad_()
    y_ref = O.lrelu(O.instance_norm(xr, gr, br), slope)
    y_ref.backward(dy)
    xd = x.to(DEV, H)
    stats = torch.empty((2, C, 2), device=DEV)
    y = torch.empty_like(xd)
    ops.inorm_stats(ctx, xd, stats)
    ops.inorm_act_fwd(ctx, xd, stats, gamma.float().to(DEV), beta.float().to(DEV), slope, y)
    _close(y, y_ref, 2e-3, 'y')                       # fp16 output rounding: 2^-11 relative
    _close(stats[..., 0], xr.mean(dim=(1, 2, 3)), 1e-5, 'mean')
    dx = torch.empty(shape, device=DEV, dtype=B)     # the gradient of an fp16 activation is bf16
    dgam, dbet = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    ops.inorm_act_bwd(ctx, dy.to(DEV, B), xd, stats, gamma.float().to(DEV), beta.float().to(DEV), slope, dx, False,
                      dgam, dbet)
    torch.cuda.synchronize()
    _close(dx, xr.grad, 1e-2, 'dx')
    _close(dgam, gr.grad, 1e-3, 'dgamma')
    _close(dbet, br.grad, 1e-3, 'dbeta')


@pytest.mark.parametrize("C,red,rate", [(16, 4, 0.5), (32, 8, 0.0), (256, 8, 0.5)])
def test_se_tail_fp16(ctx, C, red, rate):
    from m1b200 import ops
    g = _gen(2)
    shape = (2, 3, 6, 8, C)
    Cr = C // red
    raw3 = _h(torch.randn(shape, generator=g))
    raw4 = _h(torch.randn(shape, generator=g) * 1.5 + 0.2)
    P = dict(g3=torch.rand(C, generator=g) + 0.5, b3=torch.randn(C, generator=g) * 0.5,
             g4=torch.rand(C, generator=g) + 0.5, b4=torch.randn(C, generator=g) * 0.5,
             w6=torch.randn((C, Cr), generator=g) * 0.3, b6=torch.randn(Cr, generator=g) * 0.1,
             w7=torch.randn((Cr, C), generator=g) * 0.3, b7=torch.randn(C, generator=g) * 0.1)
    u = torch.rand(shape, generator=g)
    dout = _b(torch.randn(shape, generator=g))
    R = {k: v.double().requires_grad_() for k, v in P.items()}
    r3, r4 = raw3.clone().requires_grad_(), raw4.clone().requires_grad_()
    x_ = O.instance_norm(r3, R['g3'], R['b3'])
    res = O.instance_norm(r4, R['g4'], R['b4'])
    pool = x_.mean(dim=(1, 2, 3))
    gate = torch.sigmoid(O.lrelu(pool @ R['w6'] + R['b6']) @ R['w7'] + R['b7'])
    out_ref = O.dropout(O.lrelu(x_ * gate[:, None, None, None, :] * res), rate, u.double())
    out_ref.backward(dout)

    D = {k: v.to(DEV) for k, v in P.items()}
    d3, d4 = raw3.to(DEV, H), raw4.to(DEV, H)
    st3, st4 = torch.empty((2, C, 2), device=DEV), torch.empty((2, C, 2), device=DEV)
    ops.inorm_stats(ctx, d3, st3)
    ops.inorm_stats(ctx, d4, st4)
    poold, hidden, gated = (torch.empty((2, C), device=DEV), torch.empty((2, Cr), device=DEV),
                            torch.empty((2, C), device=DEV))
    # the production sequence: squeeze folded into the excite launch (pool from the statistics of raw3)
    ops.se_excite_fwd(ctx, poold, D['w6'], D['b6'], D['w7'], D['b7'], hidden, gated, st3, D['g3'], D['b3'])
    ud = u.to(DEV)
    drop = ops.make_dropout(rate, ud)
    out = torch.empty_like(d3)
    ops.se_gate_fwd(ctx, d3, d4, st3, st4, D['g3'], D['b3'], D['g4'], D['b4'], gated, drop, out)
    _close(poold, pool, 1e-5, 'pool (from the statistics: no second read of raw3)')
    _close(gated, gate, 1e-5, 'gate')
    _close(out, out_ref, 2e-3, 'out')
    red5, dgate = torch.empty((2, C, 5), device=DEV), torch.empty((2, C), device=DEV)
    dd = dout.to(DEV, B)
    ops.se_gate_bwd_reduce(ctx, dd, d3, d4, st3, st4, D['g3'], D['b3'], D['g4'], D['b4'], gated, drop, red5, dgate)
    G = {k: torch.zeros_like(v) for k, v in D.items()}
    dpool = torch.empty((2, C), device=DEV)
    # ... and the norm3 / norm4 parameter gradients accumulated by the excite backward (no param_grad launches)
    ops.se_excite_bwd(ctx, dgate, poold, hidden, gated, D['w6'], D['w7'], dpool, G['w6'], G['b6'], G['w7'], G['b7'],
                      red5, G['g3'], G['b3'], G['g4'], G['b4'])
    dr3, dr4 = torch.empty(shape, device=DEV, dtype=B), torch.empty(shape, device=DEV, dtype=B)
    ops.se_gate_bwd_apply(ctx, dd, d3, d4, st3, st4, D['g3'], D['b3'], D['g4'], D['b4'], gated, drop, red5, dpool,
                          dr3, dr4, None, None, None, None)
    torch.cuda.synchronize()
    _close(dr3, r3.grad, 1e-2, 'draw3')
    _close(dr4, r4.grad, 1e-2, 'draw4')
    for k in P:
        _close(G[k], R[k].grad, 1e-3, 'd' + k)


@pytest.mark.parametrize("F,xg,gg", [(32, (4, 16, 16), (1, 1, 1)), (64, (4, 16, 32), (2, 2, 4)), (12, (4, 8, 8), (1, 2, 2))])
def test_attention_gate_fp16(ctx, F, xg, gg):
    from m1b200 import ops
    g = _gen(3)
    n = 2
    theta = _h(torch.randn((n, *xg, F), generator=g))
    phi = _h(torch.randn((n, *gg, F), generator=g))
    x = _h(torch.randn((n, *xg, F), generator=g))
    w = torch.randn(F, generator=g).double() * 0.3
    bpsi = torch.randn(1, generator=g).double() * 0.1
    dy = _b(torch.randn((n, *xg, F), generator=g))
    tr, pr, xr, wr, br = (t.clone().requires_grad_() for t in (theta, phi, x, w, bpsi))
    up = O.upsample_nearest(pr, [xg[i] // gg[i] for i in range(3)])
    psi = torch.sigmoid((O.lrelu(tr + up) * wr).sum(-1, keepdim=True) + br)
    y_ref = psi * xr
    y_ref.backward(dy)
    td, pd, xd = theta.to(DEV, H), phi.to(DEV, H), x.to(DEV, H)
    wd, bd = w.float().to(DEV), bpsi.float().to(DEV)
    psid = torch.empty((n, *xg), device=DEV)
    y = torch.empty_like(xd)
    ops.attn_fwd(ctx, td, pd, wd, bd, xd, psid, y)
    _close(psid, psi[..., 0], 1e-5, 'psi')
    _close(y, y_ref, 2e-3, 'y')
    dx = torch.empty(x.shape, device=DEV, dtype=B)
    dth = torch.empty(theta.shape, device=DEV, dtype=B)
    dphi = torch.zeros(phi.shape, device=DEV)
    dw, db = torch.zeros(F, device=DEV), torch.zeros(1, device=DEV)
    ops.attn_bwd(ctx, dy.to(DEV, B), td, pd, wd, psid, xd, dx, False, dth, dphi, dw, db)
    torch.cuda.synchronize()
    _close(dx, xr.grad, 1e-2, 'dx')
    _close(dth, tr.grad, 1e-2, 'dtheta')
    _close(dphi, pr.grad, 1e-3, 'dphi')
    _close(dw, wr.grad, 1e-3, 'dw')
    _close(db, br.grad, 1e-3, 'db')


# ---- tcgen05 convolutions in fp16 -----------------------------------------------------------------------------
def _mk(batch, dhw, cins, couts, k, seed=0):
    g = _gen(seed)
    xs = [_h(torch.randn((batch, *dhw, c), generator=g)) for c in cins]
    cin = sum(cins)
    ws = [_h(torch.randn((*k, cin, co), generator=g) / (cin * k[0] * k[1] * k[2]) ** 0.5) for co in couts]
    bs = [torch.randn((co,), generator=g).double() * 0.1 for co in couts]
    return xs, ws, bs


FWD_CASES = [((4, 16, 16), [64], [16, 64], (3, 3, 3), (1, 1, 1), 1),
             ((6, 20, 20), [128, 64], [32, 128], (3, 3, 3), (1, 1, 1), 1),
             ((4, 16, 32), [32, 32, 32], [8, 32], (1, 3, 3), (1, 1, 1), 1),
             ((8, 16, 16), [64, 32], [32, 128], (3, 3, 3), (2, 2, 2), 1),        # strided
             ((3, 16, 44), [32, 32, 32], [16, 32], (1, 3, 3), (1, 1, 1), 2),      # halo variant
             ((2, 6, 40), [128], [64, 256], (3, 3, 3), (1, 1, 1), 2)]


@pytest.mark.parametrize("dhw,cins,couts,k,s,variant", FWD_CASES)
def test_conv_fwd_fp16(ctx, dhw, cins, couts, k, s, variant):
    """f16 x f16 tcgen05.mma (instruction-descriptor formats 0/0), fp16 outputs"""
    from m1b200 import ops, _lib
    xs, ws, bs = _mk(2, dhw, cins, couts, k, seed=41)
    geo = [ops.same_pads(dhw[i], k[i], s[i]) for i in range(3)]
    out_dhw, pad = [g[0] for g in geo], [g[1] for g in geo]
    cin = sum(cins)
    d = ops.conv_desc(_lib.CONV_FWD, 2, dhw, out_dhw, k, s, pad, cins, couts, [(cin * co, co, 1) for co in couts],
                      act_dtype=_lib.F16, engine=_lib.ENGINE_TCGEN05)
    d.tune[0] = variant
    assert ops.conv3d_tc_supported(d)
    if variant == 2:
        assert _lib.lib().m1_conv3d_halo_engine(ctypes.byref(d)) == 1
    wd = [w.float().to(DEV).contiguous() for w in ws]
    packed = ops.conv3d_pack_weights(ctx, d, wd)
    assert packed.dtype == H
    outs = [torch.full((2, *out_dhw, co), float('nan'), device=DEV, dtype=H) for co in couts]
    ops.conv3d(ctx, d, [x.to(DEV, H).contiguous() for x in xs], wd, [b.float().to(DEV) for b in bs], outs, packed)
    torch.cuda.synchronize()
    x = torch.cat(xs, -1)
    for o, w, b in zip(outs, ws, bs):
        ref = O.conv3d_same(x, w, b, s)
        _close(o, ref, 2e-3, 'fp16 conv')            # exact products, fp32 accumulation, fp16 output rounding


@pytest.mark.parametrize("dhw,k,s,cin,cout", [((3, 8, 8), (3, 3, 3), (2, 2, 2), 64, 32),
                                              ((4, 10, 6), (1, 3, 3), (1, 2, 2), 64, 32)])
def test_conv_transpose_fp16(ctx, dhw, k, s, cin, cout):
    from m1b200 import ops, _lib
    g = _gen(6)
    x = _h(torch.randn((2, *dhw, cin), generator=g))
    w = _h(torch.randn((*k, cout, cin), generator=g) / (cin * 4) ** 0.5)
    b = torch.randn((cout,), generator=g).double() * 0.1
    out_dhw = [dhw[i] * s[i] for i in range(3)]
    pad = [ops.same_pads(out_dhw[i], k[i], s[i])[1] for i in range(3)]
    d = ops.conv_desc(_lib.CONV_TRANSPOSED, 2, dhw, out_dhw, k, s, pad, [cin], [cout], [(cout * cin, 1, cin)],
                      act_dtype=_lib.F16, engine=_lib.ENGINE_TCGEN05)
    assert ops.conv3d_tc_supported(d)
    wd, bd = w.float().to(DEV).contiguous(), b.float().to(DEV)
    packed = ops.conv3d_pack_weights(ctx, d, [wd])
    out = torch.full((2, *out_dhw, cout), float('nan'), device=DEV, dtype=H)
    ops.conv3d(ctx, d, [x.to(DEV, H).contiguous()], [wd], [bd], [out], packed)
    torch.cuda.synchronize()
    _close(out, O.conv3d_transpose_same(x, w, b, s), 2e-3, 'fp16 conv transpose')


def test_mixed_operand_formats_are_refused(ctx):
    """tcgen05.mma.kind::f16 with different A and B formats traps with 'illegal instruction' on B200 (measured both
    ways in round 2: bf16 x fp16 and fp16 x bf16). The planners therefore refuse such launches - the engine never
    issues them: fp16 mode multiplies fp16 x fp16 (forward), bf16 x bf16 (data gradient, bf16 weight pack) and
    bf16 twin x bf16 (weight gradient)."""
    from m1b200 import ops, _lib
    dhw, k = (4, 16, 16), (3, 3, 3)
    pad = [1, 1, 1]
    d = ops.conv_desc(_lib.CONV_TRANSPOSED, 2, dhw, dhw, k, (1, 1, 1), pad, [64], [64], [(64 * 64, 1, 64)],
                      act_dtype=_lib.BF16, engine=_lib.ENGINE_TCGEN05, w_dtype=_lib.F16)
    assert not ops.conv3d_tc_supported(d)
    d.w_dtype = 0
    assert ops.conv3d_tc_supported(d)
    dw = ops.conv_desc(_lib.CONV_FWD, 2, dhw, dhw, k, (1, 1, 1), pad, [64], [64], [(64 * 64, 64, 1)],
                       act_dtype=_lib.F16, out_dtype=_lib.BF16, engine=_lib.ENGINE_AUTO)
    assert not ops.conv3d_wgrad_tc_supported(dw)
    dw.act_dtype = _lib.BF16
    assert ops.conv3d_wgrad_tc_supported(dw)


def test_bf16_twins_of_fp16_outputs(ctx):
    """The forward kernels that produce an activation can store it twice: fp16 for the forward convolutions and a
    bf16 twin for the tensor-core weight gradients - same registers, two roundings."""
    from m1b200 import ops
    g = _gen(7)
    C, shape = 32, (2, 3, 6, 8, 32)
    x = torch.randn(shape, generator=g).to(DEV, H)
    gamma, beta = (torch.rand(C, generator=g) + 0.5).to(DEV), (torch.randn(C, generator=g) * 0.3).to(DEV)
    st = torch.empty((2, C, 2), device=DEV)
    ops.inorm_stats(ctx, x, st)
    y, y2 = torch.empty_like(x), torch.full(shape, float('nan'), device=DEV, dtype=B)
    ops.inorm_act_fwd(ctx, x, st, gamma, beta, 0.1, y, y2)
    torch.cuda.synchronize()
    assert torch.isfinite(y2.float()).all()
    assert (y2.float() - y.float()).abs().max().item() <= 2.0 ** -8 * y.float().abs().max().item()
    gate = torch.rand((2, C), generator=g).to(DEV)
    o, o2 = torch.empty_like(x), torch.full(shape, float('nan'), device=DEV, dtype=B)
    ops.se_gate_fwd(ctx, x, y, st, st, gamma, beta, gamma, beta, gate, ops.make_dropout(0.0), o, o2)
    torch.cuda.synchronize()
    assert torch.isfinite(o2.float()).all()
    assert (o2.float() - o.float()).abs().max().item() <= 2.0 ** -8 * o.float().abs().max().item()
    theta, phi = torch.randn(shape, generator=g).to(DEV, H), torch.randn((2, 3, 6, 8, C), generator=g).to(DEV, H)
    psi = torch.empty(shape[:-1], device=DEV)
    a, a2 = torch.empty_like(x), torch.full(shape, float('nan'), device=DEV, dtype=B)
    ops.attn_fwd(ctx, theta, phi, torch.randn(C, generator=g).to(DEV), torch.zeros(1, device=DEV), x, psi, a, a2)
    torch.cuda.synchronize()
    assert torch.isfinite(a2.float()).all()
    assert (a2.float() - a.float()).abs().max().item() <= 2.0 ** -8 * a.float().abs().max().item()


@pytest.mark.parametrize("C,N", [(128, 2), (512, 6), (32, 2)])
def test_pointwise_heads_fp16(ctx, C, N):
    from m1b200 import ops, _lib
    g = _gen(21)
    dhw = (3, 7, 9)
    x = _h(torch.randn((2, *dhw, C), generator=g)).requires_grad_()
    w = (torch.randn((1, 1, 1, C, N), generator=g, dtype=torch.float64) / C ** 0.5).requires_grad_()
    b = torch.randn((N,), generator=g, dtype=torch.float64) * 0.1
    y = O.conv3d_same(x, w, b, (1, 1, 1))
    dy = torch.randn(y.shape, generator=g).double()
    y.backward(dy)
    xd = x.detach().to(DEV, H).contiguous()
    wd = w.detach().float().to(DEV).contiguous()
    d = ops.conv_desc(_lib.CONV_FWD, 2, dhw, dhw, (1, 1, 1), (1, 1, 1), (0, 0, 0), [C], [N], [(C * N, N, 1)],
                      act_dtype=_lib.F16, out_dtype=_lib.F32, engine=_lib.ENGINE_SIMT)
    out = torch.full((2, *dhw, N), float('nan'), device=DEV)
    ops.conv3d(ctx, d, [xd], [wd], [b.float().to(DEV)], [out])
    _close(out, y, 1e-4, 'head')
    dw, db = torch.zeros(w.shape, device=DEV), torch.zeros(N, device=DEV)
    dyd = dy.float().to(DEV).contiguous()
    ops.conv3d_wgrad(ctx, d, [xd], [dyd], [dw], [db])
    _close(dw, w.grad, 1e-4, 'head dW')
    dd = ops.conv_desc(_lib.CONV_TRANSPOSED, 2, dhw, dhw, (1, 1, 1), (1, 1, 1), (0, 0, 0), [N], [C], [(C * N, 1, N)],
                       act_dtype=_lib.F32, out_dtype=_lib.BF16, engine=_lib.ENGINE_SIMT)
    dx = torch.full(x.shape, float('nan'), device=DEV, dtype=B)
    ops.conv3d(ctx, dd, [dyd], [wd], None, [dx])
    torch.cuda.synchronize()
    _close(dx, x.grad, 1e-2, 'head dx')


def test_batched_weight_repack_equals_per_pack(ctx):
    """m1_pack_plan (every operand pack of the model in ONE launch) writes exactly what the per-pack launches write"""
    from m1b200 import ops
    model, cfg, x, y = _build('fp16')
    model.set_noise(None, seed=3)
    model.train_step(x, y)                       # creates the packs (fp16 forward packs, bf16 data-gradient packs)
    eng = model.eng
    assert len(eng.packs) > 100
    dtypes = {pk.dtype for _, _, pk in eng.packs.values()}
    assert dtypes == {torch.float16, torch.bfloat16}, dtypes
    for _, _, pk in eng.packs.values():
        pk.zero_()
    eng.refresh_packs()
    assert eng._pack_plan is not None and eng._pack_plan.n == len(eng.packs)
    torch.cuda.synchronize()
    batched = [pk.clone() for _, _, pk in eng.packs.values()]
    for (d, ws, pk), ref in zip(eng.packs.values(), batched):
        pk.zero_()
        ops.conv3d_pack_weights_into(ctx, d, ws, pk)
        torch.cuda.synchronize()
        assert torch.equal(pk.view(torch.int16), ref.view(torch.int16))


# ---- determinism of the forward reductions ---------------------------------------------------------------------
def test_reductions_are_bit_deterministic(ctx):
    """InstanceNorm statistics / SE pooling / loss sums: per-warp slots, per-block partials and a last-block
    finalise in a fixed order - no floating-point atomics on any value the forward pass depends on."""
    from m1b200 import ops
    g = _gen(5)
    for C, shape in ((32, (2, 6, 40, 40)), (8, (2, 6, 40, 40)), (256, (3, 4, 10, 10)), (24, (1, 5, 9, 11))):
        x = torch.randn((*shape, C), generator=g).to(DEV, H)
        ref = None
        for _ in range(4):
            st = torch.empty((shape[0], C, 2), device=DEV)
            ops.inorm_stats(ctx, x, st)
            torch.cuda.synchronize()
            ref = st.clone() if ref is None else ref
            assert torch.equal(st, ref), "instance-norm statistics differ between two runs"
        mean = x.float().mean(dim=(1, 2, 3))
        assert (ref[..., 0] - mean).abs().max().item() < 1e-4


STRIDES = ((1, 1, 1), (1, 2, 2), (1, 2, 2), (2, 2, 2), (2, 2, 2))
KERNELS = ((1, 3, 3), (1, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3))
MID = dict(filters=(32, 64, 128, 192, 256), se_reduction=(8, 8, 8, 8, 8))


def _build(precision, dims=(8, 32, 32), batch=2, xseed=11, **extra):
    from m1b200.model import losses, optimizers, unets
    kw = dict(strides=STRIDES, kernel_sizes=KERNELS, att_sub_samp=((1, 1, 1),) * 4, dropout_rate=0.5,
              dropout_mode='monte-carlo', dense_skip=True, deep_supervision=True, probabilistic=True,
              prob_latent_dims=(3, 2, 1, 0), **MID)
    model = unets.networks.M1(dims, 4, 2, summary=False, precision=precision, seed=0, **kw, **extra)
    model.compile(optimizer=optimizers.Adam(1e-3, amsgrad=True),
                  loss=[losses.Focal(alpha=[0.75, 0.25], gamma=2.0).loss, losses.EvidenceLowerBound().loss],
                  loss_weights=[1.0, 10.0])
    cfg = O.default_config(num_classes=2, dropout_rate=0.5, dropout_mode='monte-carlo', strides=STRIDES,
                           kernel_sizes=KERNELS, dense_skip=True, deep_supervision=True, probabilistic=True,
                           prob_latent_dims=(3, 2, 1, 0), **MID)
    x, y = O.synthetic_batch(batch, dims, probabilistic=True, seed=xseed)
    return model, cfg, x, y


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_forward_pass_is_bit_deterministic(ctx, precision):
    """Two eager training steps from the same weights and Philox step: identical softmax, focal and KL, bit for
    bit (the gradients still carry split-K fp32 atomics and are compared with a tolerance)."""
    model, cfg, x, y = _build(precision)
    model.set_noise(None, seed=7)
    outs = []
    for _ in range(3):
        model.noise.step = 5
        r = model.train_step(x, y, apply_update=False)
        torch.cuda.synchronize()
        outs.append((r['detection'].clone(), r['focal'].clone(), r['kl'].clone(),
                     {k: v.clone() for k, v in model.gradients().items()}))
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]), "softmax differs between two runs"
        assert torch.equal(o[1], outs[0][1]) and torch.equal(o[2], outs[0][2]), "losses differ between two runs"
    a = torch.cat([v.flatten() for v in outs[0][3].values()])
    b = torch.cat([v.flatten() for v in outs[1][3].values()])
    # the weight gradients still sum their split-K partials with fp32 atomics: close, not identical
    assert (a - b).norm().item() <= 2e-2 * a.norm().item()


# ---- whole model, fp16 mode, north-star bounds -----------------------------------------------------------------
def _perturb(ps, seed=17):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, t in ps.p.items():
            if ps.kind[n] in ('gamma', 'beta', 'se_bias'):
                t.add_(0.2 * torch.randn(t.shape, generator=g).to(t.dtype))


def _oracle(cfg, x, y, pseed, nseed):
    ps = O.ParamStore(dtype=torch.float32, seed=pseed, requires_grad=True)
    with torch.no_grad():
        O.train_loss(ps, cfg, x, y, O.Noise(0, torch.float32))
    _perturb(ps)
    noise = O.Noise(nseed, torch.float32)
    r = O.train_loss(ps, cfg, x, y, noise, alpha=(0.75, 0.25), gamma=2.0, kl_weight=10.0)
    (r['detection_loss'] + 10.0 * r['KL_loss']).backward()
    return ps, noise, r


@pytest.mark.parametrize("xseed,pseed,nseed", [(11, 3, 5), (12, 4, 6), (13, 5, 7)])
def test_probabilistic_train_step_fp16_north_star_bounds(ctx, xseed, pseed, nseed):
    """precision='fp16' (the benchmarked mode) against the fp32 oracle on IDENTICAL fp32 inputs, weights and
    injected dropout / latent noise: per-voxel softmax within 2e-2 abs (max, not a percentile), focal and KL
    within 1e-3 relative, gradient cosine >= 0.98 over all parameters and >= 0.95 for the prior and the posterior
    net separately."""
    model, cfg, x, y = _build('fp16', xseed=xseed)
    ps, noise, r = _oracle(cfg, x, y, pseed, nseed)
    before = ctx.launch_count()
    model.set_weights({n: t.detach().float().numpy() for n, t in ps.p.items()})
    model.set_noise(noise.t)
    out = model.train_step(x, y, apply_update=False)
    torch.cuda.synchronize()
    assert ctx.launch_count() > before and len(model.eng.packs) > 0, "no convolution took the tcgen05 engine"
    e = (out['detection'].double().cpu() - r['detection'].detach().double()).abs().flatten()
    fl, fl_ref = out['focal'].item(), r['detection_loss'].item()
    kl, kl_ref = out['kl'].item(), r['KL'].item()
    grads = model.gradients()

    def cos(prefix):
        names = [n for n in ps.p if n.startswith(prefix)]
        a = torch.cat([grads[n].double().flatten() for n in names])
        b = torch.cat([(ps.p[n].grad if ps.p[n].grad is not None else torch.zeros_like(ps.p[n])).double().flatten()
                       for n in names])
        return (a @ b).item() / (a.norm().item() * b.norm().item())
    c_all, c_prior, c_post = cos(''), cos('prior/'), cos('posterior/')
    print(f'fp16: softmax abs err mean {e.mean().item():.2e} max {e.max().item():.2e} | focal rel '
          f'{abs(fl - fl_ref) / abs(fl_ref):.2e} | KL rel {abs(kl - kl_ref) / abs(kl_ref):.2e} | grad cosine '
          f'{c_all:.4f} (prior {c_prior:.4f}, posterior {c_post:.4f})')
    assert e.max().item() < 2e-2, e.max().item()
    assert abs(fl - fl_ref) < 1e-3 * abs(fl_ref), (fl, fl_ref)
    assert abs(kl - kl_ref) < 1e-3 * abs(kl_ref), (kl, kl_ref)
    assert c_all >= 0.98, c_all
    assert c_prior >= 0.95 and c_post >= 0.95, (c_prior, c_post)
