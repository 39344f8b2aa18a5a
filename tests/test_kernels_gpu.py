"""K4-K9 parity: each fused bandwidth kernel (forward AND backward) against autograd of the fp64
CPU oracle on the same seeded inputs."""
import pytest
import torch

from oracle import m1_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _close(got, ref, tol, what=''):
    got = got.detach().double().cpu()
    ref = ref.detach().double()
    scale = max(1.0, ref.abs().max().item())
    err = (got - ref).abs().max().item() / scale
    assert err < tol, f'{what}: rel err {err:.3e} (tol {tol})'


def _gen(seed=0):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("C,slope", [(8, 0.1), (6, 1.0), (64, 0.1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_inorm_act_fwd_bwd(ctx, C, slope, dtype):
    from m1b200 import ops
    g = _gen(1)
    shape = (2, 3, 10, 12, C)
    x = (torch.randn(shape, generator=g) * 2 + 0.5)
    if dtype == torch.bfloat16:
        x = x.bfloat16().float()
    gamma = torch.rand(C, generator=g) + 0.5
    beta = torch.randn(C, generator=g) * 0.3
    dy = torch.randn(shape, generator=g)
    if dtype == torch.bfloat16:
        dy = dy.bfloat16().float()
    xr = x.double().requires_grad_()
    gr, br = gamma.double().requires_grad_(), beta.double().requires_grad_()
    y_ref = O.lrelu(O.instance_norm(xr, gr, br), slope)
    y_ref.backward(dy.double())

    xd = x.to(DEV, dtype)
    stats = torch.empty((2, C, 2), device=DEV)
    y = torch.empty_like(xd)
    ops.inorm_stats(ctx, xd, stats)
    ops.inorm_act_fwd(ctx, xd, stats, gamma.to(DEV), beta.to(DEV), slope, y)
    tol = 2e-5 if dtype == torch.float32 else 1e-2
    _close(y, y_ref, tol, 'y')
    _close(stats[..., 0], xr.mean(dim=(1, 2, 3)), 1e-5, 'mean')
    dx = torch.empty_like(xd)
    dgam = torch.zeros(C, device=DEV)
    dbet = torch.zeros(C, device=DEV)
    ops.inorm_act_bwd(ctx, dy.to(DEV, dtype), xd, stats, gamma.to(DEV), beta.to(DEV), slope, dx, False, dgam, dbet)
    torch.cuda.synchronize()
    _close(dx, xr.grad, 1e-4 if dtype == torch.float32 else 2e-2, 'dx')
    _close(dgam, gr.grad, 1e-4 if dtype == torch.float32 else 2e-2, 'dgamma')
    _close(dbet, br.grad, 1e-4 if dtype == torch.float32 else 2e-2, 'dbeta')


@pytest.mark.parametrize("C,red,rate", [(16, 4, 0.5), (32, 8, 0.0), (12, 4, 0.25)])
def test_se_tail_fwd_bwd(ctx, C, red, rate):
    """GAP -> conv6 -> lrelu -> conv7 -> sigmoid -> x_*g*res -> lrelu -> dropout, with norm3/norm4 fused."""
    from m1b200 import ops
    g = _gen(2)
    shape = (2, 3, 6, 8, C)
    Cr = C // red
    raw3 = torch.randn(shape, generator=g)
    raw4 = torch.randn(shape, generator=g) * 1.5 + 0.2
    P = {k: v for k, v in dict(
        g3=torch.rand(C, generator=g) + 0.5, b3=torch.randn(C, generator=g) * 0.5,
        g4=torch.rand(C, generator=g) + 0.5, b4=torch.randn(C, generator=g) * 0.5,
        w6=torch.randn((C, Cr), generator=g) * 0.3, b6=torch.randn(Cr, generator=g) * 0.1,
        w7=torch.randn((Cr, C), generator=g) * 0.3, b7=torch.randn(C, generator=g) * 0.1).items()}
    u = torch.rand(shape, generator=g)
    dout = torch.randn(shape, generator=g)

    R = {k: v.double().requires_grad_() for k, v in P.items()}
    r3, r4 = raw3.double().requires_grad_(), raw4.double().requires_grad_()
    x_ = O.instance_norm(r3, R['g3'], R['b3'])
    res = O.instance_norm(r4, R['g4'], R['b4'])
    pool = x_.mean(dim=(1, 2, 3))
    gate = torch.sigmoid(O.lrelu(pool @ R['w6'] + R['b6']) @ R['w7'] + R['b7'])
    out_ref = O.dropout(O.lrelu(x_ * gate[:, None, None, None, :] * res), rate, u.double())
    out_ref.backward(dout.double())

    D = {k: v.to(DEV) for k, v in P.items()}
    d3, d4 = raw3.to(DEV), raw4.to(DEV)
    st3 = torch.empty((2, C, 2), device=DEV)
    st4 = torch.empty((2, C, 2), device=DEV)
    ops.inorm_stats(ctx, d3, st3)
    ops.inorm_stats(ctx, d4, st4)
    poold = torch.empty((2, C), device=DEV)
    hidden = torch.empty((2, Cr), device=DEV)
    gated = torch.empty((2, C), device=DEV)
    ops.se_squeeze(ctx, d3, st3, D['g3'], D['b3'], poold)
    ops.se_excite_fwd(ctx, poold, D['w6'], D['b6'], D['w7'], D['b7'], hidden, gated)
    ud = u.to(DEV)                      # keep alive: m1_dropout only holds the raw pointer
    drop = ops.make_dropout(rate, ud)
    out = torch.empty_like(d3)
    ops.se_gate_fwd(ctx, d3, d4, st3, st4, D['g3'], D['b3'], D['g4'], D['b4'], gated, drop, out)
    _close(poold, pool, 1e-5, 'pool')
    _close(poold, P['b3'].expand(2, C), 1e-5, 'Q6: GAP(IN(x)) == beta')
    _close(gated, gate, 1e-5, 'gate')
    _close(out, out_ref, 1e-5, 'out')

    red5 = torch.empty((2, C, 5), device=DEV)
    dgate = torch.empty((2, C), device=DEV)
    dd = dout.to(DEV)
    ops.se_gate_bwd_reduce(ctx, dd, d3, d4, st3, st4, D['g3'], D['b3'], D['g4'], D['b4'], gated, drop, red5, dgate)
    G = {k: torch.zeros_like(v) for k, v in D.items()}
    dpool = torch.empty((2, C), device=DEV)
    ops.se_excite_bwd(ctx, dgate, poold, hidden, gated, D['w6'], D['w7'], dpool, G['w6'], G['b6'], G['w7'], G['b7'])
    dr3, dr4 = torch.empty_like(d3), torch.empty_like(d4)
    ops.se_gate_bwd_apply(ctx, dd, d3, d4, st3, st4, D['g3'], D['b3'], D['g4'], D['b4'], gated, drop, red5, dpool,
                          dr3, dr4, G['g3'], G['b3'], G['g4'], G['b4'])
    torch.cuda.synchronize()
    _close(dr3, r3.grad, 1e-4, 'draw3')
    _close(dr4, r4.grad, 1e-4, 'draw4')
    for k in P:
        _close(G[k], R[k].grad, 2e-4, 'd' + k)


@pytest.mark.parametrize("sub,F,xg,gg", [((1, 1, 1), 12, (4, 8, 8), (1, 2, 2)), ((2, 2, 2), 12, (4, 8, 8), (1, 2, 2)),
                                         ((1, 1, 1), 16, (4, 8, 8), (1, 2, 2)), ((2, 2, 2), 16, (4, 8, 8), (1, 2, 2)),
                                         ((1, 1, 1), 32, (4, 16, 16), (1, 1, 1)), ((1, 1, 1), 64, (4, 16, 32), (2, 2, 4)),
                                         ((1, 1, 1), 256, (2, 4, 4), (1, 2, 2))])
def test_attention_gate_fwd_bwd(ctx, sub, F, xg, gg):
    from m1b200 import ops
    g = _gen(3)
    Cx = F
    tg = tuple(a // s for a, s in zip(xg, sub))
    x = torch.randn((2, *xg, Cx), generator=g)
    theta = torch.randn((2, *tg, F), generator=g)
    phi = torch.randn((2, *gg, F), generator=g)
    wpsi = torch.randn(F, generator=g) * 0.5
    bpsi = torch.randn(1, generator=g) * 0.1
    dy = torch.randn(x.shape, generator=g)
    R = [t.double().requires_grad_() for t in (x, theta, phi, wpsi, bpsi)]
    xr, tr, pr, wr, br = R
    up = O.upsample_nearest(pr, [tg[i] // gg[i] for i in range(3)])
    f = O.lrelu(tr + up)
    psi_ref = torch.sigmoid((f * wr).sum(-1, keepdim=True) + br)
    y_ref = O.upsample_nearest(psi_ref, sub) * xr
    y_ref.backward(dy.double())

    xd, td, pd_ = x.to(DEV), theta.to(DEV), phi.to(DEV)
    wd, bd = wpsi.to(DEV), bpsi.to(DEV)
    psi = torch.empty((2, *tg), device=DEV)
    y = torch.empty_like(xd)
    ops.attn_fwd(ctx, td, pd_, wd, bd, xd, psi, y)
    _close(psi, psi_ref[..., 0], 1e-5, 'psi')
    _close(y, y_ref, 1e-5, 'y')
    dx = torch.empty_like(xd)
    dth = torch.empty_like(td)
    dphi = torch.zeros(pd_.shape, device=DEV)
    dw = torch.zeros(F, device=DEV)
    db = torch.zeros(1, device=DEV)
    ops.attn_bwd(ctx, dy.to(DEV), td, pd_, wd, psi, xd, dx, False, dth, dphi, dw, db)
    torch.cuda.synchronize()
    _close(dx, xr.grad, 1e-5, 'dx')
    _close(dth, tr.grad, 1e-4, 'dtheta')
    _close(dphi, pr.grad, 1e-4, 'dphi')
    _close(dw, wr.grad, 1e-4, 'dw_psi')
    _close(db, br.grad, 1e-4, 'db_psi')


@pytest.mark.parametrize("L", [3, 1])
def test_latent_and_kl(ctx, L):
    from m1b200 import ops
    g = _gen(4)
    shp = (2, 3, 4, 5)
    mlq = torch.randn((*shp, 2 * L), generator=g) * 0.2
    mlp = torch.randn((*shp, 2 * L), generator=g) * 0.2
    eps = torch.randn((*shp, L), generator=g)
    dz = torch.randn((*shp, L), generator=g)
    q, p = mlq.double().requires_grad_(), mlp.double().requires_grad_()
    sq = torch.exp(torch.clamp(q[..., L:], -0.1, 0.1))
    sp = torch.exp(torch.clamp(p[..., L:], -0.1, 0.1))
    z_ref = q[..., :L] + sq * eps.double()
    kl_ref = O.kl_mvn_diag((q[..., :L], sq), (p[..., :L], sp)).sum(dim=(1, 2, 3)).mean()
    (10.0 * kl_ref + (z_ref * dz.double()).sum()).backward()

    qd, pd_, ed = mlq.to(DEV), mlp.to(DEV), eps.to(DEV)
    z = torch.empty((*shp, L), device=DEV)
    ops.latent_fwd(ctx, qd, ed, 0, z)
    _close(z, z_ref, 1e-5, 'z')
    zm = torch.empty((*shp, L), device=DEV)
    ops.latent_fwd(ctx, qd, None, 1, zm)
    _close(zm, q[..., :L], 1e-6, 'z=mu')
    kl = torch.zeros(1, device=DEV)
    ops.kl_fwd(ctx, qd, pd_, kl)
    _close(kl[0], kl_ref, 1e-5, 'kl')
    klz = torch.zeros(1, device=DEV)
    ops.kl_fwd(ctx, qd, qd, klz)
    assert abs(klz.item()) < 1e-5
    dq, dp = torch.zeros_like(qd), torch.zeros_like(pd_)
    ops.kl_bwd(ctx, qd, pd_, 10.0, dq, dp)
    ops.latent_bwd(ctx, dz.to(DEV), qd, ed, 0, dq)
    torch.cuda.synchronize()
    _close(dq, q.grad, 1e-4, 'dml_q')
    _close(dp, p.grad, 1e-4, 'dml_p')


@pytest.mark.parametrize("up,gamma", [((1, 1, 1), 2.0), ((1, 2, 2), 2.0), ((2, 4, 4), 0.0)])
def test_softmax_focal(ctx, up, gamma):
    from m1b200 import ops
    g = _gen(5)
    lg = (2, 3, 4)
    xg = tuple(a * b for a, b in zip(lg, up))
    logits = torch.randn((2, *lg, 2), generator=g) * 2
    lab = (torch.rand((2, *xg), generator=g) > 0.8).long()
    y = torch.nn.functional.one_hot(lab, 2).float()
    alpha = (0.75, 0.25)
    lr = logits.double().requires_grad_()
    p_ref = torch.softmax(O.upsample_nearest(lr, up), -1)
    loss_ref = O.focal_loss(y.double(), p_ref, alpha, gamma)
    (3.0 * loss_ref).backward()
    ld = logits.to(DEV)
    prob = torch.zeros((2, *xg, 4), device=DEV)
    loss = torch.zeros(1, device=DEV)
    dl = torch.zeros_like(ld)
    ops.softmax_focal(ctx, ld, y.to(DEV), alpha, gamma, up, prob, 2, 1.0, loss, dl, 3.0)
    torch.cuda.synchronize()
    _close(prob[..., 2:], p_ref, 1e-5, 'softmax')
    assert prob[..., :2].abs().max().item() == 0.0
    _close(loss[0], loss_ref, 1e-4, 'loss')
    _close(dl, lr.grad, 1e-4, 'dlogits')


def test_adam_amsgrad_and_l2(ctx):
    from m1b200 import ops
    g = _gen(6)
    n = 1003
    w = torch.randn(n, generator=g)
    m = torch.zeros(n)
    v = torch.zeros(n)
    vh = torch.zeros(n)
    wd, md, vd, hd = (t.to(DEV).clone() for t in (w, m, v, vh))
    wr, mr, vr, hr = (t.double() for t in (w, m, v, vh))
    l2 = 1e-4
    import math
    for step in (1, 2, 3):
        gr = torch.randn(n, generator=g)
        lr_t = 1e-3 * math.sqrt(1 - 0.999 ** step) / (1 - 0.9 ** step)
        l2acc = torch.zeros(1, device=DEV)
        ops.adam_amsgrad(ctx, wd, gr.to(DEV), md, vd, hd, lr_t, 0.9, 0.999, 1e-7, l2, 0.5, l2acc)
        l2_ref = l2 * (wr ** 2).sum()
        wr, mr, vr, hr = O.adam_amsgrad_step(wr, gr.double() * 0.5 + 2 * l2 * wr, mr, vr, hr, step, 1e-3)
        torch.cuda.synchronize()
        _close(l2acc[0], l2_ref, 1e-5, 'l2 term')
        _close(wd, wr, 1e-5, f'w step {step}')
        _close(hd, hr, 1e-5, f'vhat step {step}')


def test_philox_dropout_statistics_and_replay(ctx):
    """Philox masks: keep-rate ~ 1-rate, E[y] = x, and the backward regenerates the same mask."""
    from m1b200 import ops
    C = 32
    shape = (2, 8, 16, 16, C)
    raw = torch.ones(shape, device=DEV)
    st = torch.zeros((2, C, 2), device=DEV)
    st[..., 1] = 1.0                                   # mean 0, rstd 1 -> identity norm
    one, zero = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    gate = torch.ones((2, C), device=DEV)
    drop = ops.make_dropout(0.5, None, seed=42, stream_id=7)
    out = torch.empty_like(raw)
    ops.se_gate_fwd(ctx, raw, raw, st, st, one, zero, one, zero, gate, drop, out)
    out2 = torch.empty_like(raw)
    ops.se_gate_fwd(ctx, raw, raw, st, st, one, zero, one, zero, gate, drop, out2)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    keep = (out != 0).float().mean().item()
    assert abs(keep - 0.5) < 0.01
    assert set(out.unique().tolist()) == {0.0, 2.0}
    assert abs(out.mean().item() - 1.0) < 0.02
    drop_b = ops.make_dropout(0.5, None, seed=42, stream_id=8)
    ops.se_gate_fwd(ctx, raw, raw, st, st, one, zero, one, zero, gate, drop_b, out2)
    torch.cuda.synchronize()
    assert not torch.equal(out, out2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dropout_keep_mask_equals_regenerated_noise(ctx, dtype):
    """m1_dropout.mask: the keep-bits written by se_gate_fwd give the backward kernels exactly the mask that
    re-running Philox gives them (same draw3/draw4/red), and bit i of byte e/8 is 'element e kept'."""
    from m1b200 import ops
    g = _gen(11)
    C, shape = 32, (2, 4, 8, 8, 32)
    n = 2
    raw3 = torch.randn(shape, generator=g).to(DEV, dtype)
    raw4 = torch.randn(shape, generator=g).to(DEV, dtype)
    dout = torch.randn(shape, generator=g).to(DEV, dtype)
    st3, st4 = torch.zeros((n, C, 2), device=DEV), torch.zeros((n, C, 2), device=DEV)
    ops.inorm_stats(ctx, raw3, st3, 1e-3)
    ops.inorm_stats(ctx, raw4, st4, 1e-3)
    g3, b3 = torch.rand(C, generator=g).to(DEV) + 0.5, torch.randn(C, generator=g).to(DEV)
    g4, b4 = torch.rand(C, generator=g).to(DEV) + 0.5, torch.randn(C, generator=g).to(DEV)
    gate = torch.rand((n, C), generator=g).to(DEV)
    dpool = torch.zeros((n, C), device=DEV)
    res = []
    for use_mask in (False, True):
        mask = torch.zeros(raw3.numel() // 8, dtype=torch.uint8, device=DEV) if use_mask else None
        drop = ops.make_dropout(0.5, None, seed=5, stream_id=9, mask=mask)
        out = torch.empty_like(raw3)
        ops.se_gate_fwd(ctx, raw3, raw4, st3, st4, g3, b3, g4, b4, gate, drop, out)
        red, dgate = torch.zeros((n, C, 5), device=DEV), torch.zeros((n, C), device=DEV)
        ops.se_gate_bwd_reduce(ctx, dout, raw3, raw4, st3, st4, g3, b3, g4, b4, gate, drop, red, dgate)
        d3, d4 = torch.empty_like(raw3), torch.empty_like(raw3)
        z = [torch.zeros(C, device=DEV) for _ in range(4)]
        ops.se_gate_bwd_apply(ctx, dout, raw3, raw4, st3, st4, g3, b3, g4, b4, gate, drop, red, dpool, d3, d4, *z)
        torch.cuda.synchronize()
        res.append((out, red, d3, d4, mask))
    (o0, r0, a0, c0, _), (o1, r1, a1, c1, mask) = res
    assert torch.equal(o0, o1)
    assert (r0 - r1).abs().max().item() <= 1e-4 * r0.abs().max().item()      # atomics: summation order only
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-4                           # draw* depend on red
    for p_, q_ in ((a0, a1), (c0, c1)):
        assert (p_.float() - q_.float()).abs().max().item() <= tol * max(1.0, p_.float().abs().max().item())
        # a mismatching mask bit would flip whole elements: count elements that differ by more than rounding
        assert ((p_.float() - q_.float()).abs() > 0.05 * p_.float().abs().clamp_min(0.1)).sum().item() == 0
    bits = ((mask.view(-1, 1) >> torch.arange(8, device=DEV, dtype=torch.uint8)) & 1).bool().view(-1)
    kept = (o1.float().view(-1) != 0)
    z_nonzero = o1.float().view(-1) != 0
    assert torch.equal(bits[z_nonzero], torch.ones_like(bits[z_nonzero]))     # every surviving output was kept
    assert abs(bits.float().mean().item() - 0.5) < 0.02 and kept.any()


def test_utilities(ctx):
    from m1b200 import ops
    g = _gen(7)
    x = torch.randn((2, 3, 4, 5, 4), generator=g)
    xd = x.to(DEV)
    img = torch.empty((2, 3, 4, 5, 3), device=DEV, dtype=torch.bfloat16)
    ops.copy_channels(ctx, xd, 0, img, 0, 3)
    lab = torch.empty((2, 3, 4, 5, 1), device=DEV)
    ops.copy_channels(ctx, xd, 2, lab, 0, 1)            # Q4 slice: channel 2
    torch.cuda.synchronize()
    assert torch.equal(img.float().cpu(), x[..., :3].bfloat16().float())
    assert torch.equal(lab.cpu(), x[..., 2:3])
    a, b = torch.rand(100, generator=g), torch.rand(100, generator=g)
    for i, strat in enumerate(('identity', 'noisy-or', 'bayes')):
        out = torch.empty((100, 2), device=DEV)
        ops.decision_fusion(ctx, a.to(DEV), b.to(DEV), i, out)
        torch.cuda.synchronize()
        _, ref = O.decision_fusion(a.double(), b.double(), strat)
        _close(out, ref, 1e-6, strat)
