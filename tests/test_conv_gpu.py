"""K1/K2/K3 parity: the CUDA convolution engines against the CPU oracle (tests may import oracle/)."""
import ctypes

import pytest
import torch

from oracle import m1_oracle as O

pytestmark = pytest.mark.gpu


def _mk(batch, dhw, cins, couts, k, seed=0):
    g = torch.Generator().manual_seed(seed)
    xs = [torch.randn((batch, *dhw, c), generator=g) for c in cins]
    cin = sum(cins)
    ws = [torch.randn((*k, cin, co), generator=g) / (cin * k[0] * k[1] * k[2]) ** 0.5 for co in couts]
    bs = [torch.randn((co,), generator=g) * 0.1 for co in couts]
    return xs, ws, bs


def _run_fwd(ctx, xs, ws, bs, k, s, dtype, engine, variant=0):
    from m1b200 import ops, _lib
    dev = 'cuda'
    batch, dhw = xs[0].shape[0], xs[0].shape[1:4]
    geo = [ops.same_pads(dhw[i], k[i], s[i]) for i in range(3)]
    out_dhw = [g[0] for g in geo]
    pad = [g[1] for g in geo]
    cin = sum(x.shape[-1] for x in xs)
    code = _lib.BF16 if dtype == torch.bfloat16 else _lib.F32
    d = ops.conv_desc(_lib.CONV_FWD, batch, dhw, out_dhw, k, s, pad, [x.shape[-1] for x in xs],
                      [w.shape[-1] for w in ws], [(cin * w.shape[-1], w.shape[-1], 1) for w in ws],
                      act_dtype=code, engine=engine)
    d.tune[0] = variant
    if variant == 2:
        assert _lib.lib().m1_conv3d_halo_engine(ctypes.byref(d)) == 1, "halo engine refused the launch"
    xd = [x.to(dev, dtype).contiguous() for x in xs]
    wd = [w.to(dev).contiguous() for w in ws]
    bd = [b.to(dev).contiguous() for b in bs]
    outs = [torch.full((batch, *out_dhw, w.shape[-1]), float('nan'), device=dev, dtype=dtype) for w in ws]
    packed = None
    if engine == _lib.ENGINE_TCGEN05:
        assert ops.conv3d_tc_supported(d)
        packed = ops.conv3d_pack_weights(ctx, d, wd)
    ops.conv3d(ctx, d, xd, wd, bd, outs, packed)
    torch.cuda.synchronize()
    return [o.float().cpu() for o in outs]


def _ref_fwd(xs, ws, bs, s, dtype):
    x = torch.cat(xs, -1)
    if dtype == torch.bfloat16:  # the engine sees bf16-rounded operands
        x = x.bfloat16().float()
        ws = [w.bfloat16().float() for w in ws]
    return [O.conv3d_same(x.double(), w.double(), b.double(), s).float() for w, b in zip(ws, bs)]


SIMT_CASES = [
    # dhw, cins, couts, kernel, stride
    ((4, 10, 12), [3], [8], (1, 3, 3), (1, 1, 1)),
    ((4, 10, 12), [5, 7], [6, 9], (3, 3, 3), (1, 1, 1)),
    ((4, 10, 12), [8], [16], (1, 3, 3), (1, 2, 2)),
    ((6, 10, 12), [8], [16], (3, 3, 3), (2, 2, 2)),
    ((5, 9, 11), [4], [4], (3, 3, 3), (2, 2, 2)),       # odd sizes: SAME pads (1,1)
    ((4, 8, 8), [6], [6], (2, 2, 2), (2, 2, 2)),        # attention theta conv with sub-sampling
    ((4, 8, 8), [16], [2], (1, 1, 1), (1, 1, 1)),
]


@pytest.mark.parametrize("dhw,cins,couts,k,s", SIMT_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv_fwd_simt(ctx, dhw, cins, couts, k, s, dtype):
    from m1b200 import _lib
    xs, ws, bs = _mk(2, dhw, cins, couts, k)
    got = _run_fwd(ctx, xs, ws, bs, k, s, dtype, _lib.ENGINE_SIMT)
    ref = _ref_fwd(xs, ws, bs, s, dtype)
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    for g, r in zip(got, ref):
        assert g.shape == r.shape
        assert torch.isfinite(g).all()
        assert (g - r).abs().max().item() < tol, (g - r).abs().max().item()


TC_CASES = [
    ((4, 16, 16), [64], [32], (1, 3, 3)),               # one source, SW128, exact bricks
    ((4, 16, 16), [64], [16, 64], (3, 3, 3)),           # fused conv1||conv4 split
    ((6, 20, 20), [128, 64], [32, 128], (3, 3, 3)),     # virtual concat, ragged bricks (20 % 8 != 0)
    ((4, 16, 32), [32, 32, 32], [8, 32], (1, 3, 3)),    # 32-channel sources -> SW64 k-steps
    ((4, 16, 16), [16, 48], [16], (3, 3, 3)),           # 16-channel granularity -> SW32
    ((5, 10, 10), [128], [64, 256], (3, 3, 3)),         # N = 320 -> two N tiles; res4-like tiny grid
    ((2, 8, 16), [64], [64], (1, 1, 1)),
]


@pytest.mark.parametrize("dhw,cins,couts,k", TC_CASES)
def test_conv_fwd_tcgen05(ctx, dhw, cins, couts, k):
    from m1b200 import _lib
    xs, ws, bs = _mk(2, dhw, cins, couts, k, seed=1)
    got = _run_fwd(ctx, xs, ws, bs, k, (1, 1, 1), torch.bfloat16, _lib.ENGINE_TCGEN05)
    ref = _ref_fwd(xs, ws, bs, (1, 1, 1), torch.bfloat16)
    for g, r in zip(got, ref):
        assert torch.isfinite(g).all(), "non-finite output (unwritten voxels?)"
        err = (g - r).abs().max().item()
        assert err < 2e-2, err


def _run_transposed(ctx, x, w, b, k, s, dtype, engine=None):
    """Conv3DTranspose forward: gather mode TRANSPOSED, Keras kernel (kd,kh,kw,Cout,Cin)."""
    from m1b200 import ops, _lib
    dev = 'cuda'
    batch, in_dhw = x.shape[0], x.shape[1:4]
    out_dhw = [in_dhw[i] * s[i] for i in range(3)]
    pad = [ops.same_pads(out_dhw[i], k[i], s[i])[1] for i in range(3)]
    cout, cin = w.shape[-2], w.shape[-1]
    code = _lib.BF16 if dtype == torch.bfloat16 else _lib.F32
    d = ops.conv_desc(_lib.CONV_TRANSPOSED, batch, in_dhw, out_dhw, k, s, pad, [cin], [cout],
                      [(cout * cin, 1, cin)], act_dtype=code,
                      engine=_lib.ENGINE_SIMT if engine is None else engine)
    out = torch.full((batch, *out_dhw, cout), float('nan'), device=dev, dtype=dtype)
    ops.conv3d(ctx, d, [x.to(dev, dtype).contiguous()], [w.to(dev).contiguous()],
               [b.to(dev).contiguous()], [out])
    torch.cuda.synchronize()
    return out.float().cpu()


@pytest.mark.parametrize("dhw,k,s", [((3, 5, 6), (3, 3, 3), (2, 2, 2)), ((4, 5, 6), (3, 3, 3), (1, 2, 2)),
                                     ((4, 5, 6), (1, 3, 3), (1, 2, 2)), ((3, 4, 4), (3, 3, 3), (1, 1, 1))])
def test_conv_transpose_simt(ctx, dhw, k, s):
    g = torch.Generator().manual_seed(3)
    cin, cout = 6, 5
    x = torch.randn((2, *dhw, cin), generator=g)
    w = torch.randn((*k, cout, cin), generator=g) * 0.2
    b = torch.randn((cout,), generator=g) * 0.1
    got = _run_transposed(ctx, x, w, b, k, s, torch.float32)
    ref = O.conv3d_transpose_same(x.double(), w.double(), b.double(), s).float()
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 1e-4


@pytest.mark.parametrize("k,s", [((3, 3, 3), (1, 1, 1)), ((3, 3, 3), (2, 2, 2)), ((1, 3, 3), (1, 2, 2))])
def test_conv_wgrad_dgrad_simt(ctx, k, s):
    """dgrad (TRANSPOSED gather of dy) and wgrad against autograd of the oracle conv."""
    from m1b200 import ops, _lib
    g = torch.Generator().manual_seed(4)
    dhw, cins, cout = (4, 6, 8), [5, 4], 7
    xs = [torch.randn((2, *dhw, c), generator=g, dtype=torch.float64, requires_grad=True) for c in cins]
    cin = sum(cins)
    w = (torch.randn((*k, cin, cout), generator=g, dtype=torch.float64) * 0.2).requires_grad_()
    b = torch.zeros(cout, dtype=torch.float64, requires_grad=True)
    y = O.conv3d_same(torch.cat(xs, -1), w, b, s)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    dev = 'cuda'
    out_dhw = list(y.shape[1:4])
    pad = [ops.same_pads(dhw[i], k[i], s[i])[1] for i in range(3)]
    # wgrad
    d = ops.conv_desc(_lib.CONV_FWD, 2, dhw, out_dhw, k, s, pad, cins, [cout], [(cin * cout, cout, 1)],
                      act_dtype=_lib.F32, engine=_lib.ENGINE_SIMT)
    dw = torch.zeros(w.shape, device=dev)
    db = torch.zeros(cout, device=dev)
    ops.conv3d_wgrad(ctx, d, [x.detach().float().to(dev).contiguous() for x in xs],
                     [dy.float().to(dev).contiguous()], [dw], [db])
    torch.cuda.synchronize()
    assert (dw.cpu().double() - w.grad).abs().max().item() < 1e-3
    assert (db.cpu().double() - b.grad).abs().max().item() < 1e-3
    # dgrad per gathered tensor: produced = that tensor's channels, reduced = cout
    off = 0
    for x in xs:
        c = x.shape[-1]
        dd = ops.conv_desc(_lib.CONV_TRANSPOSED, 2, out_dhw, dhw, k, s, pad, [cout], [c],
                           [(cin * cout, 1, cout)], act_dtype=_lib.F32, engine=_lib.ENGINE_SIMT)
        dx = torch.full(x.shape, float('nan'), device=dev)
        wv = w.detach().float().to(dev).contiguous()
        wslice = wv.view(-1)[off * cout:]  # first reduced... weights of channels [off, off+c)
        ops.conv3d(ctx, dd, [dy.float().to(dev).contiguous()], [wslice], None, [dx])
        torch.cuda.synchronize()
        assert (dx.cpu().double() - x.grad).abs().max().item() < 1e-3
        off += c


@pytest.mark.parametrize("dhw,cins,cout,k", [((4, 16, 16), [64, 32], 64, (3, 3, 3)),
                                             ((6, 20, 20), [128, 128, 64], 32, (1, 3, 3)),
                                             ((4, 16, 16), [32, 32, 32, 32, 32], 32, (1, 3, 3))])
@pytest.mark.parametrize("variant", [0, 3])
def test_conv_dgrad_tcgen05_multi_output(ctx, dhw, cins, cout, k, variant):
    """Data gradient of a stride-1 conv over a virtual concatenation: ONE tcgen05 launch whose produced
    channels are split over the gradients of the concatenated tensors (one pre-loaded: accumulate bit)."""
    from m1b200 import ops, _lib
    g = torch.Generator().manual_seed(8)
    cin = sum(cins)
    xs = [torch.randn((2, *dhw, c), generator=g, dtype=torch.float64, requires_grad=True) for c in cins]
    w = (torch.randn((*k, cin, cout), generator=g) / (cin * 9) ** 0.5).bfloat16().double()
    y = O.conv3d_same(torch.cat(xs, -1), w, None, (1, 1, 1))
    dy = torch.randn(y.shape, generator=g).bfloat16().double()
    y.backward(dy)
    dev = 'cuda'
    pad = [ops.same_pads(dhw[i], k[i], 1)[1] for i in range(3)]
    prior = torch.randn(xs[0].shape, generator=g).bfloat16()          # pre-existing gradient of tensor 0
    bufs = [prior.clone().to(dev)] + [torch.full(x.shape, float('nan'), device=dev, dtype=torch.bfloat16)
                                      for x in xs[1:]]
    wd = w.float().to(dev).contiguous()
    offs = [sum(cins[:i]) for i in range(len(cins))]
    wv = [wd.view(-1)[o * cout:] for o in offs]
    d = ops.conv_desc(_lib.CONV_TRANSPOSED, 2, dhw, dhw, k, (1, 1, 1), pad, [cout], cins,
                      [(cin * cout, 1, cout)] * len(cins), accumulate=[True] + [False] * (len(cins) - 1),
                      act_dtype=_lib.BF16, engine=_lib.ENGINE_TCGEN05)
    assert ops.conv3d_tc_supported(d)
    d.tune[0] = variant
    packed = ops.conv3d_pack_weights(ctx, d, wv)
    ops.conv3d(ctx, d, [dy.to(dev, torch.bfloat16).contiguous()], wv, None, bufs, packed)
    torch.cuda.synchronize()
    for i, (x, b) in enumerate(zip(xs, bufs)):
        ref = x.grad + (prior.double() if i == 0 else 0)
        err = (b.double().cpu() - ref).abs().max().item()
        assert err < 3e-2 * max(1.0, ref.abs().max().item()), (i, err)


@pytest.mark.parametrize("dhw,cins,cout,k", [((4, 16, 16), [64], 64, (3, 3, 3)),
                                             ((4, 16, 16), [64, 32], 32, (1, 3, 3)),
                                             ((6, 20, 20), [128, 128, 64], 128, (3, 3, 3)),
                                             ((4, 16, 32), [32, 32, 32, 32, 32], 32, (1, 3, 3)),
                                             ((5, 10, 10), [256], 512, (3, 3, 3)),
                                             ((4, 16, 16), [16, 48], 16, (3, 3, 3)),
                                             ((2, 8, 16), [64], 160, (1, 1, 1)),
                                             ((4, 16, 16), [16], 16, (3, 3, 3)),      # taps-in-M: 8 taps per M tile
                                             ((4, 16, 16), [32], 64, (1, 3, 3)),      # taps-in-M: 4 taps per M tile
                                             ((6, 12, 16), [16], 32, (1, 3, 3))])
def test_conv_wgrad_tcgen05(ctx, dhw, cins, cout, k):
    """Weight gradient on the tensor cores (MN-major operands straight from NDHWC) vs autograd."""
    from m1b200 import ops, _lib
    g = torch.Generator().manual_seed(9)
    cin = sum(cins)
    xs = [torch.randn((2, *dhw, c), generator=g).bfloat16().double() for c in cins]
    w = torch.zeros((*k, cin, cout), dtype=torch.float64, requires_grad=True)
    y = O.conv3d_same(torch.cat(xs, -1), w, None, (1, 1, 1))
    dy = torch.randn(y.shape, generator=g).bfloat16().double()
    y.backward(dy)
    dev = 'cuda'
    pad = [ops.same_pads(dhw[i], k[i], 1)[1] for i in range(3)]
    d = ops.conv_desc(_lib.CONV_FWD, 2, dhw, dhw, k, (1, 1, 1), pad, cins, [cout], [(cin * cout, cout, 1)],
                      act_dtype=_lib.BF16, engine=_lib.ENGINE_TCGEN05)
    dw = torch.full(w.shape, 0.5, device=dev)                       # accumulates onto existing content
    db = torch.zeros(cout, device=dev)
    ops.conv3d_wgrad(ctx, d, [x.to(dev, torch.bfloat16).contiguous() for x in xs],
                     [dy.to(dev, torch.bfloat16).contiguous()], [dw], [db])
    torch.cuda.synchronize()
    ref = w.grad + 0.5
    scale = max(1.0, ref.abs().max().item())
    err = (dw.double().cpu() - ref).abs().max().item() / scale
    assert err < 2e-3, err
    assert (db.double().cpu() - dy.sum(dim=(0, 1, 2, 3))).abs().max().item() < 1e-2 * scale


STRIDED_TC = [((4, 16, 16), [64], [16, 64], (1, 3, 3), (1, 2, 2)),
              ((8, 16, 16), [64, 32], [32, 128], (3, 3, 3), (2, 2, 2)),
              ((6, 20, 12), [128], [64], (3, 3, 3), (1, 2, 2)),
              ((5, 9, 11), [32], [32], (3, 3, 3), (2, 2, 2))]       # odd sizes: SAME pads (1,1)


@pytest.mark.parametrize("dhw,cins,couts,k,s", STRIDED_TC)
def test_conv_fwd_strided_tcgen05(ctx, dhw, cins, couts, k, s):
    """Strided convolution: the A boxes are loaded with TMA element strides."""
    from m1b200 import _lib
    xs, ws, bs = _mk(2, dhw, cins, couts, k, seed=5)
    got = _run_fwd(ctx, xs, ws, bs, k, s, torch.bfloat16, _lib.ENGINE_TCGEN05)
    ref = _ref_fwd(xs, ws, bs, s, torch.bfloat16)
    for g, r in zip(got, ref):
        assert g.shape == r.shape and torch.isfinite(g).all()
        assert (g - r).abs().max().item() < 2e-2


@pytest.mark.parametrize("dhw,k,s,cin,cout", [((3, 8, 8), (3, 3, 3), (2, 2, 2), 64, 32),
                                              ((4, 8, 8), (3, 3, 3), (1, 2, 2), 128, 64),
                                              ((4, 10, 6), (1, 3, 3), (1, 2, 2), 64, 32),
                                              ((3, 5, 7), (3, 3, 3), (2, 2, 2), 32, 48)])
@pytest.mark.parametrize("variant", [0, 3])
def test_conv_transpose_tcgen05(ctx, dhw, k, s, cin, cout, variant):
    """Conv3DTranspose forward on the tensor cores: one launch, one output phase per blockIdx.z."""
    from m1b200 import _lib
    g = torch.Generator().manual_seed(6)
    x = torch.randn((2, *dhw, cin), generator=g).bfloat16().float()
    w = (torch.randn((*k, cout, cin), generator=g) / (cin * 4) ** 0.5).bfloat16().float()
    b = torch.randn((cout,), generator=g) * 0.1
    from m1b200 import ops
    dev = 'cuda'
    out_dhw = [dhw[i] * s[i] for i in range(3)]
    pad = [ops.same_pads(out_dhw[i], k[i], s[i])[1] for i in range(3)]
    d = ops.conv_desc(_lib.CONV_TRANSPOSED, 2, dhw, out_dhw, k, s, pad, [cin], [cout], [(cout * cin, 1, cin)],
                      act_dtype=_lib.BF16, engine=_lib.ENGINE_TCGEN05)
    assert ops.conv3d_tc_supported(d)
    d.tune[0] = variant
    wd, bd = w.to(dev).contiguous(), b.to(dev).contiguous()
    packed = ops.conv3d_pack_weights(ctx, d, [wd])
    out = torch.full((2, *out_dhw, cout), float('nan'), device=dev, dtype=torch.bfloat16)
    ops.conv3d(ctx, d, [x.to(dev, torch.bfloat16).contiguous()], [wd], [bd], [out], packed)
    torch.cuda.synchronize()
    ref = O.conv3d_transpose_same(x.double(), w.double(), b.double(), s).float()
    got = out.float().cpu()
    assert torch.isfinite(got).all(), "unwritten output voxels"
    assert (got - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("dhw,cins,cout,k,s", [((4, 16, 16), [64], 64, (1, 3, 3), (1, 2, 2)),
                                               ((8, 16, 16), [64, 64], 32, (3, 3, 3), (2, 2, 2)),
                                               ((5, 9, 11), [32], 32, (3, 3, 3), (2, 2, 2))])
def test_conv_strided_wgrad_dgrad_tcgen05(ctx, dhw, cins, cout, k, s):
    """Backward of a STRIDED convolution on the tensor cores: wgrad (element-strided A boxes) and the
    phase-decomposed data gradient, against autograd of the oracle."""
    from m1b200 import ops, _lib
    g = torch.Generator().manual_seed(10)
    cin = sum(cins)
    xs = [torch.randn((2, *dhw, c), generator=g).bfloat16().double().requires_grad_() for c in cins]
    w = (torch.randn((*k, cin, cout), generator=g) / (cin * 9) ** 0.5).bfloat16().double().requires_grad_()
    y = O.conv3d_same(torch.cat(xs, -1), w, None, s)
    dy = torch.randn(y.shape, generator=g).bfloat16().double()
    y.backward(dy)
    dev = 'cuda'
    out_dhw = list(y.shape[1:4])
    pad = [ops.same_pads(dhw[i], k[i], s[i])[1] for i in range(3)]
    dyd = dy.to(dev, torch.bfloat16).contiguous()
    d = ops.conv_desc(_lib.CONV_FWD, 2, dhw, out_dhw, k, s, pad, cins, [cout], [(cin * cout, cout, 1)],
                      act_dtype=_lib.BF16, engine=_lib.ENGINE_TCGEN05)
    dw = torch.zeros(w.shape, device=dev)
    ops.conv3d_wgrad(ctx, d, [x.detach().to(dev, torch.bfloat16).contiguous() for x in xs], [dyd], [dw], None)
    torch.cuda.synchronize()
    scale = max(1.0, w.grad.abs().max().item())
    assert (dw.double().cpu() - w.grad).abs().max().item() < 2e-3 * scale
    wd = w.detach().float().to(dev).contiguous()
    offs = [sum(cins[:i]) for i in range(len(cins))]
    wv = [wd.view(-1)[o * cout:] for o in offs]
    dd = ops.conv_desc(_lib.CONV_TRANSPOSED, 2, out_dhw, dhw, k, s, pad, [cout], cins,
                       [(cin * cout, 1, cout)] * len(cins), act_dtype=_lib.BF16, engine=_lib.ENGINE_TCGEN05)
    assert ops.conv3d_tc_supported(dd)
    packed = ops.conv3d_pack_weights(ctx, dd, wv)
    bufs = [torch.full(x.shape, float('nan'), device=dev, dtype=torch.bfloat16) for x in xs]
    ops.conv3d(ctx, dd, [dyd], wv, None, bufs, packed)
    torch.cuda.synchronize()
    for x, bfr in zip(xs, bufs):
        got = bfr.double().cpu()
        assert torch.isfinite(got).all()
        assert (got - x.grad).abs().max().item() < 3e-2 * max(1.0, x.grad.abs().max().item())


def test_conv_wgrad_tcgen05_fused_outputs(ctx):
    """conv1||conv4: the weight gradients of both layers from ONE tensor-core launch (N = 16 + 64)."""
    from m1b200 import ops, _lib
    g = torch.Generator().manual_seed(12)
    dhw, cins, couts, k = (4, 16, 16), [64, 32], [16, 64], (3, 3, 3)
    cin = sum(cins)
    xs = [torch.randn((2, *dhw, c), generator=g).bfloat16().double() for c in cins]
    ws = [torch.zeros((*k, cin, co), dtype=torch.float64, requires_grad=True) for co in couts]
    x = torch.cat(xs, -1)
    dys = []
    for w in ws:
        y = O.conv3d_same(x, w, None, (1, 1, 1))
        dy = torch.randn(y.shape, generator=g).bfloat16().double()
        y.backward(dy)
        dys.append(dy)
    dev = 'cuda'
    pad = [ops.same_pads(dhw[i], k[i], 1)[1] for i in range(3)]
    d = ops.conv_desc(_lib.CONV_FWD, 2, dhw, dhw, k, (1, 1, 1), pad, cins, couts,
                      [(cin * co, co, 1) for co in couts], act_dtype=_lib.BF16, engine=_lib.ENGINE_TCGEN05)
    dws = [torch.zeros(w.shape, device=dev) for w in ws]
    dbs = [torch.zeros(co, device=dev) for co in couts]
    ops.conv3d_wgrad(ctx, d, [t.to(dev, torch.bfloat16).contiguous() for t in xs],
                     [t.to(dev, torch.bfloat16).contiguous() for t in dys], dws, dbs)
    torch.cuda.synchronize()
    for w, dw, dy, db in zip(ws, dws, dys, dbs):
        scale = max(1.0, w.grad.abs().max().item())
        assert (dw.double().cpu() - w.grad).abs().max().item() < 2e-3 * scale
        assert (db.double().cpu() - dy.sum(dim=(0, 1, 2, 3))).abs().max().item() < 1e-2 * scale


@pytest.mark.parametrize("engine_name", ["tcgen05", "simt"])
def test_conv_dgrad_k_fused(ctx, engine_name):
    """Data gradient of the fused conv1||conv4 pair in ONE launch: K runs over [dy1 | dy4], one weight tensor
    per (produced, gathered) pair (m1_conv_desc.w_by_src)."""
    from m1b200 import ops, _lib
    g = torch.Generator().manual_seed(13)
    dhw, cins, couts, k = (4, 16, 16), [64, 32], [16, 64], (3, 3, 3)
    cin = sum(cins)
    xs = [torch.randn((2, *dhw, c), generator=g, dtype=torch.float64, requires_grad=True) for c in cins]
    ws = [(torch.randn((*k, cin, co), generator=g) / (cin * 27) ** 0.5).bfloat16().double() for co in couts]
    x = torch.cat(xs, -1)
    dys = []
    for w in ws:
        y = O.conv3d_same(x, w, None, (1, 1, 1))
        dy = torch.randn(y.shape, generator=g).bfloat16().double()
        y.backward(dy, retain_graph=True)
        dys.append(dy)
    dev = 'cuda'
    pad = [ops.same_pads(dhw[i], k[i], 1)[1] for i in range(3)]
    wd = [w.float().to(dev).contiguous() for w in ws]
    offs = [sum(cins[:i]) for i in range(len(cins))]
    wv = [wd[j].view(-1)[o * couts[j]:] for o in offs for j in range(len(couts))]
    eng = _lib.ENGINE_TCGEN05 if engine_name == "tcgen05" else _lib.ENGINE_SIMT
    d = ops.conv_desc(_lib.CONV_TRANSPOSED, 2, dhw, dhw, k, (1, 1, 1), pad, couts, cins,
                      [(cin * co, 1, co) for co in couts], act_dtype=_lib.BF16, engine=eng, w_by_src=True)
    packed = ops.conv3d_pack_weights(ctx, d, wv) if engine_name == "tcgen05" else None
    bufs = [torch.full(x_.shape, float('nan'), device=dev, dtype=torch.bfloat16) for x_ in xs]
    ops.conv3d(ctx, d, [t.to(dev, torch.bfloat16).contiguous() for t in dys], wv, None, bufs, packed)
    torch.cuda.synchronize()
    for x_, b in zip(xs, bufs):
        err = (b.double().cpu() - x_.grad).abs().max().item()
        assert err < 3e-2 * max(1.0, x_.grad.abs().max().item()), err


@pytest.mark.parametrize("C,N,dtype", [(128, 2, torch.bfloat16), (256, 4, torch.bfloat16), (512, 6, torch.bfloat16),
                                       (32, 2, torch.bfloat16), (64, 8, torch.float32), (128, 3, torch.float32)])
def test_pointwise_heads(ctx, C, N, dtype):
    """mu/log-sigma style 1x1x1 heads (R:networks.py:637-641): streaming forward (fp32 out), data gradient
    (fp32 dy -> activation-dtype dx, overwrite and accumulate) and weight gradient vs autograd of the oracle."""
    from m1b200 import ops, _lib
    g = torch.Generator().manual_seed(21)
    dhw = (3, 7, 9)                                              # ragged voxel count
    x = torch.randn((2, *dhw, C), generator=g).to(dtype).double().requires_grad_()
    w = (torch.randn((1, 1, 1, C, N), generator=g, dtype=torch.float64) / C ** 0.5).requires_grad_()
    b = torch.randn((N,), generator=g, dtype=torch.float64) * 0.1
    y = O.conv3d_same(x, w, b, (1, 1, 1))
    dy = torch.randn(y.shape, generator=g).double()
    y.backward(dy)
    dev = 'cuda'
    code = _lib.BF16 if dtype == torch.bfloat16 else _lib.F32
    xd = x.detach().to(dev, dtype).contiguous()
    wd = w.detach().float().to(dev).contiguous()
    d = ops.conv_desc(_lib.CONV_FWD, 2, dhw, dhw, (1, 1, 1), (1, 1, 1), (0, 0, 0), [C], [N], [(C * N, N, 1)],
                      act_dtype=code, out_dtype=_lib.F32, engine=_lib.ENGINE_SIMT)
    out = torch.full((2, *dhw, N), float('nan'), device=dev)
    before = ctx.launch_count()
    ops.conv3d(ctx, d, [xd], [wd], [b.float().to(dev)], [out])
    assert ctx.launch_count() == before + 1
    torch.cuda.synchronize()
    assert (out.double().cpu() - y.detach()).abs().max().item() < 1e-4
    # wgrad (+ bias gradient)
    dw = torch.full(w.shape, 0.25, device=dev)
    db = torch.zeros(N, device=dev)
    dyd = dy.float().to(dev).contiguous()
    ops.conv3d_wgrad(ctx, d, [xd], [dyd], [dw], [db])
    torch.cuda.synchronize()
    scale = max(1.0, w.grad.abs().max().item())
    assert (dw.double().cpu() - (w.grad + 0.25)).abs().max().item() < 1e-4 * scale
    assert (db.double().cpu() - dy.sum(dim=(0, 1, 2, 3))).abs().max().item() < 1e-3
    # dgrad: gathered = dy (N fp32 channels), produced = dx (C channels)
    dd = ops.conv_desc(_lib.CONV_TRANSPOSED, 2, dhw, dhw, (1, 1, 1), (1, 1, 1), (0, 0, 0), [N], [C],
                       [(C * N, 1, N)], accumulate=[True], act_dtype=_lib.F32, out_dtype=code,
                       engine=_lib.ENGINE_SIMT)
    prior = torch.randn(x.shape, generator=g).to(dtype)
    dx = prior.clone().to(dev)
    ops.conv3d(ctx, dd, [dyd], [wd], None, [dx])
    torch.cuda.synchronize()
    ref = x.grad + prior.double()
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-4
    assert (dx.double().cpu() - ref).abs().max().item() < tol * max(1.0, ref.abs().max().item())


HALO_CASES = [
    ((2, 12, 40), [64], [32], (1, 3, 3)),                   # SW128, one sub-tile column, H ragged vs G*bh
    ((3, 16, 44), [32, 32, 32], [16, 32], (1, 3, 3)),       # SW64 chunks over a virtual concat, W ragged (44 = 40 + 4)
    ((4, 9, 40), [16], [16], (3, 3, 3)),                    # thin 3x3x3 layer (SE conv2 at res0/res1), plane skipping at d = 0, D-1
    ((3, 10, 20), [64, 64], [16, 64], (3, 3, 3)),           # narrow grid: bw = 20, bh = 5
    ((2, 7, 80), [32], [32], (1, 3, 3)),                    # two W tiles, H < G*bh
    ((1, 5, 160), [16, 32], [32, 32, 32], (1, 3, 3)),       # full-resolution row length
    ((2, 6, 40), [128], [64, 256], (3, 3, 3)),              # N = 320: two N tiles
    ((2, 8, 40), [32], [8, 32], (1, 1, 3)),                 # taps along W only; 8-channel produced tensor
    ((2, 8, 40), [32], [48], (3, 3, 1)),                    # taps along D and H only
]


@pytest.mark.parametrize("dhw,cins,couts,k", HALO_CASES)
def test_conv_fwd_halo(ctx, dhw, cins, couts, k):
    """Halo variant of the tcgen05 engine (row-shifted UMMA descriptors over one TMA halo tile) vs the oracle."""
    from m1b200 import _lib
    xs, ws, bs = _mk(2, dhw, cins, couts, k, seed=31)
    got = _run_fwd(ctx, xs, ws, bs, k, (1, 1, 1), torch.bfloat16, _lib.ENGINE_TCGEN05, variant=2)
    ref = _ref_fwd(xs, ws, bs, (1, 1, 1), torch.bfloat16)
    for g, r in zip(got, ref):
        assert torch.isfinite(g).all(), "non-finite output (unwritten voxels?)"
        err = (g - r).abs().max().item()
        assert err < 2e-2, err


@pytest.mark.parametrize("dhw,couts,cins,k", [((3, 12, 40), [16, 32], [32, 32, 32], (1, 3, 3)),
                                              ((4, 9, 40), [16], [16], (3, 3, 3)),
                                              ((2, 10, 80), [16, 64], [64, 64], (1, 3, 3))])
def test_conv_dgrad_halo(ctx, dhw, couts, cins, k):
    """K-fused data gradient of a stride-1 conv (TRANSPOSED gather, w_by_src) on the halo engine, first
    produced tensor accumulating onto existing content."""
    from m1b200 import ops, _lib
    g = torch.Generator().manual_seed(33)
    cin = sum(cins)
    xs = [torch.randn((2, *dhw, c), generator=g, dtype=torch.float64, requires_grad=True) for c in cins]
    ws = [(torch.randn((*k, cin, co), generator=g) / (cin * 9) ** 0.5).bfloat16().double() for co in couts]
    x = torch.cat(xs, -1)
    dys = []
    for w in ws:
        y = O.conv3d_same(x, w, None, (1, 1, 1))
        dy = torch.randn(y.shape, generator=g).bfloat16().double()
        y.backward(dy, retain_graph=True)
        dys.append(dy)
    dev = 'cuda'
    pad = [ops.same_pads(dhw[i], k[i], 1)[1] for i in range(3)]
    wd = [w.float().to(dev).contiguous() for w in ws]
    offs = [sum(cins[:i]) for i in range(len(cins))]
    wv = [wd[j].view(-1)[o * couts[j]:] for o in offs for j in range(len(couts))]
    d = ops.conv_desc(_lib.CONV_TRANSPOSED, 2, dhw, dhw, k, (1, 1, 1), pad, couts, cins,
                      [(cin * co, 1, co) for co in couts], accumulate=[True] + [False] * (len(cins) - 1),
                      act_dtype=_lib.BF16, engine=_lib.ENGINE_TCGEN05, w_by_src=True)
    d.tune[0] = 2
    assert _lib.lib().m1_conv3d_halo_engine(ctypes.byref(d)) == 1
    packed = ops.conv3d_pack_weights(ctx, d, wv)
    prior = torch.randn(xs[0].shape, generator=g).bfloat16()
    bufs = [prior.clone().to(dev)] + [torch.full(x_.shape, float('nan'), device=dev, dtype=torch.bfloat16)
                                      for x_ in xs[1:]]
    ops.conv3d(ctx, d, [t.to(dev, torch.bfloat16).contiguous() for t in dys], wv, None, bufs, packed)
    torch.cuda.synchronize()
    for i, (x_, b) in enumerate(zip(xs, bufs)):
        ref = x_.grad + (prior.double() if i == 0 else 0)
        err = (b.double().cpu() - ref).abs().max().item()
        assert err < 3e-2 * max(1.0, ref.abs().max().item()), (i, err)


@pytest.mark.parametrize("dhw,cins,couts,k", [((4, 16, 16), [128], [64], (3, 3, 3)),
                                              ((6, 20, 20), [128, 128, 64], [128], (3, 3, 3)),
                                              ((4, 16, 32), [32, 32, 32, 32, 32], [32], (1, 3, 3)),
                                              ((5, 10, 10), [256], [64, 64], (3, 3, 3)),
                                              ((3, 13, 40), [64, 32], [16, 64], (3, 3, 3)),     # fused outputs, ragged H
                                              ((2, 9, 44), [128], [32, 128], (1, 3, 3))])
def test_conv_wgrad_tcgen05_shift_mode(ctx, dhw, cins, couts, k):
    """SHIFT mode of the weight-gradient kernel (tune[1] = 2): the kw taps share one activation box through
    row-shifted MN-major descriptors; full-width lines at a common row pitch, zero-filled halo columns."""
    from m1b200 import ops, _lib
    g = torch.Generator().manual_seed(19)
    cin = sum(cins)
    xs = [torch.randn((2, *dhw, c), generator=g).bfloat16().double() for c in cins]
    ws = [torch.zeros((*k, cin, co), dtype=torch.float64, requires_grad=True) for co in couts]
    x = torch.cat(xs, -1)
    dys = []
    for w in ws:
        y = O.conv3d_same(x, w, None, (1, 1, 1))
        dy = torch.randn(y.shape, generator=g).bfloat16().double()
        y.backward(dy)
        dys.append(dy)
    dev = 'cuda'
    pad = [ops.same_pads(dhw[i], k[i], 1)[1] for i in range(3)]
    d = ops.conv_desc(_lib.CONV_FWD, 2, dhw, dhw, k, (1, 1, 1), pad, cins, couts,
                      [(cin * co, co, 1) for co in couts], act_dtype=_lib.BF16, engine=_lib.ENGINE_TCGEN05)
    d.tune[1] = 2
    assert ops.conv3d_wgrad_tc_supported(d), "SHIFT mode refused the launch"
    dws = [torch.full(w.shape, 0.5, device=dev) for w in ws]
    ops.conv3d_wgrad(ctx, d, [t.to(dev, torch.bfloat16).contiguous() for t in xs],
                     [t.to(dev, torch.bfloat16).contiguous() for t in dys], dws, None)
    torch.cuda.synchronize()
    for w, dw in zip(ws, dws):
        ref = w.grad + 0.5
        scale = max(1.0, ref.abs().max().item())
        assert torch.isfinite(dw).all()
        assert (dw.double().cpu() - ref).abs().max().item() < 2e-3 * scale


@pytest.mark.parametrize("dhw,cins,couts,k", TC_CASES)
def test_conv_fwd_tcgen05_multi_tile(ctx, dhw, cins, couts, k):
    from m1b200 import _lib
    xs, ws, bs = _mk(2, dhw, cins, couts, k, seed=1)
    got = _run_fwd(ctx, xs, ws, bs, k, (1, 1, 1), torch.bfloat16, _lib.ENGINE_TCGEN05, variant=3)
    ref = _ref_fwd(xs, ws, bs, (1, 1, 1), torch.bfloat16)
    for g, r in zip(got, ref):
        assert torch.isfinite(g).all(), "non-finite output (unwritten voxels?)"
        assert (g - r).abs().max().item() < 2e-2


@pytest.mark.parametrize("dhw,cins,couts,k,s", STRIDED_TC)
def test_conv_fwd_strided_tcgen05_multi_tile(ctx, dhw, cins, couts, k, s):
    from m1b200 import _lib
    xs, ws, bs = _mk(2, dhw, cins, couts, k, seed=5)
    got = _run_fwd(ctx, xs, ws, bs, k, s, torch.bfloat16, _lib.ENGINE_TCGEN05, variant=3)
    ref = _ref_fwd(xs, ws, bs, s, torch.bfloat16)
    for g, r in zip(got, ref):
        assert g.shape == r.shape and torch.isfinite(g).all()
        assert (g - r).abs().max().item() < 2e-2
