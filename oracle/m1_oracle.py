"""CPU ORACLE for the M1 hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain PyTorch-CPU (fp64 or fp32) restatement of the reference's arithmetic for the M1
Hierarchical Probabilistic 3D U-Net forward pass, its losses and one optimizer step.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` leg
may import this module; the product package (prostatemr_3d-cad-cspca_b200/) never does.

PARITY UNPINNED: the reference cannot be executed in this image (TensorFlow 2.5, tensorflow_addons
0.14, tensorflow_probability 0.13 and sonnet are absent, Python 3.12, no network) and ships no
tests, golden vectors or saved weights.  The semantics of those third-party layers are restated
from their documented behaviour; the oracle is pinned only by the
known-answer tests of tests/test_oracle.py (loop restatements of the defining sums of Conv3D /
Conv3DTranspose / Focal, adjoint identities, closed forms, torch.distributions).

Reference map (R: = /root/reference/tf2.5/scripts/model/unets/, L: = .../model/losses.py)
  same_pads / conv3d_same             tf.keras.layers.Conv3D(padding='same')      R:networks.py:259,472
  conv3d_transpose_same               tf.keras.layers.Conv3DTranspose('same')     R:networks.py:496-553
  instance_norm                       tfa.layers.InstanceNormalization()          R:network_blocks.py:38-44
  dropout                             tf.nn.dropout / keras Dropout               R:network_blocks.py:137-143
  se_block                            SEResNetBottleNeck.call                     R:network_blocks.py:48-80
  attention_gate                      GridAttentionBlock3D.call                   R:network_blocks.py:106-130
  m1core                              M1Core.__call__                             R:networks.py:568-759
  m1_probabilistic / m1_deterministic m1()                                        R:networks.py:232-392
  focal_loss / elbo_loss              Focal.FL / Focal.loss / EvidenceLowerBound  L:32-49, L:62-63
  l2_penalty                          kernel/bias regularizers                    R:networks.py:259-263
  adam_amsgrad_step                   tf.keras.optimizers.Adam(amsgrad=True)      train_model.py:113-120
  decision_fusion                     M1.decision_fusion                          R:networks.py:209-223
"""
import math
import zlib
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

LRELU = 0.1
IN_EPS = 1e-3


# --------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------
def default_config(**kw):
    """Constructor defaults of M1 / M1Core (R:networks.py:34-55, :418-434)."""
    cfg = dict(
        num_classes=2, dropout_rate=0.5, dropout_mode='standard',
        filters=(32, 64, 128, 256, 512),
        strides=((1, 1, 1), (1, 2, 2), (1, 2, 2), (2, 2, 2), (1, 2, 2)),
        kernel_sizes=((1, 3, 3), (1, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3)),
        se_reduction=(8, 8, 8, 8, 8),
        att_sub_samp=((1, 1, 1), (1, 1, 1), (1, 1, 1), (1, 1, 1)),
        l2_kernel=1e-4, l2_bias=1e-4,
        dense_skip=False, deep_supervision=False, probabilistic=False,
        prob_latent_dims=(1, 1, 1, 1))
    cfg.update(kw)
    return cfg


# --------------------------------------------------------------------------------------------
# parameters: created on first use, like Keras layers (so pruned layers own no weights)
# --------------------------------------------------------------------------------------------
def _seed_of(name, base):
    return (zlib.crc32(name.encode()) + 7919 * base) % (2 ** 31 - 1)


def init_orthogonal(shape, gain, seed):
    """tf.keras.initializers.Orthogonal: QR of a N(0,1) matrix (rows=prod(shape[:-1]), cols=shape[-1])."""
    rows, cols = int(np.prod(shape[:-1])), int(shape[-1])
    rng = np.random.RandomState(seed)
    a = rng.standard_normal((max(rows, cols), min(rows, cols)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    if rows < cols:
        q = q.T
    return (gain * q).reshape(shape)


def init_truncated_normal(shape, std, seed):
    """tf.keras.initializers.TruncatedNormal: redraw beyond two standard deviations."""
    rng = np.random.RandomState(seed)
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2
    return x * std


def init_glorot_uniform(shape, seed):
    """Keras default kernel initializer (SE conv6/conv7, R:network_blocks.py:45-46)."""
    rec = int(np.prod(shape[:-2]))
    fan_in, fan_out = rec * shape[-2], rec * shape[-1]
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return np.random.RandomState(seed).uniform(-lim, lim, shape)


class ParamStore:
    """name -> tensor; kinds: kernel|bias (L2-regularised conv_params layers), se_kernel|se_bias
    (Keras defaults, no L2), gamma|beta (InstanceNorm, no L2)."""

    def __init__(self, dtype=torch.float64, seed=0, requires_grad=False):
        self.p = OrderedDict()
        self.kind = OrderedDict()
        self.dtype = dtype
        self.seed = seed
        self.requires_grad = requires_grad

    def get(self, name, shape, kind):
        if name not in self.p:
            s = _seed_of(name, self.seed)
            if kind == 'kernel':
                v = init_orthogonal(shape, 1.0, s)
            elif kind == 'bias':
                v = init_truncated_normal(shape, 1e-3, s)
            elif kind == 'se_kernel':
                v = init_glorot_uniform(shape, s)
            elif kind in ('se_bias', 'beta'):
                v = np.zeros(shape)
            elif kind == 'gamma':
                v = np.ones(shape)
            else:
                raise ValueError(kind)
            t = torch.tensor(np.asarray(v), dtype=self.dtype)
            t.requires_grad_(self.requires_grad)
            self.p[name] = t
            self.kind[name] = kind
        t = self.p[name]
        assert tuple(t.shape) == tuple(shape), (name, tuple(t.shape), tuple(shape))
        return t

    def num_params(self, kinds=None):
        return sum(int(t.numel()) for n, t in self.p.items() if kinds is None or self.kind[n] in kinds)


# --------------------------------------------------------------------------------------------
# primitive layers (NDHWC tensors, Keras weight layouts)
# --------------------------------------------------------------------------------------------
def same_pads(size, k, s):
    """TF SAME: out=ceil(in/s), pad_total=max((out-1)s+k-in,0), before=total//2, after=rest."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return out, total // 2, total - total // 2


def conv3d_same(x, w, b, stride=(1, 1, 1)):
    """x (N,D,H,W,Cin), w (kd,kh,kw,Cin,Cout), b (Cout,) or None."""
    xt = x.permute(0, 4, 1, 2, 3)
    pads = []
    for dim in (2, 1, 0):  # F.pad wants last dim first: W, H, D
        _, pb, pa = same_pads(x.shape[1 + dim], w.shape[dim], stride[dim])
        pads += [pb, pa]
    xt = F.pad(xt, pads)
    y = F.conv3d(xt, w.permute(4, 3, 0, 1, 2), b, stride=stride)
    return y.permute(0, 2, 3, 4, 1)


def conv3d_transpose_same(x, w, b, stride=(1, 1, 1)):
    """Keras Conv3DTranspose(padding='same'): x (N,D,H,W,Cin), w (kd,kh,kw,Cout,Cin); output grid =
    input grid * stride; the EXACT adjoint of conv3d_same with that kernel/stride (crop
    [pad_before : pad_before + in*s] of the full scatter)."""
    xt = x.permute(0, 4, 1, 2, 3)
    full = F.conv_transpose3d(xt, w.permute(4, 3, 0, 1, 2), None, stride=stride)
    sl = [slice(None), slice(None)]
    for dim in range(3):
        n_out = x.shape[1 + dim] * stride[dim]
        _, pb, _ = same_pads(n_out, w.shape[dim], stride[dim])
        have = full.shape[2 + dim]
        if have < pb + n_out:  # kernel smaller than stride: the tail is zeros
            padspec = [0, 0] * (2 - dim) + [0, pb + n_out - have]
            full = F.pad(full, padspec)
        sl.append(slice(pb, pb + n_out))
    y = full[tuple(sl)]
    if b is not None:
        y = y + b.view(1, -1, 1, 1, 1)
    return y.permute(0, 2, 3, 4, 1)


def instance_norm(x, gamma, beta, eps=IN_EPS):
    mean = x.mean(dim=(1, 2, 3), keepdim=True)
    var = x.var(dim=(1, 2, 3), keepdim=True, unbiased=False)
    return (x - mean) * torch.rsqrt(var + eps) * gamma + beta


def lrelu(x, slope=LRELU):
    return torch.where(x > 0, x, x * slope)


def dropout(x, rate, u):
    """tf.nn.dropout: keep iff u >= rate, scale 1/(1-rate); u has x's shape."""
    if rate == 0.0:
        return x
    keep = (u >= rate).to(x.dtype)
    return x * keep * (1.0 / (1.0 - rate))


def upsample_nearest(x, factors):
    for dim, f in enumerate(factors):
        if f != 1:
            x = x.repeat_interleave(int(f), dim=1 + dim)
    return x


# --------------------------------------------------------------------------------------------
# blocks
# --------------------------------------------------------------------------------------------
class Noise:
    """Injected randomness. Keys: (pass_name, site) -> tensor (dropout uniforms or latent eps).
    Missing entries are drawn from `gen` and recorded so that the product can replay them."""

    def __init__(self, seed=0, dtype=torch.float64):
        self.t = {}
        self.gen = torch.Generator().manual_seed(seed)
        self.dtype = dtype

    def uniform(self, key, shape):
        if key not in self.t:
            self.t[key] = torch.rand(tuple(shape), generator=self.gen, dtype=torch.float32).to(self.dtype)
        return self.t[key]

    def normal(self, key, shape):
        if key not in self.t:
            self.t[key] = torch.randn(tuple(shape), generator=self.gen, dtype=torch.float32).to(self.dtype)
        return self.t[key]


# --------------------------------------------------------------------------------------------
# bf16-storage emulation (test infrastructure for the product's precision='bf16' mode)
# --------------------------------------------------------------------------------------------
# The reference computes in fp32. The product's bf16 mode keeps fp32 arithmetic INSIDE every kernel but stores
# activations and activation gradients between kernels as bf16 and feeds the tensor cores bf16 weights. With
# emulate_bf16_storage() the oracle rounds at exactly those storage points (forward values and, through the
# autograd function below, the gradients that flow back through them), so that the remaining product-vs-oracle
# difference is summation order only - the deviation of the bf16 mode from the fp32 reference is then shown
# to be storage precision, not arithmetic.
# 'fwd' / 'bwd': which KINDS of storage points round their values / their gradients ('conv' = raw convolution
# outputs before a norm or another consumer, 'act' = every other stored activation: norm+LeakyReLU outputs, SE
# block outputs, gated attention tensors, latents, the input image); 'w': bf16 tensor-core weights. The product
# rounds everything; the subsets exist for precision studies (tools/bf16_precision_study.py).
_EMU = {'on': False, 'fwd': ('conv', 'act'), 'bwd': ('conv', 'act'), 'w': True, 'vdtype': torch.bfloat16}


class _RoundBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return x.to(_EMU['vdtype']).to(x.dtype) if fwd else x          # gradients always round to bf16

    @staticmethod
    def backward(ctx, g):
        return (g.to(torch.bfloat16).to(g.dtype) if ctx.bwd else g), None, None


class _RoundBF16Fwd(torch.autograd.Function):      # bf16 operand copy of an fp32 master weight
    @staticmethod
    def forward(ctx, x):
        return x.to(_EMU['vdtype']).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def _q(x, kind='act'):
    if not _EMU['on']:
        return x
    fwd, bwd = kind in _EMU['fwd'], kind in _EMU['bwd']
    return _RoundBF16.apply(x, fwd, bwd) if (fwd or bwd) else x


def _wq(w):
    return _RoundBF16Fwd.apply(w) if (_EMU['on'] and _EMU['w']) else w


class emulate_bf16_storage:
    """with emulate_bf16_storage(): ... - see above. fwd / bwd / weights select subsets (precision studies)."""

    def __init__(self, fwd=('conv', 'act'), bwd=('conv', 'act'), weights=True, value_dtype=torch.bfloat16):
        """value_dtype: storage type of forward values and tensor-core weights (torch.float16 for the 'what if
        the forward pass were fp16' study); gradients always round to bf16."""
        self.cfg = dict(on=True, fwd=tuple(fwd), bwd=tuple(bwd), w=bool(weights), vdtype=value_dtype)

    def __enter__(self):
        self.prev = dict(_EMU)
        _EMU.update(self.cfg)

    def __exit__(self, *a):
        _EMU.update(self.prev)


def _fp32_head(name, kinds):
    """layers the product runs with fp32 weights and fp32 outputs: SE excite (conv6/conv7), the attention psi
    conv (attX/conv3), the mu/log-sigma heads and every logits head"""
    leaf = name.rsplit('/', 1)[-1]
    return (kinds[0] == 'se_kernel' or leaf.startswith('mu_logsig') or leaf.endswith('logits')
            or (leaf == 'conv3' and '/att' in name))


def _conv(ps, name, x, cout, k, s=(1, 1, 1), kinds=('kernel', 'bias')):
    w = ps.get(name + '/kernel', tuple(k) + (x.shape[-1], cout), kinds[0])
    b = ps.get(name + '/bias', (cout,), kinds[1])
    if _fp32_head(name, kinds):
        return conv3d_same(x, w, b, s)
    return _q(conv3d_same(x, _wq(w), b, s), 'conv')


def _convt(ps, name, x, cout, k, s):
    w = ps.get(name + '/kernel', tuple(k) + (cout, x.shape[-1]), 'kernel')
    b = ps.get(name + '/bias', (cout,), 'bias')
    return _q(conv3d_transpose_same(x, _wq(w), b, s), 'act')      # feeds convolutions directly: must be bf16


def _inorm(ps, name, x):
    c = x.shape[-1]
    return instance_norm(x, ps.get(name + '/gamma', (c,), 'gamma'), ps.get(name + '/beta', (c,), 'beta'))


def se_block(ps, name, x, filters, k, s, reduction):
    """SEResNetBottleNeck.call (R:network_blocks.py:48-80). NOTE Q5: the 'residual addition' is a
    multiplication; Q6: the squeeze sees norm3's output."""
    a = _q(lrelu(_inorm(ps, name + '/norm1', _conv(ps, name + '/conv1', x, filters // 4, k, s))))
    b = _q(lrelu(_inorm(ps, name + '/norm2', _conv(ps, name + '/conv2', a, filters // 4, (3, 3, 3)))))
    x_ = _inorm(ps, name + '/norm3', _conv(ps, name + '/conv3', b, filters, (1, 1, 1)))
    residual = x
    if x_.shape[-1] != residual.shape[-1]:
        residual = _inorm(ps, name + '/norm4', _conv(ps, name + '/conv4', residual, filters, k, s))
    pool = x_.mean(dim=(1, 2, 3), keepdim=True)
    g = _conv(ps, name + '/conv6', pool, filters // reduction, (1, 1, 1), kinds=('se_kernel', 'se_bias'))
    g = lrelu(g)
    g = _conv(ps, name + '/conv7', g, filters, (1, 1, 1), kinds=('se_kernel', 'se_bias'))
    g = torch.sigmoid(g)
    return lrelu(x_ * g * residual)


def attention_gate(ps, name, x, g, inter, sub_samp):
    """GridAttentionBlock3D.call (R:network_blocks.py:106-130). Returns (W_y, sigm_psi_f)."""
    theta = _conv(ps, name + '/conv1', x, inter, sub_samp, sub_samp)
    phi = _conv(ps, name + '/conv2', g, inter, (1, 1, 1))
    phi = upsample_nearest(phi, [theta.shape[1 + i] // phi.shape[1 + i] for i in range(3)])
    f = lrelu(theta + phi)
    psi = torch.sigmoid(_conv(ps, name + '/conv3', f, 1, (1, 1, 1)))
    psi = upsample_nearest(psi, [x.shape[1 + i] // psi.shape[1 + i] for i in range(3)])
    y = _q(psi * x)
    wy = _q(_inorm(ps, name + '/norm4', _conv(ps, name + '/conv4', y, inter, (1, 1, 1))))
    return wy, psi


# --------------------------------------------------------------------------------------------
# M1Core.__call__ (R:networks.py:568-759)
# --------------------------------------------------------------------------------------------
def m1core(ps, net, cfg, inputs, prob_mean=False, prob_z_q=None, noise=None, pass_name='pass',
           training=True, stop='full'):
    """stop='latents': the Keras-pruned partial pass (everything that feeds the last latent head)."""
    Fs, S, K = cfg['filters'], cfg['strides'], cfg['kernel_sizes']
    red, sub = cfg['se_reduction'], cfg['att_sub_samp']
    assert len(Fs) == 5 and len(red) == 5
    assert [len(a) for a in sub] == [3, 3, 3, 3]
    assert [len(s) for s in S] == [3] * 5 and [len(k) for k in K] == [3] * 5
    rate = cfg['dropout_rate']
    drop_on = training or cfg['dropout_mode'] == 'monte-carlo'
    dense = cfg['dense_skip']
    L = cfg['prob_latent_dims']
    prob = cfg['probabilistic']
    partial = (stop == 'latents')
    last_lat = max([i for i in range(len(L)) if L[i] != 0], default=-1) if prob else -1
    n = lambda s: net + '/' + s  # noqa: E731

    def drop(site, t, r=rate):
        # (the product stores the SE block's output once, AFTER the dropout that always follows it)
        if not drop_on or r == 0.0:
            return _q(t)
        return _q(dropout(t, r, noise.uniform((pass_name, site), t.shape)))

    out = {}
    inputs = _q(inputs)
    x = _q(lrelu(_inorm(ps, n('norme0'), _conv(ps, n('conve0'), inputs, Fs[0], K[0], S[0]))))
    conv1 = drop('drope1', se_block(ps, n('serse1'), x, Fs[1], K[1], S[1], red[1]))
    conv2 = drop('drope2', se_block(ps, n('serse2'), conv1, Fs[2], K[2], S[2], red[2]))
    conv3 = drop('drope3', se_block(ps, n('serse3'), conv2, Fs[3], K[3], S[3], red[3]))
    convm = drop('drope4', se_block(ps, n('serse4'), conv3, Fs[4], K[4], S[4], red[4]))

    need = lambda lvl: (not partial) or (last_lat >= (3 - lvl) + 1)  # noqa: E731  uconv{lvl}_ needed?
    # which skip concatenations a partial pass needs: uconv3_ feeds sersp3 (before latent idx 1),
    # uconv2_ feeds sersp2 (before latent idx 2), uconv1_ feeds sersp1 (before latent idx 3)
    att3, _ = attention_gate(ps, n('att3'), conv3, convm, Fs[3], sub[3]) if need(3) else (None, None)
    att2, _ = attention_gate(ps, n('att2'), conv2, convm, Fs[2], sub[2]) if need(2) else (None, None)
    att1, _ = attention_gate(ps, n('att1'), conv1, convm, Fs[1], sub[1]) if need(1) else (None, None)
    att0, _ = attention_gate(ps, n('att0'), x, convm, Fs[0], sub[0]) if need(0) else (None, None)

    uconv = [None] * 4
    uconv_ = [None] * 4
    if need(3):
        deconv3 = _convt(ps, n('convtd3'), convm, Fs[3], K[4], S[4])
        if dense and need(2):
            deconv3_up1 = _convt(ps, n('convtd3_up1'), deconv3, Fs[2], K[3], S[3])
            if need(1):
                deconv3_up2 = _convt(ps, n('convtd3_up2'), deconv3_up1, Fs[1], K[2], S[2])
                if need(0):
                    deconv3_up3 = _convt(ps, n('convtd3_up3'), deconv3_up2, Fs[0], K[1], S[1])
        uconv_[3] = torch.cat([deconv3, att3], -1)
    if need(2):
        uconv[3] = drop('dropd3', se_block(ps, n('sersd3'), uconv_[3], Fs[3], K[3], (1, 1, 1), red[3]))
        deconv2 = _convt(ps, n('convtd2'), uconv[3], Fs[2], K[3], S[3])
        if dense:
            if need(1):
                deconv2_up1 = _convt(ps, n('convtd2_up1'), deconv2, Fs[1], K[2], S[2])
                if need(0):
                    deconv2_up2 = _convt(ps, n('convtd2_up2'), deconv2_up1, Fs[0], K[1], S[1])
            uconv_[2] = torch.cat([deconv2, deconv3_up1, att2], -1)
        else:
            uconv_[2] = torch.cat([deconv2, att2], -1)
    if need(1):
        uconv[2] = drop('dropd2', se_block(ps, n('sersd2'), uconv_[2], Fs[2], K[2], (1, 1, 1), red[2]))
        deconv1 = _convt(ps, n('convtd1'), uconv[2], Fs[1], K[2], S[2])
        if dense:
            if need(0):
                deconv1_up1 = _convt(ps, n('convtd1_up1'), deconv1, Fs[0], K[1], S[1])
            uconv_[1] = torch.cat([deconv1, deconv2_up1, deconv3_up2, att1], -1)
        else:
            uconv_[1] = torch.cat([deconv1, att1], -1)
    if need(0):
        uconv[1] = drop('dropd1', se_block(ps, n('sersd1'), uconv_[1], Fs[1], K[1], (1, 1, 1), red[1]))
        deconv0 = _convt(ps, n('convtd0'), uconv[1], Fs[0], K[1], S[1])
        if dense:
            uconv_[0] = torch.cat([deconv0, deconv1_up1, deconv2_up2, deconv3_up3, att0], -1)
        else:
            uconv_[0] = torch.cat([deconv0, att0], -1)
    if not partial:
        uconv[0] = drop('dropd0', se_block(ps, n('sersd0'), uconv_[0], Fs[0], K[0], (1, 1, 1), red[0]), rate / 2)
        y__ = _conv(ps, n('logits'), uconv[0], cfg['num_classes'], (1, 1, 1))
        out['logits'] = y__
    out['summary_shapes'] = [tuple(t.shape[1:]) for t in (x, conv1, conv2, conv3, convm)]
    out['concat_widths'] = [None if u is None else u.shape[-1] for u in uconv_]

    ds_ops = []
    if prob:
        dists, used = [], []
        feat = convm
        # level i uses: latent dim L[i]; dec_hi{3-i}: filters Fs[::-1][i+1], kernel K[::-1][i],
        # stride S[::-1][i]; sersp{3-i}: filters Fs[::-1][i+1], kernel K[::-1][i+1]
        for i in range(4):
            lvl = 3 - i
            if partial and i > last_lat:
                break
            if L[i] != 0:
                ml = _conv(ps, n('mu_logsig%d' % lvl), feat, 2 * L[i], (1, 1, 1))
                mu, logsig = ml[..., :L[i]], ml[..., L[i]:]
                sigma = torch.exp(torch.clamp(logsig, -0.1, 0.1))
                if prob_z_q is not None:
                    z = prob_z_q[len(used)]
                elif prob_mean:
                    z = _q(mu)
                else:
                    z = _q(mu + sigma * noise.normal((pass_name, 'eps%d' % lvl), mu.shape))
                dists.append((mu, sigma))
                used.append(z)
                if partial and i == last_lat:
                    break
                hi_in = torch.cat([z, feat], -1)
            else:
                hi_in = feat
            Fr, Kr, Sr = Fs[::-1], K[::-1], S[::-1]
            up = _convt(ps, n('dec_hi%d' % lvl), hi_in, Fr[i + 1], Kr[i], Sr[i])
            feat = se_block(ps, n('sersp%d' % lvl), torch.cat([up, uconv_[lvl]], -1), Fr[i + 1], Kr[i + 1],
                            (1, 1, 1), red[::-1][i + 1])
            feat = drop('dropp%d' % lvl, feat)
            ds_ops.append(feat)
        out['prob_distributions'] = dists
        out['prob_used_latents'] = used
        if not partial:
            out['prob_decoder_features'] = feat

    if not partial:
        nc = cfg['num_classes']
        heads = [out['logits']]
        if cfg['deep_supervision']:
            srcs = [uconv[1], uconv[2], uconv[3]] if not prob else [ds_ops[-2], ds_ops[-3], ds_ops[-4]]
            # R:networks.py:739-747 — the reference appends to ds_ops after levels 3,2,1 only, so its
            # ds_ops[-1],[-2],[-3] are the res1,res2,res3 features; this loop appends after level 0 too,
            # hence the shifted indices. The prob branch is dead in the reference (Q3) and only reachable
            # here with ds_in_prob='intended'.
            ups = [np.array(S[1]), np.array(S[1]) * np.array(S[2]),
                   np.array(S[1]) * np.array(S[2]) * np.array(S[3])]
            for j, (t, u) in enumerate(zip(srcs, ups)):
                heads.append(_conv(ps, n('dsy%d_logits' % (j + 1)), upsample_nearest(t, u), nc, (1, 1, 1)))
        out['y_softmax'] = torch.cat([torch.softmax(h, -1) for h in heads], -1)
        out['y_sigmoid'] = torch.cat([torch.sigmoid(h) for h in heads], -1)
        out['y_'] = torch.argmax(out['logits'], -1)
    return out


def kl_mvn_diag(q, p):
    """tfp.distributions.kl_divergence(MultivariateNormalDiag q, p): per-voxel sum over the event axis."""
    (mq, sq), (mp, sp) = q, p
    return (torch.log(sp / sq) + (sq ** 2 + (mq - mp) ** 2) / (2 * sp ** 2) - 0.5).sum(-1)


def m1_deterministic(ps, cfg, inputs, noise=None, training=True, net='m1', stage=''):
    """m1() deterministic branch (R:networks.py:266-294) with the intended prob_mean=False,
    prob_z_q=None (Q1). stage: prefix of parameter names and noise keys (second stage of a cascade)."""
    c = dict(cfg, probabilistic=False)
    return m1core(ps, stage + net, c, inputs, False, None, noise, stage + 'det', training)


def m1_probabilistic(ps, cfg, inputs, noise, training=True, ds_in_prob='reference', with_infer=False, stage=''):
    """m1() probabilistic branch (R:networks.py:297-390): Q4 slicing, 4 live passes, KL, softmax.
    stage: prefix of parameter names and noise keys (second stage of a cascade)."""
    nc = cfg['num_classes']
    P = stage
    image = inputs[..., :-(nc - 1)]
    label = inputs[..., -(nc - 1) - 1:-1]          # Q4: this is the LAST IMAGE channel, not the label
    post_in = torch.cat([image, label], -1)
    core = dict(cfg, probabilistic=True,
                deep_supervision=(cfg['deep_supervision'] and ds_in_prob == 'intended'))  # Q3
    q_sample = m1core(ps, P + 'posterior', core, post_in, False, None, noise, P + 'q_sample', training, 'latents')
    q_mean = m1core(ps, P + 'posterior', core, post_in, True, None, noise, P + 'q_mean', training, 'latents')
    p_zq = m1core(ps, P + 'prior', core, image, False, q_sample['prob_used_latents'], noise, P + 'p_z_q', training,
                  'latents')
    p_zqm = m1core(ps, P + 'prior', core, image, False, q_mean['prob_used_latents'], noise, P + 'p_z_qmean',
                   training, 'full')
    wl = ps.get(P + 'final_decoder/logits/kernel', (1, 1, 1, cfg['filters'][0], nc), 'kernel')
    bl = ps.get(P + 'final_decoder/logits/bias', (nc,), 'bias')
    train_conv = conv3d_same(p_zqm['prob_decoder_features'], wl, bl)
    kl = 0.0
    for q, p in zip(q_sample['prob_distributions'], p_zq['prob_distributions']):
        kl = kl + kl_mvn_diag(q, p).sum(dim=(1, 2, 3)).mean()
    out = {'prob_train_conv': train_conv, 'prob_kl': kl}
    sm = torch.softmax(train_conv, -1)
    if cfg['deep_supervision']:
        sm = torch.cat([sm, p_zqm['y_softmax'][..., nc:]], -1)   # empty slice in 'reference' mode (Q3)
    out['prob_softmax'] = sm
    if with_infer:
        p_s = m1core(ps, P + 'prior', core, image, False, None, noise, P + 'p_sample', training, 'full')
        out['prob_infer_conv'] = conv3d_same(p_s['prob_decoder_features'], wl, bl)
    out['passes'] = dict(q_sample=q_sample, q_mean=q_mean, p_z_q=p_zq, p_z_qmean=p_zqm)
    return out


def m1_infer(ps, cfg, inputs, noise, pass_name='p_sample', stage=''):
    """get_detect_model() of a probabilistic model (R:networks.py:196-206): one prior pass with
    z ~ P at every level, dropout active only in 'monte-carlo' mode."""
    nc = cfg['num_classes']
    image = inputs[..., :-(nc - 1)]
    core = dict(cfg, probabilistic=True, deep_supervision=False)
    p_s = m1core(ps, stage + 'prior', core, image, False, None, noise, stage + pass_name, False, 'full')
    wl = ps.get(stage + 'final_decoder/logits/kernel', (1, 1, 1, cfg['filters'][0], nc), 'kernel')
    bl = ps.get(stage + 'final_decoder/logits/bias', (nc,), 'bias')
    return torch.softmax(conv3d_same(p_s['prob_decoder_features'], wl, bl), -1)


STAGE2 = 'stage2/'


def m1_cascade(ps, cfg, image_1, image_2, noise, strategy='identity', training=True, ds_in_prob='reference'):
    """Cascaded two-stage M1 (R:networks.py:109-193): stage 2 sees concat([stage-1 softmax[..., :nc-1], image_2])
    (for nc = 2 that is the BACKGROUND probability, as the reference slices it), the two class-(nc-1) probabilities
    are fused by decision_fusion (R:networks.py:209-223; `strategy` is the value of the `cascaded` argument, True
    read as 'identity', Q8). Outputs as the Keras model: detection_1, detection_2 (+ KL_1, KL_2).
    nc must be 2: Focal.loss over a 2-channel [1-p, p] prediction has no heads otherwise (losses.py:43-49)."""
    nc = cfg['num_classes']
    assert nc == 2, "the cascade's [1-p, p] outputs only make sense for two classes"
    run = m1_probabilistic if cfg['probabilistic'] else m1_deterministic
    kw = dict(ds_in_prob=ds_in_prob) if cfg['probabilistic'] else {}
    key = 'prob_softmax' if cfg['probabilistic'] else 'y_softmax'
    o1 = run(ps, cfg, image_1, noise, training, **kw)
    sm1 = o1[key]
    x2 = torch.cat([sm1[..., :nc - 1], image_2], -1)
    o2 = run(ps, cfg, x2, noise, training, stage=STAGE2, **kw)
    sm2 = o2[key]
    prior_pred, joint_pred = decision_fusion(sm1[..., nc - 1], sm2[..., nc - 1], strategy)
    out = dict(detection_1=prior_pred, detection_2=joint_pred, stage1=o1, stage2=o2)
    if cfg['probabilistic']:
        out['KL_1'], out['KL_2'] = o1['prob_kl'], o2['prob_kl']
    return out


def cascade_train_loss(ps, cfg, image_1, image_2, y_true, noise, strategy='identity', alpha=(0.75, 0.25), gamma=2.0,
                       kl_weight=10.0, det_weights=(1.0, 1.0)):
    """Keras objective of the cascaded model compiled with one Focal per detection output and one ELBO per KL
    output: sum_i det_weights[i] * Focal(detection_i) + kl_weight * (KL_1 + KL_2) + sum(L2)."""
    o = m1_cascade(ps, cfg, image_1, image_2, noise, strategy, True)
    f1 = focal_loss(y_true, o['detection_1'], alpha, gamma)
    f2 = focal_loss(y_true, o['detection_2'], alpha, gamma)
    total = det_weights[0] * f1 + det_weights[1] * f2 + l2_penalty(ps, cfg)
    res = dict(detection_1=o['detection_1'], detection_2=o['detection_2'], detection_1_loss=f1, detection_2_loss=f2)
    if cfg['probabilistic']:
        total = total + kl_weight * (elbo_loss(o['KL_1']) + elbo_loss(o['KL_2']))
        res.update(KL_1=o['KL_1'], KL_2=o['KL_2'])
    res['loss'] = total
    return res


# --------------------------------------------------------------------------------------------
# losses, regularisers, optimizer
# --------------------------------------------------------------------------------------------
def focal_fl(y_true, y_pred, alpha, gamma):
    """Focal.FL (L:32-39)."""
    a = torch.tensor(alpha, dtype=y_pred.dtype)
    eps = 1e-7
    y_pred = y_pred / y_pred.sum(-1, keepdim=True)
    y_pred = torch.clamp(y_pred, eps, 1 - eps)
    ce = y_true * (-torch.log(y_pred))
    gw = y_true * torch.pow(1.0 - y_pred, gamma)
    fl = a * gw * ce
    return fl.sum(dim=(1, 2, 3, 4)).mean(0)


def focal_loss(y_true, y_pred, alpha=(0.25, 0.75), gamma=2.0):
    """Focal.loss (L:43-49): mean over the C_pred // C_true heads."""
    nc = y_true.shape[-1]
    heads = y_pred.shape[-1] // nc
    return torch.stack([focal_fl(y_true, y_pred[..., nc * i:nc * (i + 1)], alpha, gamma)
                        for i in range(heads)]).mean()


def elbo_loss(y_pred, beta=1.0):
    """EvidenceLowerBound.loss (L:62-63)."""
    return beta * y_pred.sum()


def l2_penalty(ps, cfg):
    """Keras l2 regularisers on kernel+bias of every conv_params layer (R:networks.py:259-263)."""
    tot = 0.0
    for name, t in ps.p.items():
        if ps.kind[name] == 'kernel':
            tot = tot + cfg['l2_kernel'] * (t ** 2).sum()
        elif ps.kind[name] == 'bias':
            tot = tot + cfg['l2_bias'] * (t ** 2).sum()
    return tot


def train_loss(ps, cfg, inputs, y_true, noise, alpha=(0.75, 0.25), gamma=2.0, kl_weight=10.0,
               ds_in_prob='reference'):
    """Keras train_step objective: 1*Focal + kl_weight*ELBO + sum(L2) (train_model.py:124-131,231)."""
    if cfg['probabilistic']:
        o = m1_probabilistic(ps, cfg, inputs, noise, True, ds_in_prob)
        det, kl = o['prob_softmax'], o['prob_kl']
        fl = focal_loss(y_true, det, alpha, gamma)
        total = fl + kl_weight * elbo_loss(kl) + l2_penalty(ps, cfg)
        return dict(loss=total, detection_loss=fl, KL_loss=elbo_loss(kl), detection=det, KL=kl)
    o = m1_deterministic(ps, cfg, inputs, noise, True)
    fl = focal_loss(y_true, o['y_softmax'], alpha, gamma)
    return dict(loss=fl + l2_penalty(ps, cfg), detection_loss=fl, detection=o['y_softmax'])


def adam_amsgrad_step(w, g, m, v, vhat, step, lr, b1=0.9, b2=0.999, eps=1e-7):
    """tf.keras.optimizers.Adam(amsgrad=True), TF 2.5 dense update; `step` counts from 1."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    vhat = torch.maximum(vhat, v)
    lr_t = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
    w = w - lr_t * m / (torch.sqrt(vhat) + eps)
    return w, m, v, vhat


def cosine_decay_restarts(step, initial_lr, first_decay_steps, t_mul=2.0, m_mul=1.0, alpha=0.0):
    """tf.keras.optimizers.schedules.CosineDecayRestarts (README.md:53-58)."""
    completed = step / first_decay_steps
    if t_mul == 1.0:
        i_restart = math.floor(completed)
        completed -= i_restart
    else:
        i_restart = math.floor(math.log(1.0 - completed * (1.0 - t_mul)) / math.log(t_mul))
        sum_r = (1.0 - t_mul ** i_restart) / (1.0 - t_mul)
        completed = (completed - sum_r) / t_mul ** i_restart
    m_fac = m_mul ** i_restart
    cosine = 0.5 * m_fac * (1.0 + math.cos(math.pi * completed))
    return initial_lr * ((1 - alpha) * cosine + alpha)


def decision_fusion(prior, follow, strategy='identity'):
    """M1.decision_fusion (R:networks.py:209-223) on class-1 probabilities."""
    if strategy in ('identity', True):
        joint = follow
    elif strategy == 'noisy-or':
        joint = 1 - (1 - prior) * (1 - follow)
    elif strategy == 'bayes':
        joint = (prior * follow + 1e-9) / (prior * follow + 1e-9 + (1 - prior) * (1 - follow))
    else:
        raise ValueError(strategy)
    return torch.stack([1 - prior, prior], -1), torch.stack([1 - joint, joint], -1)


# --------------------------------------------------------------------------------------------
# synthetic inputs shared by tests and bench
# --------------------------------------------------------------------------------------------
def synthetic_batch(batch, dims, image_channels=3, num_classes=2, seed=1234, dtype=torch.float32,
                    probabilistic=True):
    g = torch.Generator().manual_seed(seed)
    D, H, W = dims
    img = torch.randn((batch, D, H, W, image_channels), generator=g, dtype=torch.float32)
    zz, yy, xx = torch.meshgrid(torch.arange(D), torch.arange(H), torch.arange(W), indexing='ij')
    lab = torch.zeros((batch, D, H, W), dtype=torch.float32)
    for b in range(batch):
        c = [int(torch.randint(0, s, (1,), generator=g)) for s in (D, H, W)]
        r = float(torch.randint(2, max(3, min(H, W) // 8 + 3), (1,), generator=g))
        lab[b] = (((zz - c[0]) * 2.0) ** 2 + (yy - c[1]) ** 2 + (xx - c[2]) ** 2 <= r * r).float()
    onehot = torch.stack([1 - lab, lab], -1) if num_classes == 2 else F.one_hot(lab.long(), num_classes).float()
    x = torch.cat([img, lab[..., None]], -1) if probabilistic else img
    return x.to(dtype), onehot.to(dtype)
