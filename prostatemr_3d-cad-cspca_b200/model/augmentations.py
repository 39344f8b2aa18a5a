"""Mirror of tf2.5/scripts/model/augmentations.py: the train-time augmentations, on the GPU and batched.

The reference maps `augment_tensors(features, targets, augmentation_params, train_obj)` over single volumes inside
its tf.data pipeline (train_model.py:181) - host CPU work that cannot feed a B200 training at ~100 volumes/s per
GPU. Here the batch is augmented where it already lives: one kernel launch per transform over the whole batch
(libm1b200 `m1_augment`, csrc/augment.cu), every sample with its own randomly drawn parameters.

    features, targets = augment_tensors({'image': x}, {'detection': y}, AUGM_PARAMS, train_obj='lesion', rng=rng)

x: (B, D, H, W, C) fp32 CUDA tensor (or anything torch.as_tensor takes - it is moved to the device), y: one-hot
(B, D, H, W, nc). augmentation_params as in the reference (train_model.py AUGM_PARAMS): [prob, tx_prob,
translate_factor, rotation_degree, axial_hflip, zoom_factor, gauss_noise_stddev, chan_shift_factor, sim_poor_scan,
gamma_correct]. The random decisions of one call are a list of per-sample PLANS (`draw_plans`; pass `plans=` to
replay given decisions - how the parity tests drive the oracle and the kernels with identical draws). The order of
the transforms and what is applied to the labels follow augmentations.py:56-119 exactly. No CPU fallback."""
import ctypes
import math

import numpy as np
import torch

from .. import _lib, ops

GEOMETRIC = ('zoom', 'flip', 'rotate', 'translate')                     # applied to the image AND the labels
IMAGE_ONLY = ('chan_shift', 'gamma', 'poor_scan', 'noise')
_OPS = {'zoom': _lib.AUG_ZOOM, 'flip': _lib.AUG_FLIP, 'rotate': _lib.AUG_ROTATE, 'translate': _lib.AUG_TRANSLATE,
        'chan_shift': _lib.AUG_CHANNEL_SHIFT, 'gamma': _lib.AUG_GAMMA, 'poor_scan': _lib.AUG_POOR_SCAN,
        'noise': _lib.AUG_NOISE}


def draw_plan(rng, shape, params):
    """All random decisions of the reference's augment_tensors for ONE volume (D, H, W, C), in its order of draws
    (augmentations.py:52-111), from a numpy Generator. Same structure as oracle.augment_oracle.draw_plan."""
    prob, tx_prob, tf_, rot, flip, zoom, gstd, cs, poor, gam = params
    d, h, w, c = shape
    plan = dict(apply=bool(rng.uniform() > (1 - prob)))
    u = lambda: float(rng.uniform())                                                # noqa: E731
    ri = lambda hi: int(rng.integers(0, max(1, hi)))                                # noqa: E731  maxval exclusive
    if zoom != 0.0:
        lo_s, hi_s = h, int(math.ceil(h * zoom))
        plan['zoom'] = dict(on=u() > tx_prob, scale=int(rng.integers(lo_s, max(lo_s + 1, hi_s))))
    if flip:
        plan['flip'] = dict(on=u() > 0.5)
    if rot != 0:
        plan['rotate'] = dict(on=u() > tx_prob, angle=float(rng.uniform(-rot, rot)))
    if tf_ != 0.0:
        mh, mw = int(math.ceil(h * tf_)), int(math.ceil(w * tf_))
        plan['translate'] = dict(on=u() > tx_prob, top=ri(mh), bottom=ri(mh), right=ri(mw), left=ri(mw))
    if cs != 0:
        mh, mw = int(math.ceil(h * cs)), int(math.ceil(w * cs))
        plan['chan_shift'] = dict(on=u() > tx_prob, top=ri(mh), bottom=ri(mh), right=ri(mw), left=ri(mw),
                                  channel=int(rng.integers(0, 3)))
    if np.sum(gam) != 0:
        plan['gamma'] = dict(on=u() > tx_prob, gamma=float(rng.uniform(gam[0], gam[1])),
                             channels=[u() > 0.5 for _ in range(3)])
    if poor:
        plan['poor_scan'] = dict(on=u() > tx_prob, channels=[u() > 0.5 for _ in range(3)])
    if gstd != 0:
        plan['noise'] = dict(on=u() > tx_prob, stddev=float(rng.uniform(0, gstd)))
    return plan


def draw_plans(rng, batch_shape, params):
    b = batch_shape[0]
    return [draw_plan(rng, tuple(batch_shape[1:]), params) for _ in range(b)]


def rotation_geometry(h, w, angle_deg):
    """rotate_4D_tensor (augmentations.py:217-235): pad width, tfa.image.rotate offsets of the PADDED image (fp32
    like TensorFlow) and the central-crop start; the crop must give back H x W."""
    diagonal = (h ** 2 + w ** 2) ** 0.5
    pad = int(np.ceil((diagonal - min(h, w)) / 2).astype(np.int32))
    hp, wp = h + 2 * pad, w + 2 * pad
    a = angle_deg * math.pi / 180
    ca, sa = np.float32(math.cos(a)), np.float32(math.sin(a))
    x_off = np.float32(((wp - 1) - (ca * (wp - 1) - sa * (hp - 1))) / 2.0)
    y_off = np.float32(((hp - 1) - (sa * (wp - 1) + ca * (hp - 1))) / 2.0)
    frac = h / hp
    hs, ws = int((float(hp) - float(hp) * frac) / 2), int((float(wp) - float(wp) * frac) / 2)
    if hp - 2 * hs != h or wp - 2 * ws != w:
        raise ValueError("rotate_4D_tensor: tf.image.central_crop(%g) of the %dx%d padded slice does not return %dx%d "
                         "(the reference fails on this shape too)" % (frac, hp, wp, h, w))
    return pad, float(ca), float(sa), float(x_off), float(y_off), hs, ws


def pack_plans(plans, h, w):
    """list of plan dicts -> ctypes array of m1_aug_plan (a sample with apply=False has every transform off)"""
    arr = (_lib.AugPlan * len(plans))()
    for q, pl in zip(arr, plans):
        on = lambda k: bool(pl['apply'] and k in pl and pl[k]['on'])                # noqa: E731
        if on('zoom'):
            q.zoom_on, q.zoom_scale = 1, int(pl['zoom']['scale'])
        if on('flip'):
            q.flip_on = 1
        if on('rotate'):
            pad, ca, sa, xo, yo, hs, ws = rotation_geometry(h, w, pl['rotate']['angle'])
            q.rot_on, q.rot_pad, q.rot_crop_h, q.rot_crop_w = 1, pad, hs, ws
            q.rot_cos, q.rot_sin, q.rot_xoff, q.rot_yoff = ca, sa, xo, yo
        if on('translate'):
            t = pl['translate']
            q.tr_on, q.tr_top, q.tr_bottom, q.tr_right, q.tr_left = 1, t['top'], t['bottom'], t['right'], t['left']
        if on('chan_shift'):
            t = pl['chan_shift']
            q.cs_on, q.cs_channel = 1, t['channel']
            q.cs_top, q.cs_bottom, q.cs_right, q.cs_left = t['top'], t['bottom'], t['right'], t['left']
        if on('gamma'):
            q.gamma = pl['gamma']['gamma']
            for ch in range(3):
                q.gamma_on[ch] = int(bool(pl['gamma']['channels'][ch]))
        if on('poor_scan'):
            for ch in range(3):
                q.poor_on[ch] = int(bool(pl['poor_scan']['channels'][ch]))
        if on('noise'):
            q.noise_on, q.noise_std = 1, pl['noise']['stddev']
    return arr


def _active(plans, key):
    return any(pl['apply'] and key in pl and pl[key]['on'] for pl in plans)


def augment_tensors(features, targets, augmentation_params, train_obj='lesion', debug_on=False, *, rng=None,
                    plans=None, noise_eps=None, device=None):
    """Batched GPU counterpart of augmentations.augment_tensors (same arguments up to debug_on). Returns NEW
    `features` / `targets` dicts whose 'image' / 'detection' entries are augmented CUDA tensors.
    rng: numpy Generator for the decisions (default: a fresh one); plans: explicit decisions (parity runs);
    noise_eps: explicit N(0,1) tensor (B, D, H, W, 3) for the additive noise (default: torch.randn on the device)."""
    if train_obj != 'lesion':
        raise NotImplementedError("m1b200 augmentations implement train_obj='lesion' (3 bpMRI channels [+ labels])")
    if not torch.cuda.is_available():
        raise RuntimeError("augment_tensors: no CUDA device - m1b200 has no CPU fallback")
    dev = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
    x = torch.as_tensor(features['image']).to(device=dev, dtype=torch.float32).contiguous()
    y = torch.as_tensor(targets['detection']).to(device=dev, dtype=torch.float32).contiguous()
    assert x.dim() == 5 and y.dim() == 5 and x.shape[:4] == y.shape[:4], "image / detection must be (B, D, H, W, C)"
    b, d, h, w, c = x.shape
    if plans is None:
        plans = draw_plans(rng or np.random.default_rng(), tuple(x.shape), augmentation_params)
    assert len(plans) == b
    ctx = _lib.Context.get(dev.index or 0)
    packed = pack_plans(plans, h, w)
    host = torch.frombuffer(bytearray(bytes(packed)), dtype=torch.uint8)
    plans_dev = host.to(dev)

    def run(t, keys):
        for k in keys:
            if not _active(plans, k):
                continue
            out = torch.empty_like(t)
            eps = None
            if k == 'noise':
                eps = noise_eps if noise_eps is not None else torch.randn((b, d, h, w, 3), device=dev)
                eps = torch.as_tensor(eps).to(device=dev, dtype=torch.float32).contiguous()
            ops.augment(ctx, _OPS[k], t, out, plans_dev, eps)
            t = out
        return t
    x = run(x, GEOMETRIC + IMAGE_ONLY)
    y = run(y, GEOMETRIC)
    f2, t2 = dict(features), dict(targets)
    f2['image'], t2['detection'] = x, y
    return f2, t2


assert ctypes.sizeof(_lib.AugPlan) % 4 == 0
