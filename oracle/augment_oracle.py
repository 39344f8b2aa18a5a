"""CPU restatement (numpy, fp32 like TensorFlow) of the reference's train-time augmentations - TEST INFRASTRUCTURE ONLY
(tests/, never imported by the product): tf2.5/scripts/model/augmentations.py:36-378, applied per volume (D, H, W, C)
by `train_gen.map(augment_tensors)` before batching (train_model.py:181).

parity unpinned: the arithmetic lives in TensorFlow 2.5 / TensorFlow-Addons 0.14 (`tf.image.resize`, `tf.pad`,
`tfa.image.rotate`, `tf.image.central_crop`), which cannot be imported here; the functions below restate the
documented behaviour of those ops (half-pixel-centre bilinear / nearest resizing without antialiasing, SYMMETRIC
padding, projective transform with bilinear interpolation and constant-0 fill) and are pinned by known-answer tests
only (tests/test_augment.py: identities, integer shifts, 90-degree rotations, closed-form gamma statistics).

The random decisions of `augment_tensors` are drawn ONCE per volume into a plan (`draw_plan`) that both this oracle
and the GPU kernels consume - the same injection idea as the dropout / latent noise of the model."""
import math

import numpy as np

F = np.float32


# ---- third-party semantics restated ---------------------------------------------------------------------------------
def resize_bilinear(img, out_h, out_w):
    """tf.image.resize(img, (out_h, out_w)) on (N, H, W, C): bilinear, half_pixel_centers=True, antialias=False."""
    n, h, w, c = img.shape

    def weights(out, inn):
        scale = F(inn) / F(out)
        src = (np.arange(out, dtype=F) + F(0.5)) * scale - F(0.5)
        fl = np.floor(src)
        lo = np.maximum(fl, 0).astype(np.int64)
        hi = np.minimum(np.ceil(src), inn - 1).astype(np.int64)
        return lo, hi, (src - fl).astype(F)
    y0, y1, ly = weights(out_h, h)
    x0, x1, lx = weights(out_w, w)
    img = img.astype(F)
    top = img[:, y0][:, :, x0] + (img[:, y0][:, :, x1] - img[:, y0][:, :, x0]) * lx[None, None, :, None]
    bot = img[:, y1][:, :, x0] + (img[:, y1][:, :, x1] - img[:, y1][:, :, x0]) * lx[None, None, :, None]
    return (top + (bot - top) * ly[None, :, None, None]).astype(F)


def resize_nearest(img, out_h, out_w):
    """tf.image.resize(..., method='nearest'): half_pixel_centers=True, src = min(floor((i + 0.5) * scale), in - 1)."""
    n, h, w, c = img.shape
    ys = np.minimum(np.floor((np.arange(out_h, dtype=F) + F(0.5)) * (F(h) / F(out_h))), h - 1).astype(np.int64)
    xs = np.minimum(np.floor((np.arange(out_w, dtype=F) + F(0.5)) * (F(w) / F(out_w))), w - 1).astype(np.int64)
    return img[:, ys][:, :, xs]


def pad_symmetric(img, top, bottom, left, right):
    """tf.pad(img, [[0,0],[top,bottom],[left,right],[0,0]], mode='SYMMETRIC') (augmentations.py:331-371)."""
    return np.pad(img, ((0, 0), (top, bottom), (left, right), (0, 0)), mode='symmetric')


def rotate_bilinear(img, angle_rad):
    """tfa.image.rotate(img, angle, interpolation='BILINEAR') (fill_mode='constant', fill_value=0): the output pixel
    (x, y) reads the input at (cos x - sin y + x_off, sin x + cos y + y_off), a rotation about the image centre."""
    n, h, w, c = img.shape
    ca, sa = F(math.cos(angle_rad)), F(math.sin(angle_rad))
    x_off = F(((w - 1) - (ca * (w - 1) - sa * (h - 1))) / 2.0)
    y_off = F(((h - 1) - (sa * (w - 1) + ca * (h - 1))) / 2.0)
    ys, xs = np.meshgrid(np.arange(h, dtype=F), np.arange(w, dtype=F), indexing='ij')
    sx = ca * xs - sa * ys + x_off
    sy = sa * xs + ca * ys + y_off
    x0, y0 = np.floor(sx), np.floor(sy)

    def read(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        v = img[:, np.clip(yy, 0, h - 1).astype(np.int64), np.clip(xx, 0, w - 1).astype(np.int64)].astype(F)
        return v * ok[None, :, :, None].astype(F)
    wx1, wy1 = (sx - x0).astype(F), (sy - y0).astype(F)
    wx0, wy0 = (x0 + 1 - sx).astype(F), (y0 + 1 - sy).astype(F)
    top = wx0[None, :, :, None] * read(y0, x0) + wx1[None, :, :, None] * read(y0, x0 + 1)
    bot = wx0[None, :, :, None] * read(y0 + 1, x0) + wx1[None, :, :, None] * read(y0 + 1, x0 + 1)
    return (wy0[None, :, :, None] * top + wy1[None, :, :, None] * bot).astype(F)


def central_crop(img, frac):
    """tf.image.central_crop on (N, H, W, C): start = int((H - H * frac) / 2), size = H - 2 * start."""
    n, h, w, c = img.shape
    hs, ws = int((float(h) - float(h) * frac) / 2), int((float(w) - float(w) * frac) / 2)
    return img[:, hs:h - hs, ws:w - ws]


# ---- the reference's transforms (same names, augmentations.py:139-327) ------------------------------------------------
def zoom_4D_tensor(x, scale):
    h, w = x.shape[1], x.shape[2]
    s = resize_bilinear(x, scale, scale)
    return s[:, scale - h:scale - h + h, scale - w:scale - w + w]


def axial_4D_hflip(x):
    return x[:, :, ::-1]


def translate_4D_tensor(x, pad_top, pad_bottom, pad_right, pad_left):
    h, w = x.shape[1], x.shape[2]
    p = pad_symmetric(x, pad_top, pad_bottom, pad_left, pad_right)
    return p[:, pad_bottom:pad_bottom + h, pad_right:pad_right + w]


def channel_shift_4D_tensor(x, channel, pad_top, pad_bottom, pad_right, pad_left):
    out = x.copy()
    out[..., channel:channel + 1] = translate_4D_tensor(x[..., channel:channel + 1], pad_top, pad_bottom, pad_right,
                                                        pad_left)
    return out


def rotation_pad(h, w):
    """augmentations.py:221-222"""
    diagonal = (h ** 2 + w ** 2) ** 0.5
    return int(np.ceil((diagonal - min(h, w)) / 2).astype(np.int32))


def rotate_4D_tensor(x, angle_deg):
    h, w = x.shape[1], x.shape[2]
    pad = rotation_pad(h, w)
    p = pad_symmetric(x, pad, pad, pad, pad)
    r = rotate_bilinear(p, angle_deg * math.pi / 180)
    return central_crop(r, h / p.shape[1])


def sim_poor_scan_3D_tensor(x):
    h = x.shape[1]
    lo = resize_bilinear(x, int(h * 0.75), int(h * 0.75))
    return resize_nearest(lo, h, h)


def gamma_shift_3D_tensor(x, gamma):
    x = x.astype(F)
    mn, sd = x.mean(dtype=np.float64), x.std(dtype=np.float64)
    lo, hi = x.min(), x.max()
    x_ = np.power((x - lo) / F(hi - lo + F(1e-8)), F(gamma)) * (hi - lo) + lo
    x_ = x_ - F(x_.mean(dtype=np.float64))
    x_ = x_ / F(x_.std(dtype=np.float64) + 1e-8) * F(sd)
    return (x_ + F(mn)).astype(F)


# ---- plan + application (augment_tensors, augmentations.py:36-132) -----------------------------------------------------
def draw_plan(rng, shape, params, with_noise=True):
    """All random decisions of augment_tensors for ONE volume of shape (D, H, W, C), drawn from a numpy Generator.
    params = [prob, tx_prob, translate_factor, rotation_degree, axial_hflip, zoom_factor, gauss_noise_stddev,
              chan_shift_factor, sim_poor_scan, gamma_correct] (train_model.py AUGM_PARAMS)."""
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
        if with_noise:
            plan['noise']['eps'] = rng.standard_normal((d, h, w, 3)).astype(F)
    return plan


def augment_volume(image, label, plan):
    """image (D, H, W, C), label (D, H, W, nc) -> augmented copies; D plays the batch role of the 4-D image ops."""
    x, y = image.astype(F), label.astype(F)
    if not plan['apply']:
        return x, y
    on = lambda k: k in plan and plan[k]['on']                                      # noqa: E731
    if on('zoom'):
        x = zoom_4D_tensor(x, plan['zoom']['scale'])
    if on('flip'):
        x = axial_4D_hflip(x)
    if on('rotate'):
        x = rotate_4D_tensor(x, plan['rotate']['angle'])
    if on('translate'):
        t = plan['translate']
        x = translate_4D_tensor(x, t['top'], t['bottom'], t['right'], t['left'])
    if on('chan_shift'):
        t = plan['chan_shift']
        x = channel_shift_4D_tensor(x, t['channel'], t['top'], t['bottom'], t['right'], t['left'])
    if on('gamma'):
        x = x.copy()
        for ch in range(3):
            if plan['gamma']['channels'][ch]:
                x[..., ch:ch + 1] = gamma_shift_3D_tensor(x[..., ch:ch + 1], plan['gamma']['gamma'])
    if on('poor_scan'):
        x = x.copy()
        for ch in range(3):
            if plan['poor_scan']['channels'][ch]:
                x[..., ch:ch + 1] = sim_poor_scan_3D_tensor(x[..., ch:ch + 1])
    if on('noise'):
        x = x.copy()
        x[..., :3] += F(plan['noise']['stddev']) * plan['noise']['eps']
    # label augmentations (augmentations.py:113-119)
    if on('zoom'):
        y = zoom_4D_tensor(y, plan['zoom']['scale'])
    if on('flip'):
        y = axial_4D_hflip(y)
    if on('rotate'):
        y = rotate_4D_tensor(y, plan['rotate']['angle'])
    if on('translate'):
        t = plan['translate']
        y = translate_4D_tensor(y, t['top'], t['bottom'], t['right'], t['left'])
    return np.ascontiguousarray(x, dtype=F), np.ascontiguousarray(y, dtype=F)
