"""Train-time augmentations (tf2.5/scripts/model/augmentations.py:36-378).

CPU part: the oracle (oracle/augment_oracle.py) against known answers - identities, integer shifts, quarter turns,
torch's half-pixel-centre interpolation (an independent implementation of the TF2 resize rule), closed-form gamma
statistics - and the plan logic of the host mirror. GPU part: every kernel of csrc/augment.cu and the whole chain
against the oracle on identical plans."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import augment_oracle as A

PARAMS = [1.0, 0.15, 0.10, 7.5, True, 1.15, 0.03, 0.05, True, [0.75, 1.25]]     # everything on, mostly applied


def _vol(seed=0, shape=(3, 24, 24, 4)):
    r = np.random.default_rng(seed)
    return r.standard_normal(shape).astype(np.float32)


# ---- oracle: third-party semantics --------------------------------------------------------------------------------
@pytest.mark.parametrize("out", [(18, 18), (31, 29), (24, 24), (9, 40)])
def test_resize_bilinear_is_half_pixel_centres(out):
    x = _vol(1)
    ref = torch.nn.functional.interpolate(torch.from_numpy(x).permute(0, 3, 1, 2), size=out, mode='bilinear',
                                          align_corners=False, antialias=False).permute(0, 2, 3, 1).numpy()
    assert np.abs(A.resize_bilinear(x, *out) - ref).max() < 2e-6


@pytest.mark.parametrize("out", [(18, 18), (32, 32), (24, 24)])
def test_resize_nearest_is_half_pixel_centres(out):
    x = _vol(2)
    ref = torch.nn.functional.interpolate(torch.from_numpy(x).permute(0, 3, 1, 2), size=out,
                                          mode='nearest-exact').permute(0, 2, 3, 1).numpy()
    assert np.array_equal(A.resize_nearest(x, *out), ref)


def test_identities_and_integer_shifts():
    x = _vol(3)
    assert np.array_equal(A.zoom_4D_tensor(x, x.shape[1]), x)                       # scale == H: identity
    assert np.array_equal(A.axial_4D_hflip(A.axial_4D_hflip(x)), x)
    assert np.array_equal(A.translate_4D_tensor(x, 0, 0, 0, 0), x)
    assert np.array_equal(A.translate_4D_tensor(x, 3, 3, 2, 2), x)                   # equal pads cancel
    t = A.translate_4D_tensor(x, 4, 0, 0, 0)                                         # shift down by 4, mirrored top
    assert np.array_equal(t[:, 4:], x[:, :-4]) and np.array_equal(t[:, :4], x[:, :4][:, ::-1])
    t = A.translate_4D_tensor(x, 0, 0, 5, 0)                                         # shift left by 5, mirrored right
    assert np.array_equal(t[:, :, :-5], x[:, :, 5:]) and np.array_equal(t[:, :, -5:], x[:, :, -5:][:, :, ::-1])
    c = A.channel_shift_4D_tensor(x, 1, 2, 0, 0, 0)
    assert np.array_equal(c[..., [0, 2, 3]], x[..., [0, 2, 3]]) and not np.array_equal(c[..., 1], x[..., 1])
    assert np.abs(A.rotate_4D_tensor(x, 0.0) - x).max() < 1e-6


def test_rotation_quarter_turn_and_centre():
    x = _vol(4, (2, 16, 16, 1))
    q = A.rotate_bilinear(x, math.pi / 2)               # tfa.image.rotate: counter-clockwise by the angle
    assert np.abs(q - np.rot90(x, k=1, axes=(1, 2))).max() < 1e-5
    # the centre pixel of an odd-sized image is a fixed point of any rotation
    y = _vol(5, (1, 15, 15, 1))
    assert abs(A.rotate_bilinear(y, 0.37)[0, 7, 7, 0] - y[0, 7, 7, 0]) < 1e-5
    assert A.rotation_pad(160, 160) == 34 and A.rotate_4D_tensor(_vol(6, (1, 32, 32, 2)), 5.0).shape == (1, 32, 32, 2)


def test_gamma_keeps_mean_and_std_and_poor_scan_constant():
    x = _vol(7, (4, 20, 20, 1)) * 2.0 + 0.5
    g = A.gamma_shift_3D_tensor(x, 1.2)
    assert abs(g.mean() - x.mean()) < 1e-4 and abs(g.std() - x.std()) < 1e-4       # "retain original distribution shape"
    assert np.abs(A.gamma_shift_3D_tensor(x, 1.0) - x).max() < 1e-4                 # gamma 1: identity up to rounding
    # monotone: the rank order of the intensities is preserved
    assert np.array_equal(np.argsort(g.ravel(), kind='stable'), np.argsort(x.ravel(), kind='stable'))
    c = np.full((2, 16, 16, 1), 3.25, np.float32)
    assert np.array_equal(A.sim_poor_scan_3D_tensor(c), c)


def test_plan_draws_and_chain_shapes():
    from m1b200.model import augmentations as M
    shape = (6, 32, 32, 4)
    rng_a, rng_b = np.random.default_rng(5), np.random.default_rng(5)
    for _ in range(20):
        pa, pb = A.draw_plan(rng_a, shape, PARAMS, with_noise=False), M.draw_plan(rng_b, shape, PARAMS)
        assert pa == pb                                          # host mirror and oracle draw identical decisions
        z = pa['zoom']
        assert 32 <= z['scale'] < math.ceil(32 * 1.15) and -7.5 <= pa['rotate']['angle'] <= 7.5
        assert all(0 <= pa['translate'][k] < math.ceil(32 * 0.10) for k in ('top', 'bottom', 'right', 'left'))
    off = A.draw_plan(np.random.default_rng(0), shape, [0.0] + PARAMS[1:])
    x, y = _vol(8, shape), (_vol(9, (6, 32, 32, 2)) > 0).astype(np.float32)
    xa, ya = A.augment_volume(x, y, off)
    assert not off['apply'] and np.array_equal(xa, x) and np.array_equal(ya, y)
    plan = A.draw_plan(np.random.default_rng(3), shape, [1.0, 0.0] + PARAMS[2:])       # tx_prob 0: everything fires
    xa, ya = A.augment_volume(x, y, plan)
    assert xa.shape == x.shape and ya.shape == y.shape and np.isfinite(xa).all()
    assert np.array_equal(xa[..., 3], A.augment_volume(x, y, {**plan, 'gamma': {**plan['gamma'], 'on': False},
                                                               'poor_scan': {**plan['poor_scan'], 'on': False},
                                                               'noise': {**plan['noise'], 'on': False},
                                                               'chan_shift': {**plan['chan_shift'], 'on': False}})[0][..., 3])
    pad, ca, sa, xo, yo, hs, ws = M.rotation_geometry(160, 160, 5.0)
    assert pad == 34 and hs == ws == 34


def test_aug_plan_struct_layout(tmp_path):
    from m1b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "m1b200.h"\nint main(void) {\n'
                   '  printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(m1_aug_plan), offsetof(m1_aug_plan, rot_cos),\n'
                   '         offsetof(m1_aug_plan, tr_on), offsetof(m1_aug_plan, cs_on), offsetof(m1_aug_plan, gamma_on),\n'
                   '         offsetof(m1_aug_plan, gamma), offsetof(m1_aug_plan, poor_on), offsetof(m1_aug_plan, noise_std));\n'
                   '  return 0;\n}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.dirname(_lib.HEADER_PATH), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    p = _lib.AugPlan
    assert got == [ctypes.sizeof(p), p.rot_cos.offset, p.tr_on.offset, p.cs_on.offset, p.gamma_on.offset,
                   p.gamma.offset, p.poor_on.offset, p.noise_std.offset]


# ---- GPU: kernels against the oracle --------------------------------------------------------------------------------
def _gpu_case(seed, shape=(5, 48, 48, 4), batch=3, tx_prob=0.0):
    r = np.random.default_rng(seed)
    x = r.standard_normal((batch,) + shape).astype(np.float32)
    y = (r.standard_normal((batch,) + shape[:3] + (2,)) > 0.8).astype(np.float32)
    params = [1.0, tx_prob] + PARAMS[2:]
    plans = [A.draw_plan(r, shape, params) for _ in range(batch)]
    return x, y, plans, params


@pytest.mark.gpu
@pytest.mark.parametrize("only", ['zoom', 'flip', 'rotate', 'translate', 'chan_shift', 'gamma', 'poor_scan', 'noise'])
def test_each_transform_matches_the_oracle(ctx, only):
    from m1b200.model import augmentations as M
    x, y, plans, params = _gpu_case(11)
    for pl in plans:                                   # one transform at a time; the last sample keeps everything off
        for k in M.GEOMETRIC + M.IMAGE_ONLY:
            pl[k]['on'] = pl[k]['on'] and k == only
    plans[-1]['apply'] = False
    eps = np.stack([pl['noise']['eps'] for pl in plans])
    f, t = M.augment_tensors({'image': x}, {'detection': y}, params, plans=plans, noise_eps=eps)
    torch.cuda.synchronize()
    for b, pl in enumerate(plans):
        xr, yr = A.augment_volume(x[b], y[b], pl)
        # gamma: powf vs numpy pow, fp32 statistics; rotate: the GPU contracts cos x - sin y + off into FMAs, which
        # moves the fp32 sampling coordinate by an ulp or two (values are N(0,1): 1e-4 absolute)
        tol = 2e-4 if only == 'gamma' else (1e-4 if only == 'rotate' else 2e-5)
        assert np.abs(f['image'][b].cpu().numpy() - xr).max() < tol, (only, b)
        assert np.abs(t['detection'][b].cpu().numpy() - yr).max() < 2e-5, (only, b)
    assert np.array_equal(f['image'][-1].cpu().numpy(), x[-1])


@pytest.mark.gpu
@pytest.mark.parametrize("seed,shape", [(21, (5, 48, 48, 4)), (22, (4, 64, 64, 3)), (23, (20, 160, 160, 4))])
def test_whole_chain_matches_the_oracle(ctx, seed, shape):
    from m1b200.model import augmentations as M
    x, y, plans, params = _gpu_case(seed, shape, batch=2, tx_prob=0.15)
    eps = np.stack([pl['noise']['eps'] for pl in plans])
    f, t = M.augment_tensors({'image': x}, {'detection': y}, params, plans=plans, noise_eps=eps)
    torch.cuda.synchronize()
    for b, pl in enumerate(plans):
        xr, yr = A.augment_volume(x[b], y[b], pl)
        e = np.abs(f['image'][b].cpu().numpy() - xr)
        assert e.max() < 1e-3 and e.mean() < 1e-5, (b, e.max(), e.mean())
        assert np.abs(t['detection'][b].cpu().numpy() - yr).max() < 1e-4


@pytest.mark.gpu
def test_augment_then_train_step(ctx):
    """the augmented CUDA tensors feed M1.train_step directly (no host round trip)"""
    from m1b200.model import augmentations as M
    import test_model_gpu as T
    model, cfg, x, y = T._build(T.TINY, (8, 32, 32), 2, 'fp32', True, True, True)
    f, t = M.augment_tensors({'image': x}, {'detection': y}, PARAMS, rng=np.random.default_rng(1))
    assert f['image'].is_cuda and f['image'].shape == x.shape and t['detection'].shape == y.shape
    out = model.train_step(f, t)
    torch.cuda.synchronize()
    assert math.isfinite(out['focal'].item()) and math.isfinite(out['kl'].item())
