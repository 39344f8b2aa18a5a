"""Host-side planning of the tcgen05 engines (m1_conv3d_plan_info): tiling invariants for the convolution
launch shapes of the README-configured M1 at batch 8 - no GPU needed, the planners are pure host code."""
import numpy as np
import pytest

import m1b200  # noqa: F401
from m1b200 import _lib, ops

SMEM_MAX = 227 * 1024
B = 8
# (grid, gathered channels, produced channels, kernel): the stride-1 SE convolutions of the full prior pass
# (R:networks.py:596-621,653-725 concat widths; conv1||conv4 fused along N) and the thin conv2 layers
STRIDE1 = [
    ((20, 160, 160), [32] * 6, [16, 32], (1, 3, 3)),      # sersp0 (f/4 = 8 padded to 16)
    ((20, 160, 160), [32] * 5, [16, 32], (1, 3, 3)),      # sersd0
    ((20, 80, 80), [64] * 5, [16, 64], (1, 3, 3)),        # sersp1
    ((20, 80, 80), [64] * 4, [16, 64], (1, 3, 3)),        # sersd1
    ((20, 40, 40), [128] * 4, [32, 128], (3, 3, 3)),      # sersp2
    ((20, 40, 40), [128] * 3, [32, 128], (3, 3, 3)),      # sersd2
    ((10, 20, 20), [256] * 3, [64, 256], (3, 3, 3)),      # sersp3
    ((10, 20, 20), [256] * 2, [64, 256], (3, 3, 3)),      # sersd3
    ((20, 160, 160), [16], [32], (1, 3, 3)),              # stem conve0 (4 input channels padded to 16)
    ((20, 160, 160), [16], [16], (3, 3, 3)),              # SE conv2 at res0
    ((20, 80, 80), [16], [16], (3, 3, 3)),                # SE conv2 at res1
    ((20, 40, 40), [32], [32], (3, 3, 3)),                # SE conv2 at res2
    ((10, 20, 20), [64], [64], (3, 3, 3)),                # SE conv2 at res3
    ((5, 10, 10), [128], [128], (3, 3, 3)),               # SE conv2 at res4
]


def _desc(dhw, cins, couts, k, mode=_lib.CONV_FWD):
    pad = [ops.same_pads(dhw[i], k[i], 1)[1] for i in range(3)]
    cin = sum(cins)
    return ops.conv_desc(mode, B, dhw, dhw, k, (1, 1, 1), pad, cins, couts, [(cin * co, co, 1) for co in couts],
                         act_dtype=_lib.BF16, engine=_lib.ENGINE_TCGEN05)


@pytest.mark.parametrize("dhw,cins,couts,k", STRIDE1)
def test_per_tap_plan(dhw, cins, couts, k):
    d = _desc(dhw, cins, couts, k)
    assert ops.conv3d_tc_supported(d)
    ck, n_tile, n_tiles, bd, bh, bw, stages, group, smem, tmem, ctas = ops.conv3d_plan_info(d, 0)
    n_pad = -(-sum(couts) // 16) * 16
    assert n_tile * n_tiles == n_pad and n_tile % 16 == 0 and n_tile <= 256
    assert all(c % ck == 0 for c in cins) and ck in (16, 32, 64)
    assert bd * bh * bw <= 128 and stages >= 1 and smem <= SMEM_MAX
    assert n_tile <= tmem <= 512 and tmem & (tmem - 1) == 0
    tiles = -(-dhw[0] // bd) * -(-dhw[1] // bh) * -(-dhw[2] // bw)
    assert ctas == B * tiles * n_tiles
    assert bd * bh * bw * tiles < 1.35 * np.prod(dhw)              # <= 35 % idle MMA rows


@pytest.mark.parametrize("dhw,cins,couts,k", STRIDE1)
def test_halo_plan(dhw, cins, couts, k):
    d = _desc(dhw, cins, couts, k)
    info = ops.conv3d_plan_info(d, 1)
    assert info, "stride-1 3x3 launches are plannable on the halo engine"
    ck, n_tile, n_tiles, G, bh, bw, P, L, stages, smem, tmem, ctas, stage_bytes, a_alloc = info
    kh, kw = k[1], k[2]
    assert P == bw + kw - 1 and L == G * bh + kh - 1
    assert (bh - 1) * P + bw <= 128, "a sub-tile is 128 consecutive rows of the linearised halo tile"
    assert G * n_tile <= tmem <= 512 and 1 <= G <= 4
    assert stages >= 2 and smem <= SMEM_MAX and stages * stage_bytes + 2048 == smem
    # the last shifted MMA of the last sub-tile stays inside the activation allocation
    rows = (G - 1) * bh * P + (kh - 1) * P + (kw - 1) + 128
    assert a_alloc >= max(rows, L * P) * ck * 2 and a_alloc % 1024 == 0
    assert stage_bytes >= a_alloc + kh * kw * n_tile * ck * 2
    # the heuristic prefers the halo engine exactly for the few-produced-channel layers on wide grids
    d.tune[0] = 0
    assert bool(ops.conv3d_halo_engine(d)) == (n_tile <= 96 and bh * bw / 128.0 >= 0.85)
    d.tune[0] = 1
    assert not ops.conv3d_halo_engine(d)
    d.tune[0] = 2
    assert ops.conv3d_halo_engine(d)


def test_halo_engine_refuses_strided_and_pointwise():
    d = ops.conv_desc(_lib.CONV_FWD, B, (20, 80, 80), (20, 40, 40), (3, 3, 3), (1, 2, 2), (1, 0, 0), [64], [32, 128],
                      [(64 * 32, 32, 1), (64 * 128, 128, 1)], act_dtype=_lib.BF16)
    d.tune[0] = 2
    assert ops.conv3d_tc_supported(d) and not ops.conv3d_halo_engine(d) and ops.conv3d_plan_info(d, 1) == []
    p = _desc((20, 40, 40), [32], [128], (1, 1, 1))
    p.tune[0] = 2
    assert not ops.conv3d_halo_engine(p)


@pytest.mark.parametrize("dhw,cins,couts,k", STRIDE1)
def test_wgrad_plans(dhw, cins, couts, k):
    d = _desc(dhw, cins, couts, k)
    assert ops.conv3d_wgrad_tc_supported(d)
    ck, cb, n_tile, tpg, mpg, kv, bd, bh, bw, stages, smem, tmem, shift, taps_in_m, stage_bytes = \
        ops.conv3d_plan_info(d, 2)
    assert kv == bd * bh * bw and kv % 16 == 0 and shift == 0
    assert stages >= 2 and smem <= SMEM_MAX and tpg * mpg * n_tile <= tmem <= 512
    assert taps_in_m == int(len(cins) == 1 and cins[0] == ck and ck < 128)
    # SHIFT mode: full-width lines at a common pitch, K rows a multiple of the MMA K
    d.tune[1] = 2
    info = ops.conv3d_plan_info(d, 2)
    if taps_in_m or k[2] != 3 or 3 * n_tile > 512:
        assert info == [] or info[12] == 0
        return
    ck, cb, n_tile, tpg, mpg, kv, bd, bh, P, stages, smem, tmem, shift, taps_in_m, stage_bytes = info
    assert shift == 1 and tpg == 3 and bd == 1
    assert P >= dhw[2] + 2 and kv == bh * P and kv % 16 == 0
    assert stages >= 2 and smem <= SMEM_MAX and 3 * mpg * n_tile <= tmem <= 512
    assert stage_bytes >= (128 // ck) * (kv + 8) * ck * 2 * mpg + (n_tile // cb) * kv * cb * 2
    assert dhw[2] * bh >= 0.8 * kv                                  # <= 20 % zero-filled K rows


def test_every_launch_of_the_training_step_is_planned():
    """Shape-only trace of the cfg-2 training step (tools/list_launches.py): every bf16 convolution launch is
    taken by the tcgen05 engine (only the fp32-output heads stay on the CUDA cores), the algorithmic FLOPs match
    SURVEY.md section 8(d), and the plan invariants hold for every real launch shape."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
    import list_launches as LL
    full = sum(r['flops'] for r in LL.trace(batch=B, dead=True, share=False)) / B / 1e9
    assert abs(full - 1675.31) < 0.01 * 1675.31, full                # fwd 837.65 GMAC incl. dead branches
    unshared = LL.trace(batch=B, share=False)
    gflop_per_volume = sum(r['flops'] for r in unshared) / B / 1e9
    assert abs(gflop_per_volume - 1613.15) < 0.001 * 1613.15, gflop_per_volume     # live graph: 806.58 GMAC
    log = LL.trace(batch=B)
    assert len(log) > 150 and len(log) == len(unshared) - 8          # stem + serse1 (4 launches) shared by 2 x 2 passes
    gflop_per_volume = sum(r['flops'] for r in log) / B / 1e9
    assert abs(gflop_per_volume - 1595.23) < 0.001 * 1595.23, gflop_per_volume     # executed: 797.61 GMAC
    n_halo = n_shift = 0
    for r in log:
        d = LL.fwd_desc(r)
        if r['out_fp32']:
            assert not ops.conv3d_tc_supported(d), r['names']                       # mu/log-sigma heads
            continue
        assert ops.conv3d_tc_supported(d), r['names']
        ck, n_tile, n_tiles, bd, bh, bw, stages, group, smem, tmem, ctas = ops.conv3d_plan_info(d, 0)
        assert smem <= SMEM_MAX and n_tile <= tmem <= 512 and bd * bh * bw <= 128 and stages >= 1
        halo = ops.conv3d_plan_info(d, 1)
        if halo:
            n_halo += 1
            assert not r['transposed'] and tuple(r['stride']) == (1, 1, 1)
            ck, n_tile, n_tiles, G, hb, wb, P, L, st, smem, tmem, ctas_sm, stage_bytes, a_alloc = halo
            assert (hb - 1) * P + wb <= 128 and G * n_tile <= tmem <= 512 and st * stage_bytes + 2048 == smem <= SMEM_MAX
        if not r['transposed']:
            assert ops.conv3d_wgrad_tc_supported(d), r['names']
            info = ops.conv3d_plan_info(d, 2)
            assert info[5] % 16 == 0 and info[10] <= SMEM_MAX and info[11] <= 512
            d.tune[1] = 2
            sh = ops.conv3d_plan_info(d, 2)
            if sh and sh[12]:
                n_shift += 1
                assert sh[8] >= r['out_dhw'][2] + 2 and sh[5] == sh[7] * sh[8] and sh[5] % 16 == 0
                assert sh[10] <= SMEM_MAX and sh[11] <= 512 and sh[9] >= 2
    assert n_halo >= 40 and n_shift >= 12, (n_halo, n_shift)
