"""Cascaded two-stage M1 (R:networks.py:109-193, decision_fusion :209-223) against the oracle's m1_cascade on identical
weights, inputs and injected noise: both detections, all loss terms, and the gradients of BOTH stages - stage 1 receives
gradient through the stage-2 input (its background probability) and, for noisy-or / bayes, through the fusion."""
import pytest
import torch

from oracle import m1_oracle as O

pytestmark = pytest.mark.gpu

STRIDES = ((1, 1, 1), (1, 2, 2), (1, 2, 2), (2, 2, 2), (2, 2, 2))
KERNELS = ((1, 3, 3), (1, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3))
TINY = dict(filters=(32, 16, 24, 32, 48), se_reduction=(4, 4, 4, 4, 4))     # 32 features at res0: the fused head
MID = dict(filters=(32, 64, 128, 192, 256), se_reduction=(8, 8, 8, 8, 8))


def _build(arch, precision, probabilistic, strategy, sub=((1, 1, 1),) * 4, dims=(8, 32, 32)):
    from m1b200.model import losses, optimizers, unets
    cin = 4 if probabilistic else 3
    model = unets.networks.M1(dims, cin, 2, strides=STRIDES, kernel_sizes=KERNELS, att_sub_samp=sub, dropout_rate=0.5,
                              dropout_mode='monte-carlo', dense_skip=True, deep_supervision=False,
                              probabilistic=probabilistic, prob_latent_dims=(3, 2, 1, 0), cascaded=strategy,
                              summary=False, precision=precision, seed=0, **arch)
    model.compile(optimizer=optimizers.Adam(1e-3, amsgrad=True),
                  loss=[losses.Focal(alpha=[0.75, 0.25], gamma=2.0).loss, losses.EvidenceLowerBound().loss],
                  loss_weights=[1.0, 10.0])
    cfg = O.default_config(num_classes=2, dropout_rate=0.5, dropout_mode='monte-carlo', strides=STRIDES,
                           kernel_sizes=KERNELS, att_sub_samp=sub, dense_skip=True, deep_supervision=False,
                           probabilistic=probabilistic, prob_latent_dims=(3, 2, 1, 0), **arch)
    x, y = O.synthetic_batch(2, dims, probabilistic=probabilistic, seed=21)
    x2 = x.flip(0).contiguous()
    return model, cfg, x, x2, y


def _perturb(ps, seed=17):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, t in ps.p.items():
            if ps.kind[n] in ('gamma', 'beta', 'se_bias'):
                t.add_(0.2 * torch.randn(t.shape, generator=g).to(t.dtype))


def _oracle(cfg, x, x2, y, strategy, dtype):
    ps = O.ParamStore(dtype=dtype, seed=3, requires_grad=True)
    with torch.no_grad():
        O.m1_cascade(ps, cfg, x.to(dtype), x2.to(dtype), O.Noise(0, dtype), strategy)
    _perturb(ps)
    noise = O.Noise(5, dtype)
    r = O.cascade_train_loss(ps, cfg, x.to(dtype), x2.to(dtype), y.to(dtype), noise, strategy, alpha=(0.75, 0.25),
                             gamma=2.0, kl_weight=10.0)
    data = r['detection_1_loss'] + r['detection_2_loss']
    if cfg['probabilistic']:
        data = data + 10.0 * (r['KL_1'] + r['KL_2'])
    data.backward()
    return ps, noise, r


def _check(model, ps, noise, r, x, x2, y, tol_sm, tol_loss, cos_min):
    model.set_weights({n: t.detach().float().numpy() for n, t in ps.p.items()})
    model.set_noise(noise.t)
    out = model.train_step([x, x2], y, apply_update=False)
    torch.cuda.synchronize()
    # detection_1 is one M1 on identical inputs: the north-star bound tol_sm. detection_2 is compared through the
    # CHAIN (its oracle value was computed from the oracle's stage-1 output, ours from our stage-1 output), so its
    # deviation is its own rounding plus the propagated stage-1 deviation: bounded by 2 x tol_sm, mean far below.
    for k, bound in (('detection_1', tol_sm), ('detection_2', 2 * tol_sm)):
        e = (out[k].double().cpu() - r[k].detach().double()).abs()
        print(f'{k}: softmax abs err mean {e.mean().item():.2e} max {e.max().item():.2e} (bound {bound:.0e})')
        assert e.max().item() < bound, (k, e.max().item())
        assert e.mean().item() < 0.1 * tol_sm, (k, e.mean().item())
    for ours, ref in (('focal_1', 'detection_1_loss'), ('focal_2', 'detection_2_loss'), ('kl_1', 'KL_1'), ('kl_2', 'KL_2')):
        if ref in r:
            a, b = out[ours].item(), r[ref].item()
            assert abs(a - b) <= tol_loss * abs(b), (ours, a, b)
    grads = model.gradients()
    res = {}
    for prefix in ('stage2/', ''):
        names = [n for n in ps.p if n.startswith('stage2/') == (prefix == 'stage2/')]
        a = torch.cat([grads[n].double().flatten() for n in names])
        b = torch.cat([(ps.p[n].grad if ps.p[n].grad is not None else torch.zeros_like(ps.p[n])).double().flatten()
                       for n in names])
        res[prefix or 'stage1/'] = (a @ b).item() / (a.norm().item() * b.norm().item())
    print('gradient cosine per stage', res)
    assert min(res.values()) >= cos_min, res


@pytest.mark.parametrize("strategy", ['identity', 'noisy-or', 'bayes'])
@pytest.mark.parametrize("probabilistic", [False, True])
def test_cascade_train_step_fp32(ctx, strategy, probabilistic):
    model, cfg, x, x2, y = _build(TINY, 'fp32', probabilistic, strategy)
    ps, noise, r = _oracle(cfg, x, x2, y, strategy, torch.float64)
    _check(model, ps, noise, r, x, x2, y, tol_sm=1e-4, tol_loss=1e-3, cos_min=0.9999)


def test_cascade_true_is_identity_and_sub_sampled_gates_fp16(ctx):
    """BASELINE cfg-5 in miniature: cascaded=True (read as 'identity', Q8), att_sub_samp=(2,2,2), benchmarked precision"""
    sub = ((2, 2, 2),) * 4
    model, cfg, x, x2, y = _build(MID, 'fp16', True, True, sub=sub, dims=(8, 64, 64))
    assert model.strategy == 0
    ps, noise, r = _oracle(cfg, x, x2, y, 'identity', torch.float32)
    _check(model, ps, noise, r, x, x2, y, tol_sm=2e-2, tol_loss=1e-3, cos_min=0.98)


def test_cascade_inference_fit_and_checkpoint(ctx, tmp_path):
    import numpy as np
    model, cfg, x, x2, y = _build(TINY, 'fp32', True, 'noisy-or')
    ps = O.ParamStore(dtype=torch.float64, seed=3)
    with torch.no_grad():
        O.m1_cascade(ps, cfg, x.double(), x2.double(), O.Noise(0), 'noisy-or')
        O.m1_infer(ps, cfg, x.double(), O.Noise(0))
        O.m1_infer(ps, cfg, torch.cat([x[..., :1], x2], -1).double(), O.Noise(0), stage='stage2/')
    _perturb(ps)
    noise = O.Noise(31)
    with torch.no_grad():
        p1 = O.m1_infer(ps, cfg, x.double(), noise)
        p2 = O.m1_infer(ps, cfg, torch.cat([p1[..., :1], x2.double()], -1), noise, stage='stage2/')
    model.set_weights({n: t.detach().float().numpy() for n, t in ps.p.items()}, strict=False)
    model.set_noise(noise.t)
    q1, q2 = model.get_detect_model()([x, x2])
    torch.cuda.synchronize()
    assert (q1.double().cpu() - p1).abs().max().item() < 1e-4
    assert (q2.double().cpu() - p2).abs().max().item() < 1e-4
    model.set_noise(None, seed=3)
    hist = model.fit(x=[({'image_1': x, 'image_2': x2}, {'detection_1': y, 'detection_2': y})], epochs=3,
                     steps_per_epoch=2, verbose=0)
    assert hist['loss'][-1] < hist['loss'][0]
    path = str(tmp_path / 'cascade.npz')
    model.save(path)
    from m1b200.model import unets
    m2 = unets.networks.M1.load(path)
    assert type(m2).__name__ == 'CascadedM1' and m2.strategy == 1
    w1, w2 = model.get_weights(), m2.get_weights()
    assert all(np.array_equal(w1[k], w2[k]) for k in w1)
    assert m2.optimizer.iterations == model.optimizer.iterations == 6
