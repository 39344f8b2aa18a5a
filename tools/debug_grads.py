"""Per-tensor gradient comparison of the GPU model against the oracle (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
import test_model_gpu as T
from oracle import m1_oracle as O

prec = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
arch = T.TINY if prec == 'fp32' else T.MID
dims = (8, 32, 32) if prec == 'fp32' else (8, 32, 32)
tc = (sys.argv[2] != 'notc') if len(sys.argv) > 2 else True
model, cfg, x, y = T._build(arch, dims, 2, prec, True, True, True, use_tcgen05=tc)
emu = len(sys.argv) > 3 and sys.argv[3] == 'emu'
ps, noise, r = T._oracle_step(cfg, x, y, 'reference', dtype=torch.float64 if prec == 'fp32' else torch.float32,
                              round_bf16=prec != 'fp32', emulate=emu)
model.set_weights({n: t.detach().float().numpy() for n, t in ps.p.items()})
model.set_noise(noise.t)
out = model.train_step(x, y, apply_update=False)
torch.cuda.synchronize()
e = (out['detection'].double().cpu() - r['detection'].detach().double()).abs().flatten()
print('softmax err max', e.max().item(), 'mean', e.mean().item(), 'p99', e.kthvalue(int(0.99 * e.numel())).values.item(),
      'p99.9', e.kthvalue(int(0.999 * e.numel())).values.item(), 'packs', len(model.eng.packs))
print('focal', out['focal'].item(), r['detection_loss'].item(), 'kl', out['kl'].item(), r['KL'].item())
rows = []
grads = model.gradients()
for n, t in ps.p.items():
    g_ref = (t.grad if t.grad is not None else torch.zeros_like(t)).double().flatten()
    g = grads[n].double().cpu().flatten()
    na, nb = g.norm().item(), g_ref.norm().item()
    cos = (g @ g_ref).item() / (na * nb) if na > 0 and nb > 0 else float('nan')
    rows.append((cos, n, na, nb, (g - g_ref).norm().item()))
rows.sort(key=lambda r: (r[0] if r[0] == r[0] else -2))
for cos, n, na, nb, d in rows[:40]:
    print(f'{cos:9.5f} {n:45s} |ours| {na:.3e} |ref| {nb:.3e} |diff| {d:.3e}')
print('...')
tot_d = sum(r[4] ** 2 for r in rows) ** 0.5
tot_r = sum(r[3] ** 2 for r in rows) ** 0.5
print('total rel', tot_d / tot_r)
import torch as _t
A = _t.cat([grads[n].double().cpu().flatten() for n in ps.p]); Bv = _t.cat([(t.grad if t.grad is not None else _t.zeros_like(t)).double().flatten() for t in ps.p.values()])
print('GLOBAL COS', (A @ Bv).item() / (A.norm().item() * Bv.norm().item()))
rows.sort(key=lambda r: -r[4])
import statistics
cs = [r[0] for r in rows if r[0] == r[0]]
print('per-tensor cosine: median %.5f  p10 %.5f  min %.5f  (n=%d)' % (statistics.median(cs), sorted(cs)[len(cs) // 10], min(cs), len(cs)))
print('largest abs diffs:')
for cos, n, na, nb, d in rows[:15]:
    print(f'{cos:9.5f} {n:45s} |ours| {na:.3e} |ref| {nb:.3e} |diff| {d:.3e}')
