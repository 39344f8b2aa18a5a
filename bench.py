#!/usr/bin/env python
"""bench.py — M1 training throughput (volumes/s) on N B200s, one JSON line on rank 0.

  python bench.py --gpus 1 --steps K --warmup W                      our arm (sm_100a kernels via the C-ABI)
  torchrun ... bench.py --gpus N --steps K --warmup W                data parallel, one rank per GPU (NCCL)
  python bench.py --impl reference --gpus N --steps K --warmup W     the reference's CPU path (oracle port)

Workload (BASELINE.json configs[1]): full M1 — probabilistic + dense_skip + deep_supervision — training
step (4-pass forward, focal + KL, backward, Adam-AMSGrad), bf16, batch 8 per GPU, synthetic 20x160x160
volumes with 4 input channels (3 bpMRI + label channel). A "step" is one such training step.

Timed regions (CUDA events on the launch stream, barrier + synchronize on both sides, max over ranks):
  value      K steps, inputs resident in HBM; after its warm-up the model replays the step from two CUDA graphs
             ([forward, losses, backward] and [Adam, weight re-pack]; with N > 1 the NCCL all-reduce of the flat
             gradient buffer runs between them)
  e2e        K steps through the public API with pinned HOST inputs copied in and the loss read back every step
  roofline   the same K steps launched eagerly with a CUDA-event pair around every kernel family (per position
             in the step: median over the K steps), dominant conv family vs the measured bf16 peak
  cpu_baseline / --impl reference   the oracle port of the reference on the host cores (TF 2.5 cannot run here)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

README_CFG = dict(filters=(32, 64, 128, 256, 512),
                  strides=((1, 1, 1), (1, 2, 2), (1, 2, 2), (2, 2, 2), (2, 2, 2)),
                  kernel_sizes=((1, 3, 3), (1, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3)),
                  se_reduction=(8, 8, 8, 8, 8))
METRIC = "M1 train volumes/sec (20x160x160x3)"
# algorithmic conv FLOPs of one training volume (fwd 837.65 GMAC, fwd+bwd = 3x)
FLOP_PER_VOLUME = 5025.9e9


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="volumes per GPU")
    ap.add_argument("--dims", type=int, nargs=3, default=[20, 160, 160])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tc", action="store_true", help="debug: CUDA-core convolutions only")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"],
                    help="storage type of the activation values (fp16: the mode that meets the parity bounds)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (PyTorch/oneDNN fp32 port of the reference's arithmetic; TF 2.5 cannot run here)
# ------------------------------------------------------------------------------------------------
def oracle_step_fn(dims, threads):
    """Returns f() running ONE full training step (fwd 4 passes, losses, bwd, Adam) on one volume."""
    import torch
    from oracle import m1_oracle as O
    torch.set_num_threads(threads)
    cfg = O.default_config(dense_skip=True, deep_supervision=True, probabilistic=True, prob_latent_dims=(3, 2, 1, 0),
                           dropout_mode='monte-carlo', **{k: README_CFG[k] for k in ('filters', 'strides', 'kernel_sizes')})
    ps = O.ParamStore(dtype=torch.float32, seed=0, requires_grad=True)
    x, y = O.synthetic_batch(1, tuple(dims), dtype=torch.float32)
    state = {}
    step = [0]

    def f():
        step[0] += 1
        noise = O.Noise(step[0], torch.float32)
        for t in ps.p.values():
            t.grad = None
        r = O.train_loss(ps, cfg, x, y, noise, alpha=(0.75, 0.25), gamma=2.0, kl_weight=10.0)
        r['loss'].backward()
        with torch.no_grad():
            for n, t in ps.p.items():
                if t.grad is None:
                    continue
                m, v, vh = state.get(n) or (torch.zeros_like(t), torch.zeros_like(t), torch.zeros_like(t))
                w, m, v, vh = O.adam_amsgrad_step(t, t.grad, m, v, vh, step[0], 1e-3)
                t.copy_(w)
                state[n] = (m, v, vh)
        return float(r['loss'].detach())
    return f


def pick_cpu_sample(dims, budget_s, nsteps, threads):
    """Largest crop (H, W multiples of 16 so that the four stride-2 levels nest) of the volume whose
    nsteps training steps fit the budget; the conv cost is linear in the voxel count, so
    volumes/s = voxel fraction / step time."""
    probe = (dims[0], 32, 32)
    f = oracle_step_fn(probe, threads)
    f()
    t0 = time.time(); f(); t_probe = time.time() - t0
    per_voxel = t_probe / (probe[0] * probe[1] * probe[2])
    cands = [(dims[1], dims[2]), (dims[1], dims[2] // 2), (dims[1] // 2, dims[2] // 2), (80, 48), (48, 48), (32, 32)]
    for h, w in cands:
        if h % 16 == 0 and w % 16 == 0 and per_voxel * dims[0] * h * w * nsteps <= budget_s:
            return (dims[0], h, w)
    return (dims[0], 32, 32)


def cpu_baseline(dims, budget_s=30.0):
    threads = os.cpu_count() or 1
    sample = pick_cpu_sample(dims, budget_s, 2, threads)
    f = oracle_step_fn(sample, threads)
    f()                                             # warm-up (parameter creation, oneDNN primitives)
    t0 = time.time(); f(); dt = time.time() - t0
    frac = (sample[0] * sample[1] * sample[2]) / (dims[0] * dims[1] * dims[2])
    return {"value": frac / dt, "unit": "volumes/s", "cores": threads, "kind": "port",
            "sample": "1 training step (4-pass fwd + bwd + Adam, fp32 oneDNN) on one %dx%dx%d volume (%.3g of a "
                      "20x160x160 volume; cost is linear in voxels); TF 2.5 itself cannot run in this image"
                      % (sample + (frac,))}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    dims = tuple(args.dims)
    sample = pick_cpu_sample(dims, 150.0, args.steps + args.warmup, threads)
    f = oracle_step_fn(sample, threads)
    for _ in range(args.warmup):
        f()
    t0 = time.time()
    for _ in range(args.steps):
        f()
    dt = time.time() - t0
    frac = (sample[0] * sample[1] * sample[2]) / (dims[0] * dims[1] * dims[2])
    value = frac * args.steps / dt
    desc = ("each step = 1 training step on one %dx%dx%d volume (%.3g of the 20x160x160 volume), PyTorch-CPU oracle "
            "port of the reference (TF 2.5 is not installable here)" % (sample + (frac,)))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "volumes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "volumes/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def workload_config(args, world):
    return {"workload": "cfg2: full M1 (probabilistic + dense_skip + deep_supervision, monte-carlo dropout) training "
                        "step, %dx%dx%d volumes, 4 input channels (3 bpMRI + label ch), batch %d per GPU"
                        % (tuple(args.dims) + (args.batch,)),
            "global_batch": args.batch * world, "parallelism": "dp%d" % world,
            "step_launch": "whole training step captured once into a CUDA graph and replayed (M1_CUDA_GRAPH=0: eager)",
            "l2_flush": "not needed: every step streams >10 GB of activations, far above the 126 MB L2",
            "filters": list(README_CFG['filters']), "prob_latent_dims": [3, 2, 1, 0]}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import m1b200  # noqa: F401
    from m1b200.model import losses, optimizers, unets
    from m1b200.model.distribute import init_from_env
    from m1b200 import _lib

    rank, local, world = init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local)
    dims = tuple(args.dims)
    B = args.batch
    model = unets.networks.M1(dims, 4, 2, dropout_rate=0.5, dropout_mode='monte-carlo', att_sub_samp=((1, 1, 1),) * 4,
                              dense_skip=True, deep_supervision=True, probabilistic=True,
                              prob_latent_dims=(3, 2, 1, 0), summary=False, precision=args.precision, seed=0,
                              device=dev, use_tcgen05=not args.no_tc, **README_CFG)
    sched = optimizers.CosineDecayRestarts(1e-3, 1000, t_mul=2.0, m_mul=1.0, alpha=1e-3)
    model.compile(optimizer=optimizers.Adam(learning_rate=sched, amsgrad=True),
                  loss=[losses.Focal(alpha=[0.75, 0.25], gamma=2.0).loss, losses.EvidenceLowerBound().loss],
                  loss_weights=[1.0, 10.0])
    if world > 1:
        model.distribute()
    ctx = _lib.Context.get(local)

    # synthetic batch: whitened images N(0,1), ellipsoid lesion labels, pinned on the host
    g = torch.Generator().manual_seed(1234 + rank)
    D, H, W = dims
    xh = torch.randn((B, D, H, W, 4), generator=g)
    zz, yy, xx = torch.meshgrid(torch.arange(D), torch.arange(H), torch.arange(W), indexing='ij')
    lab = torch.zeros((B, D, H, W))
    for b in range(B):
        c = [int(torch.randint(0, s, (1,), generator=g)) for s in dims]
        r = float(torch.randint(4, 13, (1,), generator=g))
        lab[b] = (((zz - c[0]) * 2.0) ** 2 + (yy - c[1]) ** 2 + (xx - c[2]) ** 2 <= r * r).float()
    xh[..., 3] = lab
    yh = torch.stack([1 - lab, lab], -1)
    xh, yh = xh.pin_memory(), yh.pin_memory()
    xd, yd = xh.to(dev), yh.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, step_fn):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(nsteps):
            step_fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def dev_step():
        model.train_step(xd, yd)

    last = {}

    def e2e_step():
        x = xh.to(dev, non_blocking=True)
        y = yh.to(dev, non_blocking=True)
        r = model.train_step(x, y)
        last['loss'] = model.total_loss(r).cpu()       # device -> host read of the step's result

    for _ in range(args.warmup):
        dev_step()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    # ---- headline: K steps, inputs resident in HBM. After its warm-up the model replays the whole step from
    # a CUDA graph (M1.train_step); kernels launched = replays x kernels captured per step + eager launches.
    ctx.launch_count(reset=True)
    replays0 = getattr(model, "graph_replays", 0)
    ms = timed(args.steps, dev_step)
    launches = ctx.launch_count(reset=True)
    replays = getattr(model, "graph_replays", 0) - replays0
    if replays:
        launches += replays * model.launches_per_graph_step
    clk = clocks.stop() if rank == 0 else None
    # ---- end to end: pinned host inputs copied in, loss read back, every step
    e2e_step()
    ms_e2e = timed(args.steps, e2e_step)
    # ---- per-kernel-family durations: the same K steps once more, launched eagerly with a CUDA-event pair
    # around every launch family (a replayed graph cannot carry per-launch events)
    model.eng.prof = []
    ncu_range = os.environ.get("M1_CUDA_PROFILER_RANGE") == "1"     # ncu --profile-from-start off: only this pass
    if ncu_range:
        torch.cuda.profiler.start()
    ms_prof = timed(args.steps, dev_step)
    if ncu_range:
        torch.cuda.profiler.stop()
    prof, model.eng.prof = model.eng.prof, None

    value = world * B * args.steps / (ms / 1e3)
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- roofline of the dominant conv kernel family (CUDA-event durations inside the timed region)
    cats = {}
    if os.environ.get("M1_DUMP_PROF"):
        rows = sorted(((a.elapsed_time(b), cat, fl, str(lbl)) for cat, fl, a, b, lbl in prof), reverse=True)
        with open(os.environ["M1_DUMP_PROF"], "w") as fh:
            for t_ms, cat, fl, lbl in rows:
                fh.write("%9.3f ms  %-20s %8.2f TFLOP/s  %s\n" % (t_ms, cat, fl / 1e9 / max(t_ms, 1e-6), lbl))
    # every step issues the same launch sequence: take, per position in the step, the MEDIAN duration over the
    # K profiled steps (the eager pass is host-bound; a Python GC pause or a late launch otherwise lands in
    # whichever event pair happens to bracket it), then scale back to K steps
    times = [a.elapsed_time(b) for _, _, a, b, _ in prof]
    per_step = len(prof) // max(1, args.steps)
    if per_step and per_step * args.steps == len(prof) and args.steps > 1:
        for i in range(per_step):
            col = sorted(times[i + k * per_step] for k in range(args.steps))
            med = col[len(col) // 2]
            for k in range(args.steps):
                times[i + k * per_step] = med
    for (cat, fl, a, b, _), t_ms in zip(prof, times):
        c = cats.setdefault(cat, [0.0, 0.0, 0])
        c[0] += fl; c[1] += t_ms; c[2] += 1
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    breakdown = {k: {"launches": v[2] // args.steps, "ms_per_step": v[1] / args.steps,
                     "tflops": (v[0] / 1e12) / (v[1] / 1e3) if v[1] > 0 else None} for k, v in cats.items()}
    conv_cats = {k: v for k, v in cats.items() if k.startswith("conv_")}
    dom = max(conv_cats, key=lambda k: conv_cats[k][1]) if conv_cats else None
    roofline = None
    if dom:
        fl, t_ms, n = cats[dom]
        ach = (fl / 1e12) / (t_ms / 1e3)
        roofline = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": ach / peak_tf, "traffic": None,
                    "traffic_note": "per-launch DRAM bytes of the family's largest launches are in "
                                    "profiles/r01_summary.md (ncu --set full): sersp2 wgrad 0.56 GB vs 0.35 GB "
                                    "algorithmic; the family mixes 168 launch shapes, so no single figure applies",
                    "peak_source": peak_src,
                    "share_of_step": t_ms / sum(v[1] for v in cats.values()), "flop_per_launch_avg": fl / n,
                    "timed_with": "CUDA events around every launch family in an eager re-run of the same %d steps "
                                  "(%.2f ms/step; the headline replays the step as a CUDA graph)"
                                  % (args.steps, ms_prof / args.steps),
                    "conv_path_tflops_whole_step": (FLOP_PER_VOLUME * B * args.steps / 1e12) / (ms / 1e3),
                    "conv_path_frac_whole_step": (FLOP_PER_VOLUME * B * args.steps / 1e12) / (ms / 1e3) / peak_tf,
                    "breakdown": breakdown}
    out = {
        "metric": METRIC, "value": value, "unit": "volumes/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": workload_config(args, world),
        "clocks": clk,
        "e2e": {"value": e2e, "unit": "volumes/s", "h2d_bytes_per_step": int(xh.numel() * 4 + yh.numel() * 4),
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "loss": float(last['loss']),
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(dims)
    print(json.dumps(out))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
