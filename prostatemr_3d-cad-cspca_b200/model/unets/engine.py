"""Execution engine underneath the M1 model: device activations, a reverse tape, and one method
per layer family of the reference that launches the corresponding libm1b200 kernels (forward) and
records the kernels of its backward.

Host logic only: shapes, buffer lifetimes, which kernel to launch with which pointers. In TRACE
mode (no GPU needed) the same code path only propagates shapes and declares parameters - that is
how M1.__init__ discovers the parameter inventory, mirroring Keras' build-on-first-call."""
import numpy as np
import torch

from ... import _lib, ops
from ..._lib import CONV_FWD, CONV_TRANSPOSED, ENGINE_AUTO, ENGINE_SIMT, F32, grad_dtype

LRELU = 0.1
IN_EPS = 1e-3


class Act:
    """An NDHWC activation: shape, device tensor (None while tracing) and its gradient."""
    __slots__ = ("shape", "t", "g", "dtype", "needs_grad", "lc", "tw")

    def __init__(self, shape, dtype, t=None, needs_grad=True, lc=None):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = dtype
        self.t = t
        self.g = None
        self.tw = None      # fp16 mode, training: bf16 twin of t - the operand of the tensor-core weight gradients
        self.needs_grad = needs_grad
        # logical (reference) channel count; shape[-1] is the physical one, zero-padded to the tensor-core
        # granularity for the few-channel tensors (f/4 = 8 bottlenecks, 3-4 channel inputs, 1-3 channel latents)
        self.lc = int(lc) if lc is not None else self.shape[-1]

    @property
    def grid(self):
        return self.shape[1:4]

    @property
    def c(self):
        return self.shape[-1]


_code = _lib.code_of

PRECISIONS = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}


class InjectedNoise:
    """Explicit dropout uniforms / latent eps (parity runs): {(pass_name, site): tensor}."""

    def __init__(self, tensors):
        self.tensors = tensors
        self._dev = {}

    def _get(self, key, shape, device):
        if key not in self._dev:
            if key not in self.tensors:
                raise KeyError(f"no injected noise for {key}")
            t = torch.as_tensor(self.tensors[key]).to(device=device, dtype=torch.float32).contiguous()
            assert tuple(t.shape) == tuple(shape), (key, tuple(t.shape), tuple(shape))
            self._dev[key] = t
        return self._dev[key]

    def dropout(self, eng, pass_name, site, shape, rate):
        return ops.make_dropout(rate, self._get((pass_name, site), shape, eng.device)), \
            self._dev[(pass_name, site)]

    def normal(self, eng, pass_name, site, shape):
        return self._get((pass_name, site), shape, eng.device)


class PhiloxNoise:
    """Counter-based on-device noise: the dropout masks are never stored, the backward kernels
    regenerate them from (seed, stream id, element index)."""
    SITES = ("drope1", "drope2", "drope3", "drope4", "dropd3", "dropd2", "dropd1", "dropd0",
             "dropp3", "dropp2", "dropp1", "dropp0", "eps3", "eps2", "eps1", "eps0")
    PASSES = ("det", "q_sample", "q_mean", "p_z_q", "p_z_qmean", "p_sample")

    def __init__(self, seed=42, rank=0):
        self.seed = int(seed) + 1000003 * int(rank)
        self.step = 0
        self.step_dev = None      # device step counter while a CUDA graph of the step is being captured
        self.sub = 0              # offset on the step: pass i of a Monte-Carlo ensemble launched as ONE graph

    def _stream(self, pass_name, site):
        """Philox stream = step * 4096 + pass * 64 + site; under graph capture the step term is added on the
        device from step_dev (M1_PHILOX_STEP_STRIDE), so eager and replayed steps draw identical noise."""
        p = self.PASSES.index(pass_name) if pass_name in self.PASSES else len(self.PASSES)
        step = (0 if self.step_dev is not None else self.step) + self.sub
        return (step * 64 + p) * 64 + self.SITES.index(site)

    def dropout(self, eng, pass_name, site, shape, rate):
        # training: the forward gate kernel leaves 1 keep-bit per element for its two backward kernels (Philox
        # costs more instructions than the rest of those kernels); inference regenerates nothing anyway
        mask = None
        if eng.record and shape[-1] % 8 == 0:
            mask = torch.empty((int(np.prod(shape)) + 7) // 8, dtype=torch.uint8, device=eng.device)
        return ops.make_dropout(rate, None, self.seed, self._stream(pass_name, site), self.step_dev, mask), mask

    def normal(self, eng, pass_name, site, shape):
        out = torch.empty(shape, dtype=torch.float32, device=eng.device)
        ops.philox_normal(eng.ctx, self.seed, self._stream(pass_name, site), out, self.step_dev)
        return out


class SETrunk:
    """dropout-independent part of an SE tail (Engine.se_trunk): the two raw conv outputs, their statistics, the
    squeeze / excite results and the parameter views"""
    __slots__ = ("raw3", "raw4", "name", "c", "cr", "g3", "b3", "g4", "b4", "w6", "b6", "w7", "b7",
                 "st3", "st4", "pool", "hidden", "gate")


class LazyHead:
    """A final 1x1x1 logits convolution that has not been launched yet: the loss kernel fuses it (K8 reads
    the decoder features directly and emits d(features), dW, db); `materialize` is the unfused fallback."""

    def __init__(self, eng, feat, name, nc):
        self.feat, self.name, self.nc = feat, name, nc
        self.w = eng.p(name + "/kernel", (1, 1, 1, feat.lc, nc), "kernel", (1, 1, 1, feat.c, nc),
                       {3: np.arange(feat.lc)})
        self.b = eng.p(name + "/bias", (nc,), "bias")
        self.shape = feat.shape[:-1] + (nc,)

    def materialize(self, eng):
        out, = eng.conv([self.feat], [(self.name, self.nc)], (1, 1, 1), out_dtype=torch.float32)
        return out


class Engine:
    def __init__(self, params, precision="fp16", device=None, use_tcgen05=True):
        """precision: storage type of the activation VALUES - 'fp16' (tcgen05 convolutions on fp16 values and
        weights, bf16 activation gradients: the mode that meets the 2e-2 / 1e-3 parity bounds at tensor-core
        speed), 'bf16' (values, gradients and tensor-core weights all bf16) or 'fp32' (CUDA-core convolutions,
        the 1e-4 parity mode). Accumulation, statistics, parameters and parameter gradients are always fp32."""
        assert precision in PRECISIONS, precision
        self.params = params
        self.precision = precision
        self.act_dtype = PRECISIONS[precision]
        self.use_tc = use_tcgen05 and precision != "fp32"
        # tcgen05.mma.kind::f16 traps when its two operands have different formats (measured on B200, both ways).
        # fp16 mode therefore runs   forward        fp16 activations x fp16 weight pack
        #                            data gradient  bf16 gradients   x bf16 weight pack (w_dtype 0 = as gathered)
        #                            weight gradient bf16 TWIN of the activations x bf16 gradients
        # The twin is written by the kernel that produces the activation (second store of the same registers)
        # or, for tensors produced by a convolution epilogue / the input, by one cast the first time a weight
        # gradient needs it.
        self.twins = precision == "fp16" and self.use_tc
        self.dgrad_w = 0
        self.tracing = device is None
        self.device = device
        self.ctx = None if self.tracing else _lib.Context.get(torch.device(device).index or 0)
        self.tape = []
        self.record = True
        self.noise = None
        self.packs = {}            # key -> (desc, [kernel names], packed tensor)
        self.prof = None           # bench.py: list of (category, flops, start_event, end_event)
        self.autotune = True       # one-off timing of candidate tcgen05 tilings per layer shape
        self.tuned = {}
        self.conv_flops = 0        # algorithmic MACs*2 of the convolutions launched (forward only)
        self.bwd_flops = 0         # ... of the data-gradient and weight-gradient launches actually executed
        self.trace_log = []        # TRACE mode: one record per convolution launch (tools/list_launches.py, tests)
        import os
        # passes that share their input compute the dropout-free head of the encoder once (M1_SHARE_TRUNK=0: as the
        # reference graph, every pass on its own - same results up to the rounding order of the summed gradients)
        self.share_trunk = os.environ.get("M1_SHARE_TRUNK", "1") != "0"
        self.shared = {}
        # Weight gradients on a SIDE stream: dW of a layer and the data gradient of the same layer only share inputs,
        # so the weight-gradient kernel (often a thin, latency-bound launch that leaves most SMs idle) runs
        # concurrently with the rest of the backward walk and is joined before anything consumes the parameter
        # gradients (bucket all-reduce, optimizer). The side stream has its own m1_ctx (scratch buffers); the tensors
        # a side launch reads are kept alive until the join. Captured into the step's CUDA graph as a fork / join.
        self.side_on = os.environ.get("M1_WGRAD_STREAM", "1") != "0"
        self.side, self.ctx_side, self._side_keep, self._side_dirty = None, None, [], False

    # ---- helpers -----------------------------------------------------------------------------
    def new(self, shape, dtype=None, zero=False):
        dtype = dtype or self.act_dtype
        if self.tracing:
            return None
        return (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)

    def input(self, t, needs_grad=False, lc=None):
        if self.tracing:
            return Act(t, self.act_dtype, None, needs_grad, lc)  # t is a shape
        return Act(t.shape, t.dtype, t, needs_grad, lc)

    def padc(self, c):
        """physical channel count of a few-channel activation: multiples of 16 feed the tensor cores"""
        return -(-c // 16) * 16 if self.use_tc else c

    def p(self, name, shape, kind, pshape=None, index=None):
        self.params.declare(name, shape, kind, pshape, index)
        return None if self.tracing else self.params.view(name)

    def pg(self, name):
        return self.params.grad(name)

    def grad_buffer(self, act, zero=False):
        """(tensor, accumulate): allocates act.g on first use."""
        if act.g is None:
            act.g = (torch.zeros if zero else torch.empty)(act.shape, dtype=grad_dtype(act.dtype), device=self.device)
            return act.g, False
        return act.g, True

    def new_twin(self, act):
        """bf16 twin buffer for an activation being produced (None unless fp16-mode training)"""
        if self.twins and self.record and not self.tracing and act.dtype == torch.float16:
            act.tw = self.new(act.shape, torch.bfloat16)
        return act.tw

    def x16(self, act):
        """the tensor a tensor-core weight gradient reads for activation `act`: its bf16 twin in fp16 mode. A twin that
        no producing kernel wrote (Conv3DTranspose outputs, the network input) is cast here - on the side stream of
        the weight gradients when that is on: only they ever read twins."""
        if not self.twins or act.dtype != torch.float16:
            return act.t
        if act.tw is None:
            act.tw = self.new(act.shape, torch.bfloat16)
            src, dst = act.t, act.tw
            if self.side_on and self.prof is None:
                self._on_side([src, dst], lambda c: ops.cast(c, src, dst))
            else:
                ops.cast(self.ctx, src, dst)
        return act.tw

    def new_grad(self, act):
        """uninitialised gradient tensor of an activation (bf16 for fp16 values, else the value type)"""
        return self.new(act.shape, grad_dtype(act.dtype))

    def _timed(self, cat, flops, fn, label=None, nbytes=0):
        """bench.py's per-family profile: flops = ALGORITHMIC (reference-channel) FLOPs of the convolution launches,
        nbytes = ALGORITHMIC bytes of the bandwidth kernels (tensor passes x elements x element size)"""
        if self.prof is None:
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        self.prof.append((cat, flops, a, b, label, nbytes))

    def begin(self, record=True):
        self.tape = []
        self.record = record
        self.conv_flops = 0
        self.bwd_flops = 0
        self.param_uses = {}
        self.shared = {}           # per-step cache of sub-graphs shared between passes (M1Core: stem + serse1 trunk)

    def _rec(self, fn, names):
        """record a backward closure and the parameters whose gradients it contributes to"""
        self.tape.append((fn, names))
        for n in names:
            self.param_uses[n] = self.param_uses.get(n, 0) + 1

    def backward(self, on_param_done=None):
        for fn, names in reversed(self.tape):
            fn()
            if on_param_done is not None:
                for n in names:
                    on_param_done(n)
        self.tape = []
        self.join_side()

    # ---- side stream of the weight gradients ------------------------------------------------------
    def _on_side(self, keep, fn):
        """run fn(ctx) on the side stream after everything enqueued so far on the current stream"""
        if self.side is None:
            self.side = torch.cuda.Stream(device=self.device)
            self.ctx_side = _lib.Context(torch.device(self.device).index or 0)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            fn(self.ctx_side)
        self._side_keep.extend(keep)
        self._side_dirty = True

    def join_side(self):
        """the current stream waits for the side stream; the tensors its launches read may be released afterwards"""
        if self._side_dirty:
            ev = torch.cuda.Event()
            ev.record(self.side)
            torch.cuda.current_stream(self.device).wait_event(ev)
            self._side_dirty = False
        self._side_keep = []

    def launch_total(self):
        return self.ctx.launch_count() + (self.ctx_side.launch_count() if self.ctx_side is not None else 0)

    # ---- K1/K2/K3 convolutions -----------------------------------------------------------------
    def _gather(self, cat, mode, batch, in_dhw, out_dhw, k, s, pad, src_t, src_c, ws, wstr, bias, out_t, out_c,
                acc, key, w_by_src=False, w_dtype=0, lflops=None):
        """One convolution-shaped launch family: outs[j] (+)= gather(concat(src)) * ws[j] (+ bias[j]).
        The tcgen05 engine takes it when every gathered tensor has a multiple of 16 channels; a
        concatenation that mixes such tensors with odd ones (the 1-3 channel latents, R:networks.py:653)
        is split into runs - tensor-core launch for the aligned run, CUDA cores for the rest."""
        taps = int(np.prod(k))
        vox = batch * int(np.prod(in_dhw if mode == CONV_TRANSPOSED else out_dhw))
        act_code, out_code = _code(src_t[0].dtype), _code(out_t[0].dtype)
        runs = [(0, len(src_t))]
        if self.use_tc and act_code != F32 and out_code != F32 and not w_by_src:
            ok = [c % 16 == 0 for c in src_c]
            if any(ok) and not all(ok):
                runs, i = [], 0
                while i < len(ok):
                    j = i
                    while j < len(ok) and ok[j] == ok[i]:
                        j += 1
                    runs.append((i, j))
                    i = j
        offs = np.cumsum([0] + list(src_c))
        first = True
        for (i0, i1) in runs:
            sub_w = list(ws) if w_by_src else [w.view(-1)[int(offs[i0]) * st[1]:] for w, st in zip(ws, wstr)]
            a_flags = list(acc) if first else [True] * len(out_t)
            d = ops.conv_desc(mode, batch, in_dhw, out_dhw, k, s, pad, list(src_c[i0:i1]), list(out_c), list(wstr),
                              accumulate=a_flags, act_dtype=act_code, out_dtype=out_code,
                              engine=ENGINE_AUTO if self.use_tc else ENGINE_SIMT, w_by_src=w_by_src, w_dtype=w_dtype)
            packed = None
            if self.use_tc and ops.conv3d_tc_supported(d):
                pk = key + (i0,)
                ent = self.packs.get(pk)
                if ent is None:
                    ent = (d, sub_w, ops.conv3d_pack_weights(self.ctx, d, sub_w))
                    self.packs[pk] = ent
                packed = ent[2]
            fl = 2 * vox * taps * int(sum(src_c[i0:i1])) * int(sum(out_c))
            if lflops is not None:            # algorithmic FLOPs (reference channel counts), split over the runs
                fl = int(lflops * sum(src_c[i0:i1]) / max(1, sum(src_c)))
            b = bias if first else None
            if packed is not None and self.autotune:
                tk = ("conv", mode, batch, tuple(in_dhw), tuple(out_dhw), k, s, tuple(src_c[i0:i1]), tuple(out_c),
                      bool(w_by_src), act_code, out_code, w_dtype)
                var = self.tuned.get(tk)
                if var is None:
                    var = self._tune_conv(d, list(src_t[i0:i1]), sub_w, list(out_t), packed)
                    self.tuned[tk] = var
                d.tune[0] = var
            # algorithmic bytes: every gathered and produced tensor once (+ the old contents of accumulating outputs)
            nb = 0 if self.prof is None else (
                sum(t.numel() * t.element_size() for t in src_t[i0:i1]) +
                sum(t.numel() * t.element_size() * (2 if a else 1) for t, a in zip(out_t, a_flags)) +
                sum(w.numel() * 2 for w in ([] if w_by_src else ws)))
            self._timed(cat + ("_tcgen05" if packed is not None else "_simt"), fl,
                        lambda: ops.conv3d(self.ctx, d, list(src_t[i0:i1]), sub_w, b, list(out_t), packed),
                        label=(key, tuple(src_c[i0:i1]), tuple(out_c), tuple(out_dhw), tuple(k), tuple(s)), nbytes=nb)
            first = False

    def conv(self, srcs, layers, k, s=(1, 1, 1), transposed=False, out_dtype=None, pad_out=None, feeds_norm=False):
        """Conv3D (possibly several layers reading the same input fused along Cout) or
        Conv3DTranspose over the virtual concatenation of `srcs`.
        layers: [(param_prefix, cout)]; returns one Act per layer. pad_out[j]: zero-pad layer j's output
        channels to the tensor-core granularity (the parameters keep their reference shape logically).
        feeds_norm: the outputs go straight into an InstanceNorm; the bias gradient is then identically zero
        (d/db of IN(conv + b) = 0: the norm removes the per-channel mean) and BiasAddGrad is not launched."""
        k, s = tuple(k), tuple(s)
        pad_out = pad_out or [False] * len(layers)
        lcin = sum(a.lc for a in srcs)
        lco = [co for _, co in layers]
        layers = [(n, self.padc(co) if po else co) for (n, co), po in zip(layers, pad_out)]
        cin = sum(a.c for a in srcs)
        # logical -> physical position of every gathered channel over the virtual concatenation
        cin_index = np.concatenate([off + np.arange(a.lc) for off, a in
                                    zip(np.cumsum([0] + [a.c for a in srcs])[:-1], srcs)])
        batch, in_dhw = srcs[0].shape[0], srcs[0].grid
        for a in srcs:
            assert a.grid == in_dhw, "concatenated tensors must share the grid"
        out_dtype = out_dtype or self.act_dtype
        if transposed:
            assert len(layers) == 1
            out_dhw = tuple(in_dhw[i] * s[i] for i in range(3))
            pad = tuple(ops.same_pads(out_dhw[i], k[i], s[i])[1] for i in range(3))
            shapes = [k + (layers[0][1], cin)]
            lshapes = [k + (lco[0], lcin)]
            index = [{3: np.arange(lco[0]), 4: cin_index}]
            wstr = [(layers[0][1] * cin, 1, cin)]
            mode = CONV_TRANSPOSED
        else:
            geo = [ops.same_pads(in_dhw[i], k[i], s[i]) for i in range(3)]
            out_dhw = tuple(g[0] for g in geo)
            pad = tuple(g[1] for g in geo)
            shapes = [k + (cin, co) for _, co in layers]
            lshapes = [k + (lcin, c) for c in lco]
            index = [{3: cin_index, 4: np.arange(c)} for c in lco]
            wstr = [(cin * co, co, 1) for _, co in layers]
            mode = CONV_FWD
        ws = [self.p(n + "/kernel", lshp, "kernel", shp, ix)
              for (n, _), shp, lshp, ix in zip(layers, shapes, lshapes, index)]
        bs = [self.p(n + "/bias", (c,), "bias", (co,), {0: np.arange(c)}) for (n, co), c in zip(layers, lco)]
        outs = [Act((batch,) + out_dhw + (co,), out_dtype, self.new((batch,) + out_dhw + (co,), out_dtype), lc=c)
                for (_, co), c in zip(layers, lco)]
        taps = k[0] * k[1] * k[2]
        vox = batch * (np.prod(in_dhw) if transposed else np.prod(out_dhw))
        self.conv_flops += 2 * int(vox) * taps * lcin * sum(lco)      # algorithmic (reference) FLOPs
        if self.tracing:
            self.trace_log.append(dict(names=[n for n, _ in layers], transposed=bool(transposed), batch=batch,
                                       in_dhw=tuple(in_dhw), out_dhw=tuple(out_dhw), kernel=k, stride=s, pad=pad,
                                       src_c=[a.c for a in srcs], out_c=[co for _, co in layers],
                                       out_fp32=out_dtype == torch.float32,
                                       needs_dgrad=all(a.needs_grad for a in srcs), feeds_norm=bool(feeds_norm),
                                       flops=2 * int(vox) * taps * lcin * sum(lco)))
            return outs
        self._gather("conv_fwd", mode, batch, in_dhw, out_dhw, k, s, pad, [a.t for a in srcs], [a.c for a in srcs],
                     ws, wstr, bs, [o.t for o in outs], [co for _, co in layers], [False] * len(layers),
                     ("fwd",) + tuple(n for n, _ in layers), lflops=2 * int(vox) * taps * lcin * sum(lco))
        if self.record:
            self._rec(lambda: self._conv_bwd(srcs, layers, outs, k, s, pad, in_dhw, out_dhw, transposed, ws, wstr,
                                             not feeds_norm),
                      [n + sfx for n, _ in layers for sfx in ("/kernel", "/bias")])
        return outs

    def refresh_packs(self):
        """Re-derive the 16-bit operand packs from the fp32 master weights (after an optimizer step): ONE launch for
        all packs (m1_pack_plan); the plan is rebuilt whenever a new pack has appeared (never during graph capture)."""
        if not self.packs:
            return
        plan = getattr(self, "_pack_plan", None)
        if plan is None or plan.n != len(self.packs):
            if torch.cuda.is_current_stream_capturing():
                for d, ws, packed in self.packs.values():
                    ops.conv3d_pack_weights_into(self.ctx, d, ws, packed)
                return
            plan = self._pack_plan = ops.PackPlan(self.ctx, list(self.packs.values()))
        plan.run()

    # (max voxels per K brick, taps sharing a dY tile [0 = kw if it fits], stage cap, M tiles per CTA)
    WG_CANDIDATES = ((128, 0, 2, 1), (128, 1, 2, 1), (64, 0, 2, 1), (64, 1, 2, 1), (128, 0, 3, 1), (64, 0, 3, 1),
                     (64, 0, 4, 1), (64, 1, 4, 1), (128, 1, 3, 1), (128, 1, 2, 2), (64, 1, 2, 2), (64, 0, 2, 2))

    # SHIFT-mode tilings of the weight-gradient kernel (tune[1] == 2; M1_WG_SHIFT=0 removes them)
    WG_SHIFT_CANDIDATES = ((192, 2, 3, 1), (192, 2, 2, 1), (96, 2, 4, 1), (192, 2, 2, 2))

    def _wgrad(self, d, srcs_t, douts_t, dws, dbs, fl, label=None):
        on_tc = self.use_tc and ops.conv3d_wgrad_tc_supported(d)
        if on_tc and self.autotune:
            key = ("wgrad", label, tuple(srcs_t[0].shape[:4]), tuple(t.shape[-1] for t in srcs_t),
                   tuple(t.shape[-1] for t in douts_t))
            cfg = self.tuned.get(key)
            if cfg is None:
                cfg = self._tune_wgrad(d, srcs_t, douts_t, dws)
                self.tuned[key] = cfg
            d.tune[0], d.tune[1], d.tune[2], d.tune[3] = cfg
        nb = 0 if self.prof is None else (sum(t.numel() * t.element_size() for t in list(srcs_t) + list(douts_t)) +
                                          sum(w.numel() * 8 for w in dws))
        if self.side_on and self.prof is None and on_tc:
            self._on_side(list(srcs_t) + list(douts_t),
                          lambda c: ops.conv3d_wgrad(c, d, srcs_t, douts_t, dws, dbs))
            return
        if self.twins and any(t.dtype == torch.bfloat16 for t in srcs_t):
            self.join_side()   # a main-stream weight gradient that reads a twin the side stream may just have cast
        self._timed("conv_wgrad_tcgen05" if on_tc else "conv_wgrad_simt", fl, nbytes=nb,
                    fn=lambda: ops.conv3d_wgrad(self.ctx, d, srcs_t, douts_t, dws, dbs),
                    label=(label, tuple(t.shape[-1] for t in srcs_t), tuple(t.shape[-1] for t in douts_t),
                           tuple(srcs_t[0].shape[1:4])))

    def _tune_conv(self, d, srcs_t, ws, outs_t, packed):
        """One-off per launch shape: time the variants of the tcgen05 convolution engine on scratch outputs and keep
        the fastest - 1: one TMA box per filter tap, one tile per CTA; 2: one halo tile shared by the in-plane taps
        (stride-1 gathers with in-plane taps only); 3: multi-tile CTAs whose TMA ring streams across tile
        boundaries (what short-K launches need: 1x1x1 convolutions, phases of transposed convolutions)."""
        import os
        cands = [1]
        d.tune[0] = 2
        if ops.conv3d_halo_engine(d):
            cands.append(2)
        if os.environ.get("M1_CONV_MULTI_TUNE", "1") == "1":
            cands.append(3)
        if len(cands) == 1:
            d.tune[0] = 0
            return 1
        scratch = [torch.empty_like(o) for o in outs_t]
        best, best_t = 1, float("inf")
        for var in cands:
            d.tune[0] = var
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ops.conv3d(self.ctx, d, srcs_t, ws, None, scratch, packed)
                b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b))
            t = sorted(ts[1:])[1]                    # second fastest of 4 warm runs: robust to one outlier either way
            if t < best_t * (0.97 if var == 3 else 1.0):     # the multi-tile variant has to win clearly
                best, best_t = var, t
        d.tune[0] = 0
        return best

    def _tune_wgrad(self, d, srcs_t, douts_t, dws):
        """One-off per layer shape: time the candidate tilings of the tcgen05 weight-gradient kernel on
        scratch gradient buffers (CUDA events) and keep the fastest."""
        scratch = [torch.zeros(w.numel(), dtype=torch.float32, device=self.device) for w in dws]
        best, best_t = (0, 0, 0, 0), float("inf")
        import os
        cands = self.WG_CANDIDATES
        if os.environ.get("M1_WG_SHIFT", "1") == "1":
            cands = cands + self.WG_SHIFT_CANDIDATES
        for cand in cands:
            d.tune[0], d.tune[1], d.tune[2], d.tune[3] = cand
            if not ops.conv3d_wgrad_tc_supported(d):
                continue
            ts = []
            for _ in range(4):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ops.conv3d_wgrad(self.ctx, d, srcs_t, douts_t, scratch, None)
                b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b))
            t = sorted(ts[1:])[1]                    # median of 3 warm runs
            if t < best_t:
                best, best_t = cand, t
        d.tune[0], d.tune[1], d.tune[2], d.tune[3] = 0, 0, 0, 0
        return best

    def _conv_bwd(self, srcs, layers, outs, k, s, pad, in_dhw, out_dhw, transposed, ws, wstr, bias_grad=True):
        live = [j for j, o in enumerate(outs) if o.g is not None]
        if not live:
            return
        batch = srcs[0].shape[0]
        cin = sum(a.c for a in srcs)
        auto = ENGINE_AUTO if self.use_tc else ENGINE_SIMT
        taps = int(np.prod(k))
        need = [a for a in srcs if a.needs_grad]
        assert not need or len(need) == len(srcs), "mixed needs_grad inside one concatenation"
        # algorithmic (reference-channel) FLOPs of what runs below: one weight gradient, one data gradient if needed
        lvox = batch * int(np.prod(in_dhw if transposed else out_dhw))
        lfl = 2 * lvox * taps * sum(a.lc for a in srcs) * sum(outs[j].lc for j in live)
        self.bwd_flops += lfl * (2 if need else 1)
        if not transposed:
            # ---- wgrad: dW_j[tap, r, n] += gathered(src)[r] * dout_j[n] for all layers j of the fused launch in
            # one call (the gathered operand is streamed once); BiasAddGrad per layer
            cos = [layers[j][1] for j in live]
            # fp16 mode: 16-bit output gradients are bf16, so the activations are read through their bf16 twins
            # (one operand format per MMA); fp32-gradient heads (CUDA cores) read the fp16 activations
            xs = [self.x16(a) for a in srcs] if outs[live[0]].g.dtype != torch.float32 else [a.t for a in srcs]
            d = ops.conv_desc(CONV_FWD, batch, in_dhw, out_dhw, k, s, pad, [a.c for a in srcs], cos,
                              [wstr[j] for j in live], act_dtype=_code(xs[0].dtype),
                              out_dtype=_code(outs[live[0]].g.dtype), engine=auto)
            self._wgrad(d, xs, [outs[j].g for j in live],
                        [self.pg(layers[j][0] + "/kernel") for j in live],
                        [self.pg(layers[j][0] + "/bias") for j in live] if bias_grad else None, lfl,
                        layers[live[0]][0])
            # ---- dgrad: [dx_s for every gathered tensor] (+)= convT(dout_j, W_j): ONE launch per layer j whose
            # produced channels are split over the gradients of the concatenated tensors
            if need:
                offs = np.cumsum([0] + [a.c for a in srcs])[:-1]
                bufs, accs = zip(*[self.grad_buffer(a) for a in srcs])
                cos = [layers[j][1] for j in live]
                if len(live) > 1 and all(c % 16 == 0 for c in cos) and self.use_tc:
                    # all fused layers in ONE launch: K runs over [dout_j ...] (per-(produced, gathered) weights)
                    wv = [ws[j].view(-1)[int(off) * layers[j][1]:] for off in offs for j in live]
                    self._gather("conv_dgrad", CONV_TRANSPOSED, batch, out_dhw, in_dhw, k, s, pad,
                                 [outs[j].g for j in live], cos, wv, [(cin * c, 1, c) for c in cos], None,
                                 list(bufs), [a.c for a in srcs], list(accs),
                                 ("dgrad",) + tuple(layers[j][0] for j in live), w_by_src=True, w_dtype=self.dgrad_w,
                                 lflops=lfl)
                else:
                    for i, j in enumerate(live):
                        co = layers[j][1]
                        wv = [ws[j].view(-1)[int(off) * co:] for off in offs]
                        self._gather("conv_dgrad", CONV_TRANSPOSED, batch, out_dhw, in_dhw, k, s, pad, [outs[j].g],
                                     [co], wv, [(cin * co, 1, co)] * len(srcs), None, list(bufs),
                                     [a.c for a in srcs], list(accs) if i == 0 else [True] * len(srcs),
                                     ("dgrad", layers[j][0]), w_dtype=self.dgrad_w if outs[j].g.dtype != torch.float32 else 0,
                                     lflops=lfl * outs[j].lc // max(1, sum(outs[q].lc for q in live)))
        else:
            co = layers[0][1]
            dy = outs[0].g
            # ---- wgrad with swapped roles: dWt[tap, co, ci] += dy[i*s + k - pad, co] * x[i, ci]
            gk = self.pg(layers[0][0] + "/kernel").view(-1)
            off = 0
            for a in srcs:
                xa = self.x16(a)
                d = ops.conv_desc(CONV_FWD, batch, out_dhw, in_dhw, k, s, pad, [co], [a.c], [(co * cin, cin, 1)],
                                  act_dtype=_code(dy.dtype), out_dtype=_code(xa.dtype), engine=auto)
                self._wgrad(d, [dy], [xa], [gk[off:]], None, lfl * a.lc // max(1, sum(q.lc for q in srcs)),
                            layers[0][0] + "(T)")
                off += a.c
            if bias_grad:
                gb = self.pg(layers[0][0] + "/bias")
                if self.side_on and self.prof is None:           # reads dy only: next to the weight gradients
                    self._on_side([dy], lambda c: ops.bias_grad(c, dy, gb))
                else:
                    ops.bias_grad(self.ctx, dy, gb)
            # ---- dgrad: dx_s[i, ci] (+)= sum_k dy[i*s + k - pad, co] * Wt[k, co, off + ci] (strided FWD gather);
            # tensors with odd channel counts (latents) go to the CUDA cores, aligned ones to tcgen05
            if need:
                off = 0
                for idx, a in enumerate(srcs):
                    gbuf, acc = self.grad_buffer(a)
                    self._gather("conv_dgrad", CONV_FWD, batch, out_dhw, in_dhw, k, s, pad, [dy], [co],
                                 [ws[0].view(-1)[off:]], [(co * cin, cin, 1)], None, [gbuf], [a.c], [acc],
                                 ("dgradT", layers[0][0], idx), w_dtype=self.dgrad_w,
                                 lflops=lfl * a.lc // max(1, sum(q.lc for q in srcs)))
                    off += a.c
        for o in outs:
            o.g = None

    # ---- K4 instance norm + activation ---------------------------------------------------------
    def inorm_act(self, x, name, slope):
        c = x.c
        ix = {0: np.arange(x.lc)}
        gamma = self.p(name + "/gamma", (x.lc,), "gamma", (c,), ix)     # padded channels: gamma = beta = 0
        beta = self.p(name + "/beta", (x.lc,), "beta", (c,), ix)
        y = Act(x.shape, x.dtype, self.new(x.shape, x.dtype), lc=x.lc)
        if self.tracing:
            return y
        stats = self.new((x.shape[0], c, 2), torch.float32)
        self.new_twin(y)
        def fwd():
            ops.inorm_stats(self.ctx, x.t, stats, IN_EPS)
            ops.inorm_act_fwd(self.ctx, x.t, stats, gamma, beta, slope, y.t, y.tw)
        es = x.t.element_size()
        nel = x.t.numel()
        self._timed("inorm_fwd", 0, fwd, nbytes=3 * nel * es)              # stats read + apply read + write

        def bwd():
            if y.g is None:
                return
            gbuf, acc = self.grad_buffer(x)
            self._timed("inorm_bwd", 0, nbytes=5 * nel * es, fn=lambda: ops.inorm_act_bwd(
                self.ctx, y.g, x.t, stats, gamma, beta, slope, gbuf, acc, self.pg(name + "/gamma"),
                self.pg(name + "/beta")))
            y.g = None
        if self.record:
            self._rec(bwd, [name + "/gamma", name + "/beta"])
        return y

    # ---- K5 SE tail: norm3/norm4 + squeeze + excite + gate*residual + lrelu + dropout -----------
    def se_tail(self, raw3, raw4, name, reduction, drop):
        """drop: None or (pass_name, site, rate)."""
        return self.se_gate(self.se_trunk(raw3, raw4, name, reduction), drop)

    def se_trunk(self, raw3, raw4, name, reduction):
        """Everything of the SE tail that does not depend on the dropout draw: statistics of raw3 / raw4 (norm3,
        norm4), squeeze, excite. Several `se_gate` calls may follow on one trunk (passes of the probabilistic model
        that share their input differ only in the dropout mask applied to the block output)."""
        c = raw3.c
        cr = c // reduction
        tr = SETrunk()
        tr.raw3, tr.raw4, tr.name, tr.c, tr.cr = raw3, raw4, name, c, cr
        tr.g3 = self.p(name + "/norm3/gamma", (c,), "gamma")
        tr.b3 = self.p(name + "/norm3/beta", (c,), "beta")
        tr.g4 = self.p(name + "/norm4/gamma", (c,), "gamma")
        tr.b4 = self.p(name + "/norm4/beta", (c,), "beta")
        tr.w6 = self.p(name + "/conv6/kernel", (1, 1, 1, c, cr), "se_kernel")
        tr.b6 = self.p(name + "/conv6/bias", (cr,), "se_bias")
        tr.w7 = self.p(name + "/conv7/kernel", (1, 1, 1, cr, c), "se_kernel")
        tr.b7 = self.p(name + "/conv7/bias", (c,), "se_bias")
        if self.tracing:
            return tr
        n = raw3.shape[0]
        f32 = torch.float32
        tr.st3, tr.st4 = self.new((n, c, 2), f32), self.new((n, c, 2), f32)
        tr.pool, tr.hidden, tr.gate = self.new((n, c), f32), self.new((n, cr), f32), self.new((n, c), f32)

        def fwd():
            ops.inorm_stats(self.ctx, raw3.t, tr.st3, IN_EPS)
            ops.inorm_stats(self.ctx, raw4.t, tr.st4, IN_EPS)
            ops.se_excite_fwd(self.ctx, tr.pool, tr.w6, tr.b6, tr.w7, tr.b7, tr.hidden, tr.gate, tr.st3, tr.g3,
                              tr.b3)                                                   # squeeze folded in
        nel, es = raw3.t.numel(), raw3.t.element_size()
        self._timed("se_tail_fwd", 0, fwd, nbytes=2 * nel * es)          # the two statistics reads
        return tr

    def se_gate(self, tr, drop):
        """out = dropout(lrelu(norm3(raw3) * gate * norm4(raw4))) on a trunk; drop: None or (pass_name, site, rate)."""
        raw3, raw4, name = tr.raw3, tr.raw4, tr.name
        out = Act(raw3.shape, raw3.dtype, self.new(raw3.shape, raw3.dtype))
        if self.tracing:
            return out
        n, c = raw3.shape[0], tr.c
        f32 = torch.float32
        self.new_twin(out)
        if drop is not None and drop[2] > 0.0:
            dr, keep_alive = self.noise.dropout(self, drop[0], drop[1], raw3.shape, drop[2])
        else:
            dr, keep_alive = ops.make_dropout(0.0), None
        nel, es = raw3.t.numel(), raw3.t.element_size()
        self._timed("se_tail_fwd", 0, nbytes=3 * nel * es, fn=lambda: ops.se_gate_fwd(          # 2 reads, 1 write
            self.ctx, raw3.t, raw4.t, tr.st3, tr.st4, tr.g3, tr.b3, tr.g4, tr.b4, tr.gate, dr, out.t, out.tw))

        def bwd():
            if out.g is None:
                return
            _ = keep_alive
            red = self.new((n, c, 5), f32)
            dgate, dpool = self.new((n, c), f32), self.new((n, c), f32)
            # a second gate on the same trunk adds to the gradients the first one left in raw3.g / raw4.g
            acc = raw3.g is not None
            assert acc == (raw4.g is not None)
            if not acc:
                raw3.g = self.new_grad(raw3)
                raw4.g = self.new_grad(raw4)

            def run():
                ops.se_gate_bwd_reduce(self.ctx, out.g, raw3.t, raw4.t, tr.st3, tr.st4, tr.g3, tr.b3, tr.g4, tr.b4,
                                       tr.gate, dr, red, dgate)
                ops.se_excite_bwd(self.ctx, dgate, tr.pool, tr.hidden, tr.gate, tr.w6, tr.w7, dpool,
                                  self.pg(name + "/conv6/kernel"), self.pg(name + "/conv6/bias"),
                                  self.pg(name + "/conv7/kernel"), self.pg(name + "/conv7/bias"), red,
                                  self.pg(name + "/norm3/gamma"), self.pg(name + "/norm3/beta"),
                                  self.pg(name + "/norm4/gamma"), self.pg(name + "/norm4/beta"))
                ops.se_gate_bwd_apply(self.ctx, out.g, raw3.t, raw4.t, tr.st3, tr.st4, tr.g3, tr.b3, tr.g4, tr.b4,
                                      tr.gate, dr, red, dpool, raw3.g, raw4.g, None, None, None, None,
                                      accumulate=acc)
            self._timed("se_tail_bwd", 0, run, nbytes=(10 if acc else 8) * nel * es)   # 2 x 3 reads + 2 writes (+ 2)
            out.g = None
        if self.record:
            self._rec(bwd, [name + sfx for sfx in ("/norm3/gamma", "/norm3/beta", "/norm4/gamma", "/norm4/beta",
                                                   "/conv6/kernel", "/conv6/bias", "/conv7/kernel", "/conv7/bias")])
        return out

    # ---- K6 attention gate core ------------------------------------------------------------------
    def attn_core(self, theta, phi, x, name):
        f = theta.c
        wpsi = self.p(name + "/conv3/kernel", (1, 1, 1, f, 1), "kernel")
        bpsi = self.p(name + "/conv3/bias", (1,), "bias")
        y = Act(x.shape, x.dtype, self.new(x.shape, x.dtype))
        if self.tracing:
            return y
        psi = self.new((theta.shape[0],) + theta.grid, torch.float32)
        self.new_twin(y)
        nb_att = (theta.t.numel() + 2 * x.t.numel()) * x.t.element_size()
        self._timed("attn_fwd", 0, lambda: ops.attn_fwd(self.ctx, theta.t, phi.t, wpsi, bpsi, x.t, psi, y.t, y.tw),
                    nbytes=nb_att)

        def bwd():
            if y.g is None:
                return
            assert theta.g is None
            theta.g = self.new_grad(theta)
            dphi = self.new(phi.shape, torch.float32, zero=True)
            if x.needs_grad:
                gx, acc = self.grad_buffer(x)
            else:
                gx, acc = self.new_grad(x), False
            self._timed("attn_bwd", 0, nbytes=2 * nb_att + x.t.numel() * x.t.element_size(), fn=lambda: ops.attn_bwd(
                self.ctx, y.g, theta.t, phi.t, wpsi, psi, x.t, gx, acc, theta.g, dphi,
                self.pg(name + "/conv3/kernel"), self.pg(name + "/conv3/bias")))
            if phi.g is None:
                phi.g = self.new_grad(phi)
                ops.cast(self.ctx, dphi, phi.g)
            else:
                tmp = self.new_grad(phi)
                ops.cast(self.ctx, dphi, tmp)
                ops.axpy(self.ctx, tmp, 1.0, phi.g)
            y.g = None
        if self.record:
            self._rec(bwd, [name + "/conv3/kernel", name + "/conv3/bias"])
        return y

    # ---- K7 latent heads ----------------------------------------------------------------------------
    def latent(self, ml, mode, eps):
        """ml: fp32 [.., 2L]; mode 0 sample (eps fp32 tensor), 1 mean."""
        L = ml.c // 2
        zc = self.padc(L)
        z = Act(ml.shape[:-1] + (zc,), self.act_dtype, self.new(ml.shape[:-1] + (zc,)), lc=L)
        if self.tracing:
            return z
        ops.latent_fwd(self.ctx, ml.t, eps, mode, z.t)

        def bwd():
            if z.g is None:
                return
            gbuf, _ = self.grad_buffer(ml, zero=True)
            ops.latent_bwd(self.ctx, z.g, ml.t, eps, mode, gbuf, z.dtype)
            z.g = None
        if self.record:
            self._rec(bwd, [])
        return z

    def kl(self, ml_q, ml_p, kl_out):
        if self.tracing:
            return
        ops.kl_fwd(self.ctx, ml_q.t, ml_p.t, kl_out)

    def kl_seed_grad(self, ml_q, ml_p, scale):
        gq, _ = self.grad_buffer(ml_q, zero=True)
        gp, _ = self.grad_buffer(ml_p, zero=True)
        ops.kl_bwd(self.ctx, ml_q.t, ml_p.t, scale, gq, gp)
