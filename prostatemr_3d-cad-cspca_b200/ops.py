"""One thin Python wrapper per C entry point of include/m1b200.h.

Everything here takes torch CUDA tensors purely as device-memory handles (exported through
DLPack, see _lib.ptr) and launches on torch's current CUDA stream. No arithmetic happens in
Python or in PyTorch.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (BF16, CONV_FWD, CONV_TRANSPOSED, ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05, F16, F32, PROBS,
                   ConvDesc, Dropout, check, current_stream, dtype_code, lib, ptr, ptr_array)


def same_pads(size, k, s):
    """TF 'SAME' padding: out=ceil(in/s); pad_total=max((out-1)s+k-in,0); before=total//2."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return out, total // 2, total - total // 2


def conv_desc(mode, batch, in_dhw, out_dhw, kernel, stride, pad, src_c, out_c, w_strides,
              accumulate=False, act_dtype=F32, engine=ENGINE_AUTO, out_dtype=None, w_by_src=False, w_dtype=0):
    d = ConvDesc()
    d.w_dtype = w_dtype
    d.w_by_src = 1 if w_by_src else 0
    d.mode = mode
    d.batch = batch
    for i in range(3):
        d.in_dhw[i] = in_dhw[i]
        d.out_dhw[i] = out_dhw[i]
        d.kernel[i] = kernel[i]
        d.stride[i] = stride[i]
        d.pad[i] = pad[i]
    d.nsrc = len(src_c)
    for i, c in enumerate(src_c):
        d.src_c[i] = c
    d.nout = len(out_c)
    for j, c in enumerate(out_c):
        d.out_c[j] = c
    for j, st in enumerate(w_strides):        # per produced tensor, or per gathered tensor when w_by_src
        d.w_stride_tap[j], d.w_stride_red[j], d.w_stride_out[j] = st
    if isinstance(accumulate, (list, tuple)):          # per-output flags -> bitmask
        d.accumulate = sum(1 << j for j, a in enumerate(accumulate) if a)
    else:
        d.accumulate = ((1 << len(out_c)) - 1) if accumulate else 0
    d.act_dtype = act_dtype
    d.out_dtype = act_dtype if out_dtype is None else out_dtype
    d.engine = engine
    return d


def conv3d(ctx, d, srcs, ws, biases, outs, w_packed=None):
    """m1_conv3d: srcs/outs lists of activation tensors, ws fp32 master weights (one per output)."""
    bias_arr = ptr_array([ptr(b) for b in biases]) if biases is not None else None
    w_arr = ptr_array([ptr(w) for w in ws]) if ws is not None else None
    check(lib().m1_conv3d(ctx.handle, C.byref(d), ptr_array([ptr(s) for s in srcs]), w_arr,
                          ptr(w_packed), bias_arr, ptr_array([ptr(o) for o in outs]),
                          current_stream()))


def conv3d_tc_supported(d):
    return bool(lib().m1_conv3d_tc_supported(C.byref(d)))


def conv3d_halo_engine(d):
    """True if m1_conv3d would run d on the halo variant of the tcgen05 engine (honours d.tune[0])."""
    return bool(lib().m1_conv3d_halo_engine(C.byref(d)))


def conv3d_plan_info(d, which):
    """Tiling of the tcgen05 engines for d (0 per-tap conv, 1 halo conv, 2 weight gradient) as a list of ints;
    [] if the engine does not take the launch. Pure host-side planning: works without a GPU."""
    out = (C.c_int32 * 16)()
    n = lib().m1_conv3d_plan_info(C.byref(d), which, out)
    return [int(out[i]) for i in range(n)]


def conv3d_wgrad_tc_supported(d):
    return bool(lib().m1_conv3d_wgrad_tc_supported0(C.byref(d)))


def conv3d_pack_weights(ctx, d, ws):
    nbytes = lib().m1_conv3d_packed_bytes(C.byref(d))
    if nbytes == 0:
        raise _lib.M1Error("conv launch not supported by the tcgen05 engine")
    wd = d.w_dtype or d.act_dtype
    packed = torch.empty(nbytes // 2, dtype=torch.float16 if wd == F16 else torch.bfloat16, device=ws[0].device)
    check(lib().m1_conv3d_pack_weights(ctx.handle, C.byref(d), ptr_array([ptr(w) for w in ws]),
                                       ptr(packed), current_stream()))
    return packed


def conv3d_pack_weights_into(ctx, d, ws, packed):
    check(lib().m1_conv3d_pack_weights(ctx.handle, C.byref(d), ptr_array([ptr(w) for w in ws]),
                                       ptr(packed), current_stream()))


def conv3d_wgrad(ctx, d, srcs, douts, dws, dbiases):
    db = ptr_array([ptr(b) for b in dbiases]) if dbiases is not None else None
    check(lib().m1_conv3d_wgrad(ctx.handle, C.byref(d), ptr_array([ptr(s) for s in srcs]),
                                ptr_array([ptr(g) for g in douts]),
                                ptr_array([ptr(g) for g in dws]), db, current_stream()))


# ---- K4 instance norm ----------------------------------------------------------------------
def _nvc(x):
    """(batch, voxels, C) of an NDHWC tensor."""
    return x.shape[0], x.numel() // (x.shape[0] * x.shape[-1]), x.shape[-1]


def inorm_stats(ctx, x, stats, eps=1e-3):
    n, v, c = _nvc(x)
    check(lib().m1_inorm_stats(ctx.handle, ptr(x), dtype_code(x), n, v, c, eps, ptr(stats), current_stream()))


def inorm_act_fwd(ctx, x, stats, gamma, beta, slope, y, y_bf16=None):
    """y_bf16: optional bf16 twin of the output (fp16 mode: operand of the tensor-core weight gradients)"""
    n, v, c = _nvc(x)
    check(lib().m1_inorm_act_fwd(ctx.handle, ptr(x), ptr(stats), ptr(gamma), ptr(beta), dtype_code(x), n, v, c,
                                 slope, ptr(y), ptr(y_bf16), current_stream()))


def inorm_act_bwd(ctx, dy, x, stats, gamma, beta, slope, dx, accumulate, dgamma, dbeta):
    n, v, c = _nvc(x)
    check(lib().m1_inorm_act_bwd(ctx.handle, ptr(dy), ptr(x), ptr(stats), ptr(gamma), ptr(beta), dtype_code(x),
                                 n, v, c, slope, ptr(dx), 1 if accumulate else 0, ptr(dgamma), ptr(dbeta),
                                 current_stream()))


# ---- K5 squeeze-excite -----------------------------------------------------------------------
def make_dropout(rate, u=None, seed=0, stream_id=0, step=None, mask=None):
    """step: device uint64 tensor (1 element) holding the training-step counter of a replayed graph;
    mask: uint8 tensor of numel/8 bytes - keep-bits written by se_gate_fwd and read back by its backward."""
    d = Dropout()
    d.u = ptr(u) if u is not None else None
    d.seed = seed
    d.stream_id = stream_id
    d.rate = rate
    d.step = ptr(step) if step is not None else None
    d.mask = ptr(mask) if mask is not None else None
    return d


def se_squeeze(ctx, raw3, stats3, gamma3, beta3, pool):
    n, v, c = _nvc(raw3)
    check(lib().m1_se_squeeze(ctx.handle, ptr(raw3), ptr(stats3), ptr(gamma3), ptr(beta3), dtype_code(raw3),
                              n, v, c, ptr(pool), current_stream()))


def se_excite_fwd(ctx, pool, w6, b6, w7, b7, hidden, gate, stats3=None, gamma3=None, beta3=None):
    """stats3 / gamma3 / beta3 given: the squeeze (pool from the statistics of raw3) is folded into this launch"""
    n, c = pool.shape
    cr = hidden.shape[-1]
    check(lib().m1_se_excite_fwd(ctx.handle, ptr(pool), ptr(w6), ptr(b6), ptr(w7), ptr(b7), n, c, cr,
                                 ptr(hidden), ptr(gate), ptr(stats3), ptr(gamma3), ptr(beta3), current_stream()))


def se_excite_bwd(ctx, dgate, pool, hidden, gate, w6, w7, dpool, dw6, db6, dw7, db7, red5=None, dg3=None, db3=None,
                  dg4=None, db4=None):
    """red5 + the four norm parameter gradients given: they are accumulated here (se_gate_bwd_apply then gets None)"""
    n, c = pool.shape
    cr = hidden.shape[-1]
    check(lib().m1_se_excite_bwd(ctx.handle, ptr(dgate), ptr(pool), ptr(hidden), ptr(gate), ptr(w6), ptr(w7),
                                 n, c, cr, ptr(dpool), ptr(dw6), ptr(db6), ptr(dw7), ptr(db7), ptr(red5), ptr(dg3),
                                 ptr(db3), ptr(dg4), ptr(db4), current_stream()))


def se_gate_fwd(ctx, raw3, raw4, st3, st4, g3, b3, g4, b4, gate, drop, out, out_bf16=None):
    n, v, c = _nvc(raw3)
    check(lib().m1_se_gate_fwd(ctx.handle, ptr(raw3), ptr(raw4), ptr(st3), ptr(st4), ptr(g3), ptr(b3), ptr(g4),
                               ptr(b4), ptr(gate), C.byref(drop), dtype_code(raw3), n, v, c, ptr(out), ptr(out_bf16),
                               current_stream()))


def se_gate_bwd_reduce(ctx, dout, raw3, raw4, st3, st4, g3, b3, g4, b4, gate, drop, red, dgate):
    n, v, c = _nvc(raw3)
    check(lib().m1_se_gate_bwd_reduce(ctx.handle, ptr(dout), ptr(raw3), ptr(raw4), ptr(st3), ptr(st4), ptr(g3),
                                      ptr(b3), ptr(g4), ptr(b4), ptr(gate), C.byref(drop), dtype_code(raw3),
                                      n, v, c, ptr(red), ptr(dgate), current_stream()))


def se_gate_bwd_apply(ctx, dout, raw3, raw4, st3, st4, g3, b3, g4, b4, gate, drop, red, dpool, draw3, draw4,
                      dg3, db3, dg4, db4, accumulate=False):
    n, v, c = _nvc(raw3)
    check(lib().m1_se_gate_bwd_apply(ctx.handle, ptr(dout), ptr(raw3), ptr(raw4), ptr(st3), ptr(st4), ptr(g3),
                                     ptr(b3), ptr(g4), ptr(b4), ptr(gate), C.byref(drop), ptr(red), ptr(dpool),
                                     dtype_code(raw3), n, v, c, ptr(draw3), ptr(draw4), 1 if accumulate else 0, ptr(dg3),
                                     ptr(db3), ptr(dg4), ptr(db4), current_stream()))


# ---- K6 attention gate -----------------------------------------------------------------------
def _grid(t):
    return (C.c_int32 * 3)(*t.shape[1:4])


def attn_fwd(ctx, theta, phi, w_psi, b_psi, x, psi, y, y_bf16=None):
    check(lib().m1_attn_fwd(ctx.handle, ptr(theta), ptr(phi), ptr(w_psi), ptr(b_psi), ptr(x), dtype_code(x),
                            x.shape[0], _grid(theta), _grid(phi), _grid(x), theta.shape[-1], x.shape[-1],
                            ptr(psi), ptr(y), ptr(y_bf16), current_stream()))


def attn_bwd(ctx, dy, theta, phi, w_psi, psi, x, dx, acc_dx, dtheta, dphi, dw_psi, db_psi):
    check(lib().m1_attn_bwd(ctx.handle, ptr(dy), ptr(theta), ptr(phi), ptr(w_psi), ptr(psi), ptr(x),
                            dtype_code(x), x.shape[0], _grid(theta), _grid(phi), _grid(x), theta.shape[-1],
                            x.shape[-1], ptr(dx), 1 if acc_dx else 0, ptr(dtheta), ptr(dphi), ptr(dw_psi),
                            ptr(db_psi), current_stream()))


# ---- K7 latent heads / KL --------------------------------------------------------------------
def latent_fwd(ctx, ml, eps, mode, z):
    n, v, c2 = _nvc(ml)
    check(lib().m1_latent_fwd(ctx.handle, ptr(ml), ptr(eps), mode, n, v, c2 // 2, dtype_code(z), z.shape[-1],
                              ptr(z), current_stream()))


def latent_bwd(ctx, dz, ml, eps, mode, dml, z_dtype=None):
    """z_dtype: torch dtype of the latent VALUE z (dz is stored as grad_dtype(z_dtype)); None: that of dz"""
    n, v, c2 = _nvc(ml)
    code = dtype_code(dz) if z_dtype is None else _lib.code_of(z_dtype)
    check(lib().m1_latent_bwd(ctx.handle, ptr(dz), ptr(ml), ptr(eps), mode, n, v, c2 // 2, code,
                              dz.shape[-1], ptr(dml), current_stream()))


def kl_fwd(ctx, ml_q, ml_p, kl_out):
    n, v, c2 = _nvc(ml_q)
    check(lib().m1_kl_fwd(ctx.handle, ptr(ml_q), ptr(ml_p), n, v, c2 // 2, ptr(kl_out), current_stream()))


def kl_bwd(ctx, ml_q, ml_p, scale, dml_q, dml_p):
    n, v, c2 = _nvc(ml_q)
    check(lib().m1_kl_bwd(ctx.handle, ptr(ml_q), ptr(ml_p), n, v, c2 // 2, scale, ptr(dml_q), ptr(dml_p),
                          current_stream()))


# ---- K8 softmax + focal ----------------------------------------------------------------------
def softmax_focal(ctx, logits, y_true, alpha, gamma, up, prob, head_off, head_weight, loss_out, dlogits,
                  grad_scale, from_probs=False):
    nc = logits.shape[-1]
    al = (C.c_float * nc)(*[float(a) for a in alpha]) if alpha is not None else None
    check(lib().m1_softmax_focal(
        ctx.handle, ptr(logits), PROBS if from_probs else F32, ptr(y_true), dtype_code(y_true) if y_true is not None else F32,
        C.cast(al, C.c_void_p) if al is not None else None, gamma, logits.shape[0], _grid(logits),
        (C.c_int32 * 3)(*up), nc, ptr(prob), prob.shape[-1] if prob is not None else 0, head_off, head_weight,
        ptr(loss_out), ptr(dlogits), grad_scale, current_stream()))


# ---- K9 optimizer + utilities ----------------------------------------------------------------
def adam_amsgrad(ctx, w, g, m, v, vhat, lr_t, b1, b2, eps, l2, gscale, l2_out=None, amsgrad=True):
    check(lib().m1_adam_amsgrad(ctx.handle, ptr(w), ptr(g), ptr(m), ptr(v), ptr(vhat), w.numel(), lr_t, b1, b2,
                                eps, l2, gscale, ptr(l2_out), 1 if amsgrad else 0, current_stream()))


def adam_amsgrad_dev(ctx, w, g, m, v, vhat, lr_t_dev, b1, b2, eps, l2, gscale, l2_out=None, amsgrad=True):
    check(lib().m1_adam_amsgrad_dev(ctx.handle, ptr(w), ptr(g), ptr(m), ptr(v), ptr(vhat), w.numel(), ptr(lr_t_dev),
                                    b1, b2, eps, l2, gscale, ptr(l2_out), 1 if amsgrad else 0, current_stream()))


def cast(ctx, src, dst):
    check(lib().m1_cast(ctx.handle, ptr(src), dtype_code(src), ptr(dst), dtype_code(dst), src.numel(),
                        current_stream()))


def copy_channels(ctx, src, src_off, dst, dst_off, c):
    rows = src.numel() // src.shape[-1]
    check(lib().m1_copy_channels(ctx.handle, ptr(src), dtype_code(src), src.shape[-1], src_off, ptr(dst),
                                 dtype_code(dst), dst.shape[-1], dst_off, c, rows, current_stream()))


def axpy(ctx, x, a, y):
    check(lib().m1_axpy(ctx.handle, ptr(x), dtype_code(x), a, ptr(y), x.numel(), current_stream()))


def decision_fusion(ctx, prior, follow, strategy, out):
    check(lib().m1_decision_fusion(ctx.handle, ptr(prior), ptr(follow), strategy, follow.numel(), ptr(out),
                                   current_stream()))


def bias_grad(ctx, dout, dbias):
    rows = dout.numel() // dout.shape[-1]
    check(lib().m1_bias_grad(ctx.handle, ptr(dout), dtype_code(dout), rows, dout.shape[-1], ptr(dbias),
                             current_stream()))


def philox_normal(ctx, seed, stream_id, out, step=None):
    if step is None:
        check(lib().m1_philox_normal(ctx.handle, seed, stream_id, ptr(out), out.numel(), current_stream()))
    else:
        check(lib().m1_philox_normal_step(ctx.handle, seed, stream_id, ptr(step), ptr(out), out.numel(),
                                          current_stream()))


def logits_softmax_focal(ctx, feat, w, bias, y_true, alpha, gamma, prob, head_off, head_weight, loss_out, dfeat,
                         acc_dfeat, dw, db, grad_scale):
    """Returns False when (C, nc) has no fused instantiation (caller falls back to conv3d + softmax_focal)."""
    n, v, c = _nvc(feat)
    nc = w.shape[-1]
    al = (C.c_float * nc)(*[float(a) for a in alpha]) if alpha is not None else None
    rc = lib().m1_logits_softmax_focal(
        ctx.handle, ptr(feat), dtype_code(feat), ptr(w), ptr(bias), ptr(y_true),
        dtype_code(y_true) if y_true is not None else F32, C.cast(al, C.c_void_p) if al is not None else None,
        gamma, n, v, c, nc, ptr(prob), prob.shape[-1] if prob is not None else 0, head_off, head_weight,
        ptr(loss_out), ptr(dfeat), 1 if acc_dfeat else 0, ptr(dw), ptr(db), grad_scale, current_stream())
    if rc == 2:
        return False
    check(rc)
    return True


# ---- cascade ---------------------------------------------------------------------------------------
def logits_prob_bwd(ctx, feat, w, prob, head_off, dprob, dfeat, acc_dfeat, dw, db):
    """backward of logits conv + softmax for a gradient w.r.t. the probabilities (m1_logits_prob_bwd)"""
    n, v, c = _nvc(feat)
    nc = w.shape[-1]
    rc = lib().m1_logits_prob_bwd(ctx.handle, ptr(feat), dtype_code(feat), ptr(w), ptr(prob), prob.shape[-1], head_off,
                                  ptr(dprob), n * v, c, nc, ptr(dfeat), 1 if acc_dfeat else 0, ptr(dw), ptr(db),
                                  current_stream())
    if rc == 2:
        return False
    check(rc)
    return True


def fusion_focal(ctx, prob1, ch1, prob2, ch2, strategy, y_true, alpha, gamma, det1, det2, weight, loss_out, dp1, dp2,
                 grad_scale):
    """decision fusion + focal loss of the joint prediction + d/dp1, d/dp2 (m1_fusion_focal)"""
    n = prob1.shape[0]
    v = prob1.numel() // (n * prob1.shape[-1])
    al = (C.c_float * 2)(*[float(a) for a in alpha]) if alpha is not None else None
    check(lib().m1_fusion_focal(ctx.handle, ptr(prob1), prob1.shape[-1], ch1, ptr(prob2), prob2.shape[-1], ch2, strategy,
                                ptr(y_true), dtype_code(y_true) if y_true is not None else F32,
                                C.cast(al, C.c_void_p) if al is not None else None, gamma, n, v, ptr(det1), ptr(det2),
                                weight, ptr(loss_out), ptr(dp1), ptr(dp2), grad_scale, current_stream()))


class PackPlan:
    """m1_pack_plan: every tensor-core operand pack of a model re-derived from the fp32 master weights in ONE launch.
    jobs: [(ConvDesc, [fp32 weight views], packed tensor)]. Create outside CUDA-graph capture; run() is capturable."""

    def __init__(self, ctx, jobs):
        n = len(jobs)
        self.ctx, self.n = ctx, n
        self._keep = jobs                                                  # the device buffers must outlive the plan
        desc_ptrs = (C.c_void_p * n)(*[C.addressof(d) for d, _, _ in jobs])
        ws_arrays = [ptr_array([ptr(w) for w in ws]) for _, ws, _ in jobs]
        ws_ptrs = (C.c_void_p * n)(*[C.cast(a, C.c_void_p).value for a in ws_arrays])
        packed = (C.c_void_p * n)(*[ptr(pk) for _, _, pk in jobs])
        h = C.c_void_p()
        check(lib().m1_pack_plan_create(ctx.handle, n, C.cast(desc_ptrs, C.POINTER(ConvDesc)),
                                        C.cast(ws_ptrs, C.POINTER(C.c_void_p)), C.cast(packed, C.POINTER(C.c_void_p)),
                                        C.byref(h)))
        self.handle = h

    def run(self):
        check(lib().m1_pack_plan_run(self.ctx.handle, self.handle, current_stream()))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                lib().m1_pack_plan_destroy(self.handle)
        except Exception:      # noqa: BLE001 - interpreter shutdown
            pass


# ---- K10 augmentations ---------------------------------------------------------------------------
def augment(ctx, op, x, out, plans_dev, eps=None):
    """m1_augment: one transform (_lib.AUG_*) over the batch x (B, D, H, W, C) fp32 -> out; plans_dev: uint8 device
    tensor holding `B` packed m1_aug_plan structs; eps: (B, D, H, W, 3) fp32 for AUG_NOISE."""
    b, d, h, w, c = x.shape
    check(lib().m1_augment(ctx.handle, int(op), ptr(x), ptr(out), ptr(eps), ptr(plans_dev), b, d, h, w, c,
                           current_stream()))
