"""One thin Python wrapper per C entry point of include/m1b200.h.

Everything here takes torch CUDA tensors purely as device-memory handles (exported through
DLPack, see _lib.ptr) and launches on torch's current CUDA stream. No arithmetic happens in
Python or in PyTorch.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (BF16, CONV_FWD, CONV_TRANSPOSED, ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05, F32,
                   ConvDesc, Dropout, check, current_stream, dtype_code, lib, ptr, ptr_array)


def same_pads(size, k, s):
    """TF 'SAME' padding: out=ceil(in/s); pad_total=max((out-1)s+k-in,0); before=total//2."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return out, total // 2, total - total // 2


def conv_desc(mode, batch, in_dhw, out_dhw, kernel, stride, pad, src_c, out_c, w_strides,
              accumulate=False, act_dtype=F32, engine=ENGINE_AUTO):
    d = ConvDesc()
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
        d.w_stride_tap[j], d.w_stride_red[j], d.w_stride_out[j] = w_strides[j]
    d.accumulate = 1 if accumulate else 0
    d.act_dtype = act_dtype
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


def conv3d_pack_weights(ctx, d, ws):
    nbytes = lib().m1_conv3d_packed_bytes(C.byref(d))
    if nbytes == 0:
        raise _lib.M1Error("conv launch not supported by the tcgen05 engine")
    packed = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=ws[0].device)
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
