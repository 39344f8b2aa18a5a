"""ctypes binding of libm1b200.so (include/m1b200.h) and DLPack device-pointer export.

There is deliberately no fallback: if the shared library is missing the import of any compute
entry point raises, and every call fails loudly on a machine without an sm_100 GPU.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libm1b200.so")

M1_MAX_SRC = 8
M1_MAX_OUT = 8
F32, BF16, F16 = 0, 1, 2
PROBS = 16      # m1_softmax_focal: the fp32 input already holds probabilities
CONV_FWD, CONV_TRANSPOSED = 0, 1
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1, 2


class ConvDesc(C.Structure):
    """m1_conv_desc of include/m1b200.h"""
    _fields_ = [
        ("mode", C.c_int32), ("batch", C.c_int32),
        ("in_dhw", C.c_int32 * 3), ("out_dhw", C.c_int32 * 3),
        ("kernel", C.c_int32 * 3), ("stride", C.c_int32 * 3), ("pad", C.c_int32 * 3),
        ("nsrc", C.c_int32), ("src_c", C.c_int32 * M1_MAX_SRC),
        ("nout", C.c_int32), ("out_c", C.c_int32 * M1_MAX_OUT),
        ("w_stride_tap", C.c_int64 * M1_MAX_OUT), ("w_stride_red", C.c_int64 * M1_MAX_OUT),
        ("w_stride_out", C.c_int64 * M1_MAX_OUT),
        ("w_by_src", C.c_int32), ("accumulate", C.c_int32), ("act_dtype", C.c_int32), ("out_dtype", C.c_int32),
        ("engine", C.c_int32), ("w_dtype", C.c_int32), ("tune", C.c_int32 * 4),
    ]


class Dropout(C.Structure):
    """m1_dropout of include/m1b200.h"""
    _fields_ = [("u", C.c_void_p), ("seed", C.c_uint64), ("stream_id", C.c_uint64),
                ("rate", C.c_float), ("step", C.c_void_p), ("mask", C.c_void_p)]


class AugPlan(C.Structure):
    """m1_aug_plan of include/m1b200.h"""
    _fields_ = [("zoom_on", C.c_int32), ("zoom_scale", C.c_int32), ("flip_on", C.c_int32),
                ("rot_on", C.c_int32), ("rot_pad", C.c_int32), ("rot_crop_h", C.c_int32), ("rot_crop_w", C.c_int32),
                ("rot_cos", C.c_float), ("rot_sin", C.c_float), ("rot_xoff", C.c_float), ("rot_yoff", C.c_float),
                ("tr_on", C.c_int32), ("tr_top", C.c_int32), ("tr_bottom", C.c_int32), ("tr_right", C.c_int32),
                ("tr_left", C.c_int32),
                ("cs_on", C.c_int32), ("cs_channel", C.c_int32), ("cs_top", C.c_int32), ("cs_bottom", C.c_int32),
                ("cs_right", C.c_int32), ("cs_left", C.c_int32),
                ("gamma_on", C.c_int32 * 3), ("gamma", C.c_float), ("poor_on", C.c_int32 * 3),
                ("noise_on", C.c_int32), ("noise_std", C.c_float)]


AUG_ZOOM, AUG_FLIP, AUG_ROTATE, AUG_TRANSLATE, AUG_CHANNEL_SHIFT, AUG_GAMMA, AUG_POOR_SCAN, AUG_NOISE = range(8)


class M1Error(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes), PARSED from include/m1b200.h so the binding cannot drift from the ABI
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "m1b200.h")
_PP = C.POINTER(C.c_void_p)


def _ctype_of(decl):
    d = " ".join(decl.replace("const", " ").split())
    if "*" in d:
        base = d.split("*")[0].strip()
        stars = d.count("*")
        if base == "m1_conv_desc":
            return C.POINTER(ConvDesc)
        if base == "m1_dropout":
            return C.POINTER(Dropout)
        if stars >= 2:
            return _PP
        if base == "char":
            return C.c_char_p
        return C.c_void_p
    base = d.split()[0] if d.split() else "void"
    return {"int": C.c_int, "int32_t": C.c_int32, "int64_t": C.c_int64, "float": C.c_float,
            "uint64_t": C.c_uint64, "void": None}[base]


def parse_header(path=HEADER_PATH):
    import re
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = text[text.index("typedef struct m1_ctx m1_ctx;"):]
    sigs = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(m1_\w+)\s*\(([^;{}]*?)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef"):
            continue
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                # drop the parameter name (last identifier) unless the declarator ends with '*'
                a = re.sub(r"\b\w+$", "", a) if not a.endswith("*") else a
                argtypes.append(_ctype_of(a))
        sigs[name] = (_ctype_of(ret), argtypes)
    return sigs


SIGNATURES = parse_header()


def lib():
    """Loads libm1b200.so (once). Raises if it has not been built - there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise M1Error(
                f"{LIB_PATH} not found: build it with prostatemr_3d-cad-cspca_b200/csrc/build.sh "
                "(or __graft_entry__.build()); m1b200 has no CPU / PyTorch fallback")
        handle = C.CDLL(LIB_PATH)
        missing = [name for name in SIGNATURES if not hasattr(handle, name)]
        if missing and os.environ.get("M1_BRINGUP") == "1":   # kernel bring-up only
            for name in missing:
                SIGNATURES.pop(name)
            missing = []
        if missing:
            raise M1Error(f"{LIB_PATH} lacks symbols declared in include/m1b200.h: {missing}")
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status):
    if status != 0:
        raise M1Error(lib().m1_last_error().decode())


# ---- DLPack zero-copy pointer export -------------------------------------------------------
class _DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int32), ("device_id", C.c_int32)]


class _DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class _DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", _DLDevice), ("ndim", C.c_int32),
                ("dtype", _DLDataType), ("shape", C.POINTER(C.c_int64)),
                ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


C.pythonapi.PyCapsule_GetPointer.restype = C.c_void_p
C.pythonapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
_KDL_CUDA = 2


def dlpack_device_ptr(obj, expect_cuda=True):
    """Device pointer of any object exporting ``__dlpack__`` (torch / cupy / jax / TF-experimental
    tensors), without copying: reads DLManagedTensor.dl_tensor.{data,byte_offset}."""
    capsule = obj.__dlpack__()
    raw = C.pythonapi.PyCapsule_GetPointer(capsule, b"dltensor")
    dl = C.cast(raw, C.POINTER(_DLTensor)).contents
    if expect_cuda and dl.device.device_type != _KDL_CUDA:
        raise M1Error("m1b200 kernels need CUDA device memory (DLPack device_type "
                      f"{dl.device.device_type}); there is no CPU path")
    ptr = (dl.data or 0) + dl.byte_offset
    # the capsule still owns the DLManagedTensor; dropping it calls the deleter (no copy was made)
    del capsule
    return ptr


_STRICT_DLPACK = os.environ.get("M1_PTR_VIA_DLPACK") == "1"


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL). Foreign tensors (anything exporting __dlpack__:
    cupy / jax / TF-experimental) always go through the DLPack capsule; for torch tensors the capsule's
    data pointer IS tensor.data_ptr() (asserted by tests/test_host_logic.py), so the per-launch fast path
    reads that directly - M1_PTR_VIA_DLPACK=1 forces the capsule route everywhere."""
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        if not t.is_cuda:
            raise M1Error("m1b200 kernels need CUDA device memory; there is no CPU path")
        if _STRICT_DLPACK:
            assert t.is_contiguous(), "m1b200 kernels take dense tensors"
            return dlpack_device_ptr(t.detach(), expect_cuda=True) if t.numel() else None
        return t.data_ptr() or None
    return dlpack_device_ptr(t)


def code_of(dtype):
    """m1_dtype code of a torch dtype"""
    if dtype == torch.float32:
        return F32
    if dtype == torch.bfloat16:
        return BF16
    if dtype == torch.float16:
        return F16
    raise M1Error(f"unsupported activation dtype {dtype}")


def dtype_code(t):
    return code_of(t.dtype)


TORCH_DTYPE = {F32: torch.float32, BF16: torch.bfloat16, F16: torch.float16}


def grad_dtype(dtype):
    """storage type of the GRADIENT of an activation stored as `dtype` (M1_GRAD_DTYPE of include/m1b200.h):
    fp16 values have bf16 gradients - fp16 has too little range for gradients."""
    return torch.bfloat16 if dtype == torch.float16 else dtype


def ptr_array(ptrs):
    arr = (C.c_void_p * len(ptrs))(*[p if p else None for p in ptrs])
    return C.cast(arr, C.POINTER(C.c_void_p))


def current_stream():
    return torch.cuda.current_stream().cuda_stream


class Context:
    """m1_ctx wrapper: one per GPU / process."""
    _instances = {}

    def __init__(self, device):
        self.device = device
        h = C.c_void_p()
        check(lib().m1_ctx_create(int(device), C.byref(h)))
        self.handle = h

    @classmethod
    def get(cls, device=None):
        if device is None:
            device = torch.cuda.current_device()
        if device not in cls._instances:
            cls._instances[device] = cls(device)
        return cls._instances[device]

    def launch_count(self, reset=False):
        return lib().m1_ctx_launch_count(self.handle, 1 if reset else 0)
