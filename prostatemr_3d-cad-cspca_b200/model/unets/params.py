"""Parameter table of an M1 model: one flat fp32 master buffer (+ gradient and Adam m / v / v-hat
buffers of the same layout) with per-parameter views in the KERAS layouts of the reference
(Conv3D kernel (kd,kh,kw,Cin,Cout), Conv3DTranspose kernel (kd,kh,kw,Cout,Cin), R:networks.py:472-565),
so that importing reference weights is a rename, not a transpose hunt.

Layout: [kernels | biases | unregularised (SE conv6/7, InstanceNorm gamma/beta)], every parameter
aligned to 64 floats. One group = one launch of the fused Adam kernel with that group's L2
coefficient; the gradient buffer is also the unit that data-parallel training all-reduces in
contiguous buckets."""
import zlib
from collections import OrderedDict

import numpy as np
import torch

ALIGN = 64
GROUPS = ("kernel", "bias", "plain")
KIND_GROUP = {"kernel": "kernel", "bias": "bias", "se_kernel": "plain", "se_bias": "plain",
              "gamma": "plain", "beta": "plain"}


class ParamSpec:
    __slots__ = ("name", "shape", "kind", "offset", "size")

    def __init__(self, name, shape, kind):
        self.name, self.shape, self.kind = name, tuple(int(s) for s in shape), kind
        self.size = int(np.prod(self.shape))
        self.offset = -1


class ParamTable:
    def __init__(self):
        self.specs = OrderedDict()
        self.finalized = False
        self.group_range = {}
        self.total = 0
        self.w = self.g = self.m = self.v = self.vhat = None
        self._views = {}
        self._gviews = {}

    # ---- registration (trace time) -----------------------------------------------------------
    def declare(self, name, shape, kind):
        sp = self.specs.get(name)
        if sp is None:
            assert not self.finalized, f"parameter {name} requested after the table was finalised"
            sp = ParamSpec(name, shape, kind)
            self.specs[name] = sp
        assert sp.shape == tuple(shape), (name, sp.shape, tuple(shape))
        return sp

    def finalize(self):
        off = 0
        for grp in GROUPS:
            start = off
            for sp in self.specs.values():
                if KIND_GROUP[sp.kind] == grp:
                    sp.offset = off
                    off += -(-sp.size // ALIGN) * ALIGN
            self.group_range[grp] = (start, off)
        self.total = off
        self.finalized = True

    def num_params(self, kinds=None):
        return sum(sp.size for sp in self.specs.values() if kinds is None or sp.kind in kinds)

    # ---- storage -----------------------------------------------------------------------------
    def allocate(self, device):
        assert self.finalized
        mk = lambda: torch.zeros(self.total, dtype=torch.float32, device=device)  # noqa: E731
        self.w, self.g, self.m, self.v, self.vhat = mk(), mk(), mk(), mk(), mk()
        self._views.clear()
        self._gviews.clear()

    def initialize(self, init_for_kind, seed):
        """Host-side initialisation (one-off): init_for_kind[kind](shape, seed) -> np.ndarray."""
        host = np.zeros(self.total, dtype=np.float32)
        for sp in self.specs.values():
            s = (zlib.crc32(sp.name.encode()) + 7919 * seed) % (2 ** 31 - 1)
            host[sp.offset:sp.offset + sp.size] = init_for_kind[sp.kind](sp.shape, s).reshape(-1)
        self.w.copy_(torch.from_numpy(host))

    def view(self, name):
        v = self._views.get(name)
        if v is None:
            sp = self.specs[name]
            v = self.w[sp.offset:sp.offset + sp.size].view(sp.shape)
            self._views[name] = v
        return v

    def grad(self, name):
        v = self._gviews.get(name)
        if v is None:
            sp = self.specs[name]
            v = self.g[sp.offset:sp.offset + sp.size].view(sp.shape)
            self._gviews[name] = v
        return v

    def state_dict(self):
        return {name: self.view(name).detach().cpu().numpy() for name in self.specs}

    def load_state_dict(self, weights, strict=True):
        missing = [n for n in self.specs if n not in weights]
        if strict and missing:
            raise KeyError(f"missing weights: {missing[:5]}{'...' if len(missing) > 5 else ''}")
        for name, sp in self.specs.items():
            if name in weights:
                t = torch.as_tensor(np.asarray(weights[name]), dtype=torch.float32).reshape(sp.shape)
                self.view(name).copy_(t)
