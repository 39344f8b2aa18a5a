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
GROUPS = ("kernel", "bias", "plain", "frozen")
KIND_GROUP = {"kernel": "kernel", "bias": "bias", "se_kernel": "plain", "se_bias": "plain",
              "gamma": "plain", "beta": "plain"}


class ParamSpec:
    """shape = the reference (Keras) shape; pshape = the shape held in HBM. They differ only for layers
    whose channel counts are zero-padded to the tensor-core granularity (16): `index` maps, per padded
    axis, logical position -> physical position. Padded entries are zero, receive exactly-zero
    gradients (their activations are identically zero) and therefore stay zero under Adam."""
    __slots__ = ("name", "shape", "kind", "offset", "size", "pshape", "psize", "index", "frozen")

    def __init__(self, name, shape, kind, pshape=None, index=None, frozen=False):
        self.name, self.shape, self.kind = name, tuple(int(s) for s in shape), kind
        # frozen: a variable the reference creates but that is not reachable from the model outputs (the prior
        # net's sersd0 + logits in probabilistic mode, Q3): Keras neither trains nor regularises it
        self.frozen = frozen
        self.size = int(np.prod(self.shape))
        self.pshape = tuple(int(s) for s in (pshape if pshape is not None else shape))
        self.psize = int(np.prod(self.pshape))
        self.index = index if (index and self.pshape != self.shape) else None
        self.offset = -1

    def to_physical(self, logical):
        """numpy logical array -> zero-padded physical array"""
        if self.index is None:
            return np.asarray(logical, dtype=np.float32).reshape(self.pshape)
        out = np.zeros(self.pshape, dtype=np.float32)
        ix = np.ix_(*[self.index.get(ax, np.arange(n)) for ax, n in enumerate(self.shape)])
        out[ix] = np.asarray(logical, dtype=np.float32).reshape(self.shape)
        return out

    def to_logical(self, physical):
        if self.index is None:
            return np.asarray(physical).reshape(self.shape)
        ix = np.ix_(*[self.index.get(ax, np.arange(n)) for ax, n in enumerate(self.shape)])
        return np.asarray(physical).reshape(self.pshape)[ix]


class ParamTable:
    def __init__(self):
        self.specs = OrderedDict()
        self.finalized = False
        self.group_range = {}
        self.total = 0
        self.w = self.g = self.m = self.v = self.vhat = None
        self.declare_frozen = False     # parameters declared while set are 'frozen' (see ParamSpec)
        self._views = {}
        self._gviews = {}

    # ---- registration (trace time) -----------------------------------------------------------
    def declare(self, name, shape, kind, pshape=None, index=None):
        sp = self.specs.get(name)
        if sp is None:
            assert not self.finalized, f"parameter {name} requested after the table was finalised"
            sp = ParamSpec(name, shape, kind, pshape, index, self.declare_frozen)
            self.specs[name] = sp
        assert sp.shape == tuple(shape), (name, sp.shape, tuple(shape))
        assert sp.pshape == tuple(pshape if pshape is not None else shape), (name, sp.pshape, pshape)
        return sp

    def finalize(self):
        off = 0
        for grp in GROUPS:
            start = off
            for sp in self.specs.values():
                if ("frozen" if sp.frozen else KIND_GROUP[sp.kind]) == grp:
                    sp.offset = off
                    off += -(-sp.psize // ALIGN) * ALIGN
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
            host[sp.offset:sp.offset + sp.psize] = sp.to_physical(init_for_kind[sp.kind](sp.shape, s)).reshape(-1)
        self.w.copy_(torch.from_numpy(host))

    def view(self, name):
        v = self._views.get(name)
        if v is None:
            sp = self.specs[name]
            v = self.w[sp.offset:sp.offset + sp.psize].view(sp.pshape)
            self._views[name] = v
        return v

    def grad(self, name):
        v = self._gviews.get(name)
        if v is None:
            sp = self.specs[name]
            v = self.g[sp.offset:sp.offset + sp.psize].view(sp.pshape)
            self._gviews[name] = v
        return v

    def state_dict(self):
        """reference-shaped (logical) numpy arrays"""
        return {name: np.ascontiguousarray(sp.to_logical(self.view(name).detach().cpu().numpy()))
                for name, sp in self.specs.items()}

    def grad_dict(self):
        return {name: torch.from_numpy(np.ascontiguousarray(sp.to_logical(self.grad(name).detach().cpu().numpy())))
                for name, sp in self.specs.items()}

    def load_state_dict(self, weights, strict=True):
        missing = [n for n in self.specs if n not in weights]
        if strict and missing:
            raise KeyError(f"missing weights: {missing[:5]}{'...' if len(missing) > 5 else ''}")
        for name, sp in self.specs.items():
            if name in weights:
                t = torch.from_numpy(sp.to_physical(np.asarray(weights[name], dtype=np.float32)))
                self.view(name).copy_(t)
