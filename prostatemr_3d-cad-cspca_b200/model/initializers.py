"""Keras initializers used by M1 (R:networks.py:45-46, README.md:42-43), restated on numpy.
Instances are plain descriptors: ``__call__(shape, seed) -> np.ndarray`` runs once on the host."""
import math

import numpy as np


class Initializer:
    def get_config(self):
        return {k: v for k, v in self.__dict__.items()}


class Orthogonal(Initializer):
    """tf.keras.initializers.Orthogonal: QR of a N(0,1) matrix of shape (prod(shape[:-1]), shape[-1])."""

    def __init__(self, gain=1.0, seed=None):
        self.gain, self.seed = gain, seed

    def __call__(self, shape, seed):
        rows, cols = int(np.prod(shape[:-1])), int(shape[-1])
        rng = np.random.RandomState(seed if self.seed is None else self.seed)
        a = rng.standard_normal((max(rows, cols), min(rows, cols))).astype(np.float32)
        q, r = np.linalg.qr(a)
        q = q * np.sign(np.diag(r))
        if rows < cols:
            q = q.T
        return (self.gain * q).reshape(shape).astype(np.float32)


class TruncatedNormal(Initializer):
    """tf.keras.initializers.TruncatedNormal: samples beyond two standard deviations are redrawn."""

    def __init__(self, mean=0.0, stddev=0.05, seed=None):
        self.mean, self.stddev, self.seed = mean, stddev, seed

    def __call__(self, shape, seed):
        rng = np.random.RandomState(seed if self.seed is None else self.seed)
        x = rng.standard_normal(shape)
        bad = np.abs(x) > 2
        while bad.any():
            x[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(x) > 2
        return (x * self.stddev + self.mean).astype(np.float32)


class GlorotUniform(Initializer):
    """Keras default kernel initializer (SE conv6/conv7, R:network_blocks.py:45-46)."""

    def __init__(self, seed=None):
        self.seed = seed

    def __call__(self, shape, seed):
        rec = int(np.prod(shape[:-2]))
        lim = math.sqrt(6.0 / (rec * shape[-2] + rec * shape[-1]))
        rng = np.random.RandomState(seed if self.seed is None else self.seed)
        return rng.uniform(-lim, lim, shape).astype(np.float32)


class Constant(Initializer):
    def __init__(self, value=0.0):
        self.value = value

    def __call__(self, shape, seed):
        return np.full(shape, self.value, dtype=np.float32)


def Zeros():
    return Constant(0.0)


def Ones():
    return Constant(1.0)
