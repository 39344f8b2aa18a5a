"""Mirror of tf2.5/scripts/model/losses.py:20-63 (Focal, EvidenceLowerBound).

The objects are descriptors: M1.compile reads alpha/gamma/beta and the fused K8 kernel
(m1_softmax_focal) evaluates loss and gradient on the logits. ``Focal.loss`` / ``.FL`` can also be
called directly on device tensors (softmax predictions), like the reference methods."""
import torch

from .. import ops
from .._lib import Context


class Focal:
    """[1] T.Y. Lin et al. (2017). Requires 'y_pred': softmax prediction, 'y_true': one-hot label.
    Defaults as losses.py:27."""

    def __init__(self, alpha=[0.25, 0.75], gamma=2.00):  # noqa: B006 - reference signature
        self.alpha = alpha
        self.gamma = gamma

    def FL(self, y_true, y_pred):
        """losses.py:32-39 on CUDA tensors; y_pred are probabilities (one head)."""
        return _focal_on_probabilities(self, y_true, y_pred)

    def loss(self, y_true, y_pred):
        """losses.py:43-49: mean of FL over the C_pred // C_true heads."""
        nc = y_true.shape[-1]
        heads = y_pred.shape[-1] // nc
        vals = [self.FL(y_true, y_pred[..., nc * i:nc * (i + 1)].contiguous()) for i in range(heads)]
        return sum(vals) / len(vals)


class EvidenceLowerBound:
    """losses.py:52-63: beta * sum(y_pred); the KL itself is computed inside the model."""

    def __init__(self, beta=1.00):
        self.beta = beta

    def loss(self, y_true, y_pred):
        return self.beta * y_pred.sum()


def _focal_on_probabilities(focal, y_true, y_pred):
    assert y_pred.is_cuda, "m1b200 has no CPU path"
    ctx = Context.get(y_pred.device.index)
    probs = y_pred.contiguous()
    if probs.dtype != torch.float32:
        p32 = torch.empty(probs.shape, dtype=torch.float32, device=probs.device)
        ops.cast(ctx, probs, p32)
        probs = p32
    loss = torch.zeros(1, dtype=torch.float32, device=y_pred.device)
    ops.softmax_focal(ctx, probs, y_true.contiguous(), focal.alpha, float(focal.gamma), (1, 1, 1), None, 0, 1.0,
                      loss, None, 0.0, from_probs=True)
    return loss[0]
