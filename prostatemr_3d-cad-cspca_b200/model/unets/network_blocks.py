"""Mirror of tf2.5/scripts/model/unets/network_blocks.py: the feature-extraction and probabilistic
blocks of M1, expressed as launches of the fused sm_100a kernels through the Engine.

  SEResNetBottleNeck     R:network_blocks.py:23-80
  GridAttentionBlock3D   R:network_blocks.py:88-130
  MonteCarloDropout      R:network_blocks.py:137-143   (fused into the SE gate kernel, K5)
  StitchingProbDecoder   R:network_blocks.py:244-278
"""
from .engine import LRELU


class SEResNetBottleNeck:
    """conv1||conv4 run as ONE implicit GEMM (same input, same kernel/stride: N = f/4 + f); norm3,
    norm4, the squeeze-excite gate, the multiplicative "residual addition" (Q5), LeakyReLU(0.1) and
    the dropout that always follows the block are one fused kernel (Engine.se_tail)."""

    def __init__(self, filters, kernel_size, strides, reduction, name):
        self.filters, self.kernel_size, self.strides = filters, tuple(kernel_size), tuple(strides)
        self.reduction, self.name = reduction, name

    def __call__(self, eng, srcs, drop=None):
        return eng.se_gate(self.trunk(eng, srcs), drop)

    def trunk(self, eng, srcs):
        """the block up to (not including) the gate kernel: everything that does not depend on the dropout draw"""
        f, n = self.filters, self.name
        cin = sum(a.lc for a in srcs)
        if cin == f:
            # R:network_blocks.py:63 - identity residual; never reached by M1 (every block changes the
            # channel count) and it would need the concatenation materialised.
            raise NotImplementedError("SEResNetBottleNeck with an identity residual (Cin == filters)")
        # the f/4 bottleneck tensors are zero-padded to 16 channels when f/4 is not a multiple of 16 (f = 32)
        raw1, raw4 = eng.conv(srcs, [(n + "/conv1", f // 4), (n + "/conv4", f)], self.kernel_size, self.strides,
                              pad_out=[True, False], feeds_norm=True)
        a = eng.inorm_act(raw1, n + "/norm1", LRELU)
        raw2, = eng.conv([a], [(n + "/conv2", f // 4)], (3, 3, 3), pad_out=[True], feeds_norm=True)
        b = eng.inorm_act(raw2, n + "/norm2", LRELU)
        raw3, = eng.conv([b], [(n + "/conv3", f)], (1, 1, 1), feeds_norm=True)
        return eng.se_trunk(raw3, raw4, n, self.reduction)


class GridAttentionBlock3D:
    """theta = conv1_{k=s=sub_samp}(x); phi = conv2_{111}(g); psi = sigmoid(conv3(lrelu(theta + up(phi))));
    y = up(psi) * x; W_y = norm4(conv4_{111}(y)). The nearest up-samplings are index arithmetic inside
    the fused kernel (Engine.attn_core)."""

    def __init__(self, inter_channels, sub_samp, name):
        self.inter_channels, self.sub_samp, self.name = inter_channels, tuple(sub_samp), name

    def __call__(self, eng, conv_tensor, gating_tensor):
        n, f = self.name, self.inter_channels
        theta, = eng.conv([conv_tensor], [(n + "/conv1", f)], self.sub_samp, self.sub_samp)
        phi, = eng.conv([gating_tensor], [(n + "/conv2", f)], (1, 1, 1))
        y = eng.attn_core(theta, phi, conv_tensor, n)
        raw, = eng.conv([y], [(n + "/conv4", f)], (1, 1, 1), feeds_norm=True)
        return eng.inorm_act(raw, n + "/norm4", 1.0)


class MonteCarloDropout:
    """tf.nn.dropout active in training AND inference (R:network_blocks.py:137-143). A descriptor:
    the mask is applied inside the preceding block's fused gate kernel."""

    def __init__(self, rate, always_on=True):
        self.rate, self.always_on = rate, always_on

    def active(self, training):
        return self.rate > 0.0 and (self.always_on or training)


class StitchingProbDecoder:
    """1x1x1 logits on the probabilistic decoder features (R:network_blocks.py:275-278)."""

    def __init__(self, num_classes, name="final_decoder"):
        self.num_classes, self.name = num_classes, name

    def __call__(self, eng, decoder_features):
        from .engine import LazyHead
        return LazyHead(eng, decoder_features, self.name + "/logits", self.num_classes)
