"""Mirror of tf2.5/scripts/model/unets/networks.py: the M1 (Hierarchical Probabilistic) 3D U-Net.

  M1        top-level model, constructor kwargs / compile / fit / get_detect_model   R:networks.py:24-223
  m1 wiring deterministic and probabilistic branches, 4 live passes, KL, softmax     R:networks.py:232-392
  M1Core    stem, SE-ResNet encoder, attention gates, nested decoder, latent decoder R:networks.py:402-782

Everything here is host logic: it decides WHICH sm_100a kernels run on WHICH buffers (through
Engine) - all arithmetic happens in libm1b200.so. Quirks Q1-Q9 of the reference are
reproduced or shimmed exactly as documented next to each.
"""
import time

import numpy as np
import torch

from ... import ops
from .. import initializers, regularizers
from ..losses import EvidenceLowerBound, Focal
from ..optimizers import Adam
from .engine import LRELU, Engine, InjectedNoise, LazyHead, PhiloxNoise
from .modelio import LoadableModel, store_config_args
from .network_blocks import GridAttentionBlock3D, SEResNetBottleNeck, StitchingProbDecoder
from .params import ParamTable


# ---------------------------------------------------------------------------------------------
# M1Core (R:networks.py:402-782)
# ---------------------------------------------------------------------------------------------
class M1Core:
    def __init__(self, net,
                 num_classes=2, dropout_mode='standard', dropout_rate=0.50,
                 filters=(32, 64, 128, 256, 512),
                 strides=((1, 1, 1), (1, 2, 2), (1, 2, 2), (2, 2, 2), (1, 2, 2)),
                 kernel_sizes=((1, 3, 3), (1, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3)),
                 se_reduction=(8, 8, 8, 8, 8),
                 att_sub_samp=((1, 1, 1), (1, 1, 1), (1, 1, 1), (1, 1, 1)),
                 dense_skip=False, deep_supervision=False, probabilistic=False,
                 prob_latent_dims=(1, 1, 1, 1)):
        self.net = net
        self.num_classes, self.dropout_mode, self.dropout_rate = num_classes, dropout_mode, dropout_rate
        self.filters, self.strides, self.kernel_sizes = filters, strides, kernel_sizes
        self.se_reduction, self.att_sub_samp = se_reduction, att_sub_samp
        self.dense_skip, self.deep_supervision = dense_skip, deep_supervision
        self.probabilistic, self.prob_latent_dims = probabilistic, prob_latent_dims
        # R:networks.py:465-469
        assert len(self.filters) == 5, "ERROR: Expected Tuple/Array with 5 Values (One Per Resolution)."
        assert len(self.se_reduction) == 5, "ERROR: Expected Tuple/Array with 5 Values (One Per Resolution)."
        assert [len(a) for a in self.att_sub_samp] == [3, 3, 3, 3], \
            "ERROR: Expected 4x3 Tuple/Array (3D Sub-Sampling Factors for 4 Attention Gates)."
        assert [len(s) for s in self.strides] == [3, 3, 3, 3, 3], \
            "ERROR: Expected 5x3 Tuple/Array (3D Strides for 5 Resolutions)."
        assert [len(k) for k in self.kernel_sizes] == [3, 3, 3, 3, 3], \
            "ERROR: Expected 5x3 Tuple/Array (3D Kernels for 5 Resolutions)."
        assert dropout_mode in ('standard', 'monte-carlo')
        F, S, K, R = filters, strides, kernel_sizes, se_reduction
        n = lambda s: net + '/' + s  # noqa: E731
        one = (1, 1, 1)
        self.serse = [None] + [SEResNetBottleNeck(F[i], K[i], S[i], R[i], n('serse%d' % i)) for i in range(1, 5)]
        self.att = [GridAttentionBlock3D(F[i], att_sub_samp[i], n('att%d' % i)) for i in range(4)]
        self.sersd = [SEResNetBottleNeck(F[i], K[i], one, R[i], n('sersd%d' % i)) for i in range(4)]
        Fr, Kr, Rr = F[::-1], K[::-1], R[::-1]
        # sersp{3-i}: filters F[::-1][i+1], kernel K[::-1][i+1], reduction R[::-1][i+1]  (R:networks.py:554-561)
        self.sersp = {3 - i: SEResNetBottleNeck(Fr[i + 1], Kr[i + 1], one, Rr[i + 1], n('sersp%d' % (3 - i)))
                      for i in range(4)}
        self.shapes = {}

    def __call__(self, eng, inputs, prob_mean=False, prob_z_q=None, pass_name='det', training=True,
                 stop='full', need_logits=True):
        """inputs: list of Acts forming the (virtual) input concatenation.
        stop='latents' is the Keras-pruned partial pass: only what feeds the last latent head."""
        F, S, K = self.filters, self.strides, self.kernel_sizes
        net = self.net
        n = lambda s: net + '/' + s  # noqa: E731
        L = self.prob_latent_dims
        prob = self.probabilistic
        partial = stop == 'latents'
        last_lat = max([i for i in range(len(L)) if L[i] != 0], default=-1) if prob else -1
        drop_on = training or self.dropout_mode == 'monte-carlo'
        rate = self.dropout_rate

        def drop(site, r=rate):
            return (pass_name, site, r) if (drop_on and r > 0.0) else None

        need = lambda lvl: (not partial) or (3 - lvl) < last_lat  # noqa: E731 - is uconv{lvl}_ needed?
        out = {}
        # stem (R:networks.py:574-576)
        # The stem and the first encoder block up to its dropout see nothing stochastic: two passes of one network
        # over the SAME input tensors (q_sample / q_mean, p_z_q / p_z_qmean) compute them once and differ from the
        # gate kernel of serse1 on (its dropout mask); their gradients meet again in raw3.g / raw4.g / x.g.
        key = (net, 'stem+serse1') + tuple(id(a) for a in inputs)
        hit = eng.shared.get(key) if eng.share_trunk else None
        if hit is None:
            raw, = eng.conv(inputs, [(n('conve0'), F[0])], K[0], S[0], feeds_norm=True)
            x = eng.inorm_act(raw, n('norme0'), LRELU)
            trunk1 = self.serse[1].trunk(eng, [x])
            eng.shared[key] = (x, trunk1, inputs)
        else:
            x, trunk1, _ = hit
        # encoder (R:networks.py:579-582); the dropouts are fused into the SE gate kernel
        conv1 = eng.se_gate(trunk1, drop('drope1'))
        conv2 = self.serse[2](eng, [conv1], drop('drope2'))
        conv3 = self.serse[3](eng, [conv2], drop('drope3'))
        convm = self.serse[4](eng, [conv3], drop('drope4'))
        enc = [x, conv1, conv2, conv3]
        # attention gates (R:networks.py:585-588): the gating signal is always convm
        att = [self.att[i](eng, enc[i], convm) if need(i) else None for i in range(4)]

        dense = self.dense_skip
        uconv, uconv_ = [None] * 4, [None] * 4
        ct = lambda name, src, f, k, s: eng.conv([src], [(n(name), f)], k, s, transposed=True)[0]  # noqa: E731
        if need(3):       # decoder stage 3 (R:networks.py:591-597)
            deconv3 = ct('convtd3', convm, F[3], K[4], S[4])
            if dense and need(2):
                deconv3_up1 = ct('convtd3_up1', deconv3, F[2], K[3], S[3])
                if need(1):
                    deconv3_up2 = ct('convtd3_up2', deconv3_up1, F[1], K[2], S[2])
                    if need(0):
                        deconv3_up3 = ct('convtd3_up3', deconv3_up2, F[0], K[1], S[1])
            uconv_[3] = [deconv3, att[3]]
        if need(2):       # stage 2 (R:networks.py:600-607)
            uconv[3] = self.sersd[3](eng, uconv_[3], drop('dropd3'))
            deconv2 = ct('convtd2', uconv[3], F[2], K[3], S[3])
            if dense:
                if need(1):
                    deconv2_up1 = ct('convtd2_up1', deconv2, F[1], K[2], S[2])
                    if need(0):
                        deconv2_up2 = ct('convtd2_up2', deconv2_up1, F[0], K[1], S[1])
                uconv_[2] = [deconv2, deconv3_up1, att[2]]
            else:
                uconv_[2] = [deconv2, att[2]]
        if need(1):       # stage 1 (R:networks.py:610-616)
            uconv[2] = self.sersd[2](eng, uconv_[2], drop('dropd2'))
            deconv1 = ct('convtd1', uconv[2], F[1], K[2], S[2])
            if dense:
                if need(0):
                    deconv1_up1 = ct('convtd1_up1', deconv1, F[0], K[1], S[1])
                uconv_[1] = [deconv1, deconv2_up1, deconv3_up2, att[1]]
            else:
                uconv_[1] = [deconv1, att[1]]
        if need(0):       # stage 0 (R:networks.py:619-624)
            uconv[1] = self.sersd[1](eng, uconv_[1], drop('dropd1'))
            deconv0 = ct('convtd0', uconv[1], F[0], K[1], S[1])
            if dense:
                uconv_[0] = [deconv0, deconv1_up1, deconv2_up2, deconv3_up3, att[0]]
            else:
                uconv_[0] = [deconv0, att[0]]
        if not partial and need_logits:
            uconv[0] = self.sersd[0](eng, uconv_[0], drop('dropd0', rate / 2))      # R:networks.py:523
            out['logits'] = LazyHead(eng, uconv[0], n('logits'), self.num_classes)
        self.shapes = dict(inputs=[a.shape for a in inputs], x=x.shape, conv1=conv1.shape, conv2=conv2.shape,
                           conv3=conv3.shape, convm=convm.shape,
                           att=[None if a is None else a.shape for a in att],
                           uconv_=[None if u is None else u[0].shape[:-1] + (sum(a.lc for a in u),) for u in uconv_],
                           uconv=[None if u is None else u.shape for u in uconv])

        ds_feats = {}
        if prob:          # hierarchical latent decoder (R:networks.py:633-734)
            dists, used = [], []
            feat = convm
            Fr, Kr, Sr = F[::-1], K[::-1], S[::-1]
            for i in range(4):
                lvl = 3 - i
                if partial and i > last_lat:
                    break
                if L[i] != 0:
                    ml, = eng.conv([feat], [(n('mu_logsig%d' % lvl), 2 * L[i])], (1, 1, 1), out_dtype=torch.float32)
                    if prob_z_q is not None:
                        z = prob_z_q[len(used)]
                    elif prob_mean:
                        z = eng.latent(ml, 1, None)
                    else:
                        eps = None if eng.tracing else eng.noise.normal(eng, pass_name, 'eps%d' % lvl,
                                                                        ml.shape[:-1] + (L[i],))
                        z = eng.latent(ml, 0, eps)
                    dists.append(ml)
                    used.append(z)
                    if partial and i == last_lat:
                        break
                    hi_in = [z, feat]
                else:
                    hi_in = [feat]
                up, = eng.conv(hi_in, [(n('dec_hi%d' % lvl), Fr[i + 1])], Kr[i], Sr[i], transposed=True)
                feat = self.sersp[lvl](eng, [up] + uconv_[lvl], drop('dropp%d' % lvl))
                ds_feats[lvl] = feat
            out['prob_distributions'] = dists
            out['prob_used_latents'] = used
            if not partial:
                out['prob_decoder_features'] = feat

        if self.deep_supervision and not partial:
            # R:networks.py:737-747. 1x1x1 conv and nearest up-sampling commute exactly, so the DS logits
            # are computed at low resolution and up-sampled inside the loss kernel.
            srcs = [uconv[1], uconv[2], uconv[3]] if not prob else [ds_feats[1], ds_feats[2], ds_feats[3]]
            s1, s2, s3 = (np.array(S[i]) for i in (1, 2, 3))
            ups = [tuple(int(v) for v in s1), tuple(int(v) for v in s1 * s2), tuple(int(v) for v in s1 * s2 * s3)]
            out['ds_logits'] = []
            for j, (t, u) in enumerate(zip(srcs, ups)):
                lg, = eng.conv([t], [(n('dsy%d_logits' % (j + 1)), self.num_classes)], (1, 1, 1),
                               out_dtype=torch.float32)
                out['ds_logits'].append((lg, u))
        return out

    def summary(self):
        """R:networks.py:761-782 (the shape listing; Q2)."""
        s = self.shapes
        rows = [('Input Volume:', s['inputs'][0][:-1] + (sum(i[-1] for i in s['inputs']),)),
                ('Initial Convolutional Layer (Stage 0):', s['x']),
                ('Attention Gating: Stage 0:', s['att'][0]), ('Encoder: Stage 1; SE-Residual Block:', s['conv1']),
                ('Attention Gating: Stage 1:', s['att'][1]), ('Encoder: Stage 2; SE-Residual Block:', s['conv2']),
                ('Attention Gating: Stage 2:', s['att'][2]), ('Encoder: Stage 3; SE-Residual Block:', s['conv3']),
                ('Attention Gating: Stage 3:', s['att'][3]), ('Middle: High-Dim Latent Features:', s['convm']),
                ('Decoder: Stage 3; Nested U-Net Concat.:', s['uconv_'][3]),
                ('Decoder: Stage 3; Nested U-Net End:', s['uconv'][3]),
                ('Decoder: Stage 2; Nested U-Net Concat.:', s['uconv_'][2]),
                ('Decoder: Stage 2; Nested U-Net End:', s['uconv'][2]),
                ('Decoder: Stage 1; Nested U-Net Concat.:', s['uconv_'][1]),
                ('Decoder: Stage 1; Nested U-Net End:', s['uconv'][1]),
                ('Decoder: Stage 0; Nested U-Net Concat.:', s['uconv_'][0]),
                ('Decoder: Stage 0; Nested U-Net End:', s['uconv'][0])]
        for label, shp in rows:
            print((label + '-' * 60)[:60], None if shp is None else (None,) + tuple(shp[1:]))


# ---------------------------------------------------------------------------------------------
# one stage (R:networks.py:232-392) = parameters + engine + the pass wiring
# ---------------------------------------------------------------------------------------------
FUSION = {'identity': 0, True: 0, 'noisy-or': 1, 'bayes': 2}


class M1(LoadableModel):
    """
    [1] Z. Zhou et al. (2019), UNet++. [2] J. Hu et al.(2019), Squeeze-and-Excitation Networks.
    [3] S. Kohl et al. (2019), Hierarchical Probabilistic U-Net. [4] O. Oktay et al. (2018), Attention U-Net.

    Constructor arguments up to `name` are the reference's (R:networks.py:34-55), same names, order,
    defaults and assertions. The keyword-only arguments after it are additions of this implementation:
      precision   'fp16' (default: tcgen05 tensor-core convolutions on fp16 activations and weights, bf16
                  activation gradients, fp32 accumulation - meets the 2e-2 softmax / 1e-3 loss parity bounds)
                  | 'bf16' (everything the tensor cores see is bf16: same speed, 8-bit mantissa, misses the bounds)
                  | 'fp32' (CUDA-core fp32 convolutions: the 1e-4 parity mode)
      ds_in_prob  'reference': probabilistic + deep_supervision == no deep supervision (Q3, 2 output channels)
                  'intended' : wires the dead ds_ops branch (R:networks.py:743-747), 4 heads
      seed        parameter-initialisation and Philox seed;   device: CUDA device (None: current)
      build       False -> host-only object (configuration + parameter inventory, no GPU needed)
    """

    def __new__(cls, *args, **kwargs):
        """M1(..., cascaded='identity' | 'noisy-or' | 'bayes' | True) is the two-stage model of
        R:networks.py:109-193: the object returned is a CascadedM1 (same API, two single-stage M1 inside)."""
        if cls is M1:
            cascaded = kwargs.get('cascaded', args[14] if len(args) > 14 else False)
            if cascaded is not False:
                from .cascade import CascadedM1
                return object.__new__(CascadedM1)
        return object.__new__(cls)

    @store_config_args
    def __init__(self,
                 input_spatial_dims,
                 input_channels,
                 num_classes,
                 dropout_rate=0.50,
                 dropout_mode='standard',
                 filters=(32, 64, 128, 256, 512),
                 strides=((1, 1, 1), (1, 2, 2), (1, 2, 2), (2, 2, 2), (1, 2, 2)),
                 kernel_sizes=((1, 3, 3), (1, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3)),
                 se_reduction=(8, 8, 8, 8, 8),
                 att_sub_samp=((1, 1, 1), (1, 1, 1), (1, 1, 1)),
                 kernel_initializer=initializers.Orthogonal(gain=1.0),
                 bias_initializer=initializers.TruncatedNormal(mean=0.0, stddev=0.001),
                 kernel_regularizer=regularizers.l2(1e-4),
                 bias_regularizer=regularizers.l2(1e-4),
                 cascaded=False,
                 dense_skip=False,
                 deep_supervision=False,
                 probabilistic=False,
                 prob_latent_dims=(3, 2, 1),
                 summary=True,
                 name='UNET-TYPE-M1',
                 *, precision='fp16', ds_in_prob='reference', seed=0, device=None, build=None,
                 use_tcgen05=True, compute_dead_branches=False):
        ndims = len(input_spatial_dims)
        assert ndims in [1, 2, 3], 'Variable (ndims) should be  1, 2 or 3. Found: %d.' % ndims
        assert ndims == 3, 'the sm_100a kernels implement the 3-D model (the only one M1Core can build)'
        assert ds_in_prob in ('reference', 'intended')
        assert precision in ('fp16', 'bf16', 'fp32'), "precision must be 'fp16', 'bf16' or 'fp32'"
        assert cascaded is False, "cascaded models are built by M1.__new__ as CascadedM1 (model/unets/cascade.py)"
        self.name = name
        self.input_spatial_dims = tuple(int(d) for d in input_spatial_dims)
        self.input_channels, self.num_classes = int(input_channels), int(num_classes)
        self.probabilistic, self.deep_supervision = bool(probabilistic), bool(deep_supervision)
        self.ds_in_prob, self.precision = ds_in_prob, precision
        self.compute_dead_branches = compute_dead_branches
        self.use_tcgen05 = use_tcgen05
        self.l2_kernel = regularizers.coefficient(kernel_regularizer)
        self.l2_bias = regularizers.coefficient(bias_regularizer)
        core_kw = dict(num_classes=num_classes, dropout_mode=dropout_mode, dropout_rate=dropout_rate,
                       filters=tuple(filters), strides=tuple(tuple(s) for s in strides),
                       kernel_sizes=tuple(tuple(k) for k in kernel_sizes), se_reduction=tuple(se_reduction),
                       att_sub_samp=tuple(tuple(a) for a in att_sub_samp), dense_skip=dense_skip)
        if not probabilistic:
            # Q1: the reference calls M1Core(...)(inputs=inputs) without prob_mean/prob_z_q (TypeError on
            # HEAD); the intended semantics prob_mean=False, prob_z_q=None are implemented.
            self.core = M1Core('m1', deep_supervision=deep_supervision, probabilistic=False, **core_kw)
            self.heads = 4 if deep_supervision else 1
        else:
            # Q3: deep_supervision is NOT forwarded to the prior/posterior cores by the reference
            ds = bool(deep_supervision) and ds_in_prob == 'intended'
            kw = dict(core_kw, deep_supervision=ds, probabilistic=True, prob_latent_dims=tuple(prob_latent_dims))
            self.prior = M1Core('prior', **kw)
            self.posterior = M1Core('posterior', **dict(kw, deep_supervision=False))
            self.final_decoder = StitchingProbDecoder(num_classes)
            self.heads = 4 if ds else 1

        # ---- parameter inventory: a shape-only trace of one training step (Keras builds on first call)
        # Live parameters first; then the variables the reference creates but that no model output depends on (the
        # prior net's sersd0 + logits in probabilistic mode, whose output only feeds an empty slice, Q3): they are
        # declared 'frozen' - kept in the checkpoint inventory, never updated, never regularised - exactly what
        # Keras does with layers that are not on the path between the functional model's inputs and outputs.
        self.params = ParamTable()
        tracer = Engine(self.params, precision, device=None, use_tcgen05=use_tcgen05)
        self._graph(tracer, batch=1, training=True, trace=True, dead=self._dead_branches_live())
        if probabilistic:
            self._infer_graph(tracer, batch=1, trace=True)
            if not self._dead_branches_live():
                self.params.declare_frozen = True
                self._graph(Engine(self.params, precision, device=None, use_tcgen05=use_tcgen05), batch=1,
                            training=True, trace=True, dead=True)
                self.params.declare_frozen = False
        self.params.finalize()
        self.train_flops_per_volume = None

        # ---- references (R:networks.py:93-106)
        self.references = LoadableModel.ReferenceContainer()
        self.references.cascaded = cascaded
        self.references.probabilistic = probabilistic
        self.references.num_classes = num_classes
        self.references.m1_model = self
        if summary:
            self.summary()

        self.optimizer = None
        self.focal, self.elbo, self.loss_weights = None, None, [1.0, 1.0]
        self.world_size, self.rank, self.grad_sync = 1, 0, None
        self.eng = None
        self._init_kw = dict(kernel=kernel_initializer, bias=bias_initializer, seed=seed)
        self.noise = None
        self.history = None
        if build is None:
            build = torch.cuda.is_available()
        if build:
            self.build(device)

    # ---- construction ---------------------------------------------------------------------------
    def build(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("M1.build: no CUDA device - m1b200 has no CPU / PyTorch fallback")
        if device is None:
            device = 'cuda:%d' % torch.cuda.current_device()
        self.device = torch.device(device)
        self.params.allocate(self.device)
        kinds = dict(kernel=self._init_kw['kernel'], bias=self._init_kw['bias'],
                     se_kernel=initializers.GlorotUniform(), se_bias=initializers.Zeros(),
                     gamma=initializers.Ones(), beta=initializers.Zeros())
        self.params.initialize(kinds, self._init_kw['seed'])
        self.eng = Engine(self.params, self.precision, device=self.device, use_tcgen05=self.use_tcgen05)
        self.noise = PhiloxNoise(seed=42 + self._init_kw['seed'], rank=self.rank)
        return self

    def summary(self):
        print('-' * 85)
        if self.probabilistic:
            print('Hierarchical Prob. 3D U-Net (Type: M1) - Prior Network')
            print('-' * 85)
            self.prior.summary()
            print('-' * 85)
            print('Hierarchical Prob. 3D U-Net (Type: M1) - Posterior Network (live layers)')
        else:
            print('Deterministic 3D U-Net (Type: M1)')
            print('-' * 85)
            self.core.summary()
        print('-' * 85)
        print('Parameters: %d (conv %d, SE excite %d, InstanceNorm %d)' % (
            self.params.num_params(), self.params.num_params(('kernel', 'bias')),
            self.params.num_params(('se_kernel', 'se_bias')), self.params.num_params(('gamma', 'beta'))))
        print('-' * 85)

    # ---- graph wiring (R:networks.py:266-390) ------------------------------------------------------
    def _inputs(self, eng, batch, x, trace, needs_grad=False):
        """Model input -> activation tensors. Probabilistic (Q4, R:networks.py:300-301):
        image = inputs[..., :-(nc-1)], label = inputs[..., -(nc-1)-1:-1] - for nc=2, C=4 the 'label' is
        image channel 2, NOT channel 3; the slicing is reproduced exactly."""
        D = self.input_spatial_dims
        C, nc = self.input_channels, self.num_classes

        def padded(lc, copies):
            """(B,D,H,W,pad16(lc)) activation, zero beyond the real channels (tensor-core granularity)"""
            pc = eng.padc(lc)
            if trace:
                return eng.input((batch,) + D + (pc,), needs_grad=needs_grad, lc=lc)
            t = eng.new((batch,) + D + (pc,), zero=pc != lc)
            for src_off, dst_off, n in copies:
                ops.copy_channels(eng.ctx, x, src_off, t, dst_off, n)
            return eng.input(t, needs_grad=needs_grad, lc=lc)

        if not self.probabilistic:
            return [padded(C, [(0, 0, C)])], None
        ci = C - (nc - 1)
        lab_lo, lab_hi = C - (nc - 1) - 1, C - 1
        cl = lab_hi - lab_lo
        img = padded(ci, [(0, 0, ci)])
        # posterior input = concat([image, label]) as ONE tensor (R:networks.py:348)
        post = padded(ci + cl, [(0, 0, ci), (lab_lo, ci, cl)])
        return [img], [post]

    def _dead_branches_live(self):
        """True if the prior net's sersd0 + logits are computed (and then trained): on request, or when the
        intended deep-supervision wiring makes the prior core produce outputs of its own."""
        return bool(self.compute_dead_branches or (self.probabilistic and self.prior.deep_supervision))

    def _graph(self, eng, batch, training, trace=False, x=None, dead=None, input_needs_grad=False):
        """One training-graph forward. Returns dict(heads=[(logits Act, up)], kl_pairs=[(ml_q, ml_p)], inputs=[Acts]).
        input_needs_grad: also differentiate w.r.t. the model input (second stage of a cascade)."""
        img, lab = self._inputs(eng, batch, x, trace, input_needs_grad)
        if not self.probabilistic:
            o = self.core(eng, img, pass_name='det', training=training)
            heads = [(o['logits'], (1, 1, 1))] + (o.get('ds_logits') or [])
            return dict(heads=heads, kl_pairs=[], inputs=list(img))
        post_in = lab                 # [image || label] already concatenated by _inputs
        q_sample = self.posterior(eng, post_in, False, None, 'q_sample', training, 'latents')
        q_mean = self.posterior(eng, post_in, True, None, 'q_mean', training, 'latents')
        p_zq = self.prior(eng, img, False, q_sample['prob_used_latents'], 'p_z_q', training, 'latents')
        if dead is None:
            dead = self._dead_branches_live()
        p_zqm = self.prior(eng, img, False, q_mean['prob_used_latents'], 'p_z_qmean', training, 'full',
                           need_logits=dead)
        train_conv = self.final_decoder(eng, p_zqm['prob_decoder_features'])
        heads = [(train_conv, (1, 1, 1))] + (p_zqm.get('ds_logits') or [])
        return dict(heads=heads, inputs=list(img) + list(post_in),
                    kl_pairs=list(zip(q_sample['prob_distributions'], p_zq['prob_distributions'])))

    def _infer_graph(self, eng, batch, trace=False, x=None, pass_name='p_sample', inputs=None):
        """get_detect_model() graph (R:networks.py:196-206, :350,:355). inputs: activation list of an earlier call
        (passes of a Monte-Carlo ensemble over ONE batch share the stem + serse1 trunk through Engine.shared)."""
        img = inputs if inputs is not None else self._inputs(eng, batch, x, trace)[0]
        if not self.probabilistic:
            o = self.core(eng, img, pass_name='det', training=False)
            return o['logits']
        p_s = self.prior(eng, img, False, None, pass_name, False, 'full', need_logits=False)
        return self.final_decoder(eng, p_s['prob_decoder_features'])

    # ---- Keras-style API ---------------------------------------------------------------------------
    def compile(self, optimizer=None, loss=None, loss_weights=None, **_):
        """unet_model.compile(optimizer, loss=[Focal.loss, ELBO.loss], loss_weights=[1, 10])
        (train_model.py:231, README.md:60-61). `loss` entries are the bound .loss methods (or the objects)."""
        prev = self.optimizer
        self.optimizer = optimizer if optimizer is not None else Adam(1e-3, amsgrad=True)
        if not isinstance(self.optimizer, Adam):
            raise NotImplementedError("M1.compile: only model.optimizers.Adam (amsgrad True or False) has a fused "
                                      "update kernel, got %r" % type(self.optimizer).__name__)
        if prev is not None and prev is not self.optimizer and self.optimizer.iterations == 0:
            # a checkpoint restored m / v / v-hat and the step: a fresh optimizer object continues from there
            self.optimizer.iterations = prev.iterations
        if isinstance(loss, dict):
            loss = [loss.get('detection'), loss.get('KL')]
        losses = list(loss) if isinstance(loss, (list, tuple)) else [loss]
        objs = [getattr(ls, '__self__', ls) for ls in losses if ls is not None]
        for o in objs:
            if not isinstance(o, (Focal, EvidenceLowerBound)):
                raise NotImplementedError(
                    "M1.compile: loss %r has no fused kernel - the detection output is trained with losses.Focal "
                    "(losses.py:20-49) and the KL output with losses.EvidenceLowerBound (losses.py:52-63); "
                    "SoftDicePlusBoundarySurface (losses.py:66-128) and custom callables are not implemented"
                    % (getattr(o, '__name__', None) or type(o).__name__))
        focal = [o for o in objs if isinstance(o, Focal)]
        if objs and not focal:
            raise ValueError("M1.compile: no losses.Focal among the supplied losses - the detection output needs one")
        self.focal = focal[0] if focal else Focal()                 # only loss=None falls back to the defaults
        self.elbo = next((o for o in objs if isinstance(o, EvidenceLowerBound)), EvidenceLowerBound())
        if len(self.focal.alpha) != self.num_classes:      # train_model.py:148-149
            raise Exception("Number of Class Weights Declared in Loss Function != Number of Classes in "
                            "Labels/Loss Objective")
        lw = list(loss_weights) if loss_weights is not None else [1.0] * max(1, len(losses))
        self.loss_weights = [float(lw[0]), float(lw[1]) if len(lw) > 1 else 1.0]
        return self

    def distribute(self, bucket_bytes=32 << 20, group=None):
        """Data-parallel training over the initialised torch.distributed group (one process per GPU):
        the MirroredStrategy scope of train_model.py:167-170. Replicas must start from identical
        weights (same `seed`); each draws its own Philox sub-stream."""
        import torch.distributed as dist
        from ..distribute import BucketedGradSync
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.grad_sync = BucketedGradSync(self.params, bucket_bytes, group)
        if isinstance(self.noise, PhiloxNoise):
            self.noise = PhiloxNoise(seed=42 + self._init_kw['seed'], rank=self.rank)
        return self

    def set_noise(self, tensors=None, seed=None):
        """noise injection for parity runs: {(pass_name, site): tensor}; None -> Philox."""
        if tensors is not None:
            self.noise = InjectedNoise(tensors)
        else:
            self.noise = PhiloxNoise(seed if seed is not None else 42, self.rank)

    def _to_device(self, a, dtype=torch.float32):
        if isinstance(a, dict):
            a = a.get('image', a.get('detection'))
        t = torch.as_tensor(a) if not isinstance(a, torch.Tensor) else a
        return t.to(device=self.device, dtype=dtype, non_blocking=True).contiguous()

    # ---- CUDA-graph replay of the training step ------------------------------------------------------
    # One M1 step is ~1 700 kernel launches; issued from Python they cost more host time than the kernels
    # take on a B200. After GRAPH_WARMUP eager steps (tcgen05 autotuning, weight packs, function attributes)
    # the whole step - forward, losses, backward, gradient all-reduce, Adam, weight re-pack - is captured
    # once into a CUDA graph over static input buffers and replayed. Everything that changes per step lives
    # in device memory: the Philox step counter (m1_dropout.step / m1_philox_normal_step) and the Adam step
    # size (m1_adam_amsgrad_dev), both written by a tiny fill launch before the replay. No collective is ever
    # captured: data-parallel runs all-reduce the flat gradient buffer between the two graphs of a step.
    GRAPH_WARMUP = 2

    def _graph_ok(self, apply_update):
        import os
        if not apply_update or self.eng.prof is not None or not isinstance(self.noise, PhiloxNoise):
            return False
        mode = os.environ.get("M1_CUDA_GRAPH", "1")
        if mode == "0" or getattr(self, "_graph_failed", False):
            return False
        if self.world_size > 1 and os.environ.get("M1_CUDA_GRAPH_DP", "1") == "0":
            return False
        return True

    def _train_step_graphed(self, x, y):
        st = getattr(self, "_gs", None)
        if st is not None and (tuple(x.shape) != tuple(st['x'].shape) or tuple(y.shape) != tuple(st['y'].shape)):
            st = self._gs = None                      # new input shape: capture again
            self._graph_eager_steps = 0
        if st is None:
            n = getattr(self, "_graph_eager_steps", 0)
            if n < self.GRAPH_WARMUP:
                self._graph_eager_steps = n + 1
                return None
            st = dict(x=torch.empty_like(x), y=torch.empty_like(y),
                      step=torch.zeros(1, dtype=torch.int64, device=self.device),
                      lr=torch.zeros(1, dtype=torch.float32, device=self.device))
            st['x'].copy_(x); st['y'].copy_(y)
            st['step'].fill_(self.noise.step); st['lr'].fill_(self.optimizer.lr_t())
            torch.cuda.synchronize(self.device)
            # two graphs: [forward, losses, backward] and [Adam, weight re-pack]. In data-parallel runs the
            # gradient all-reduce is issued between the two replays, OUTSIDE any capture (one NCCL call on
            # the flat gradient buffer: 256 MB, ~1 ms over NVLink, against ~85 ms of kernels).
            g_bwd, g_upd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            self.noise.step_dev, self._lr_dev = st['step'], st['lr']
            before = self.eng.launch_total()
            import os
            # Data parallel, default: ONE flat all-reduce of the gradient buffer between the two graphs (measured at
            # 2 GPUs this round: 75.6 vs 74.6 ms per step = 98.7 % weak-scaling efficiency; 0.984 at 8 GPUs in round 1).
            # M1_CUDA_GRAPH_DP=overlap captures the BUCKETED all-reduce inside the backward graph instead - every
            # bucket's NCCL call issued as soon as the last kernel contributing to it has been enqueued, as in the eager
            # path. That mode HUNG at 2 GPUs on this round's box (torch 2.11 / NCCL 2.28.9, side-stream weight
            # gradients in the same capture) and is therefore opt-in until it has been debugged on hardware.
            st['dp_in_graph'] = (self.world_size > 1 and self.grad_sync is not None
                                 and os.environ.get("M1_CUDA_GRAPH_DP", "flat") == "overlap")
            try:
                try:
                    self._graph_dp_overlap = st['dp_in_graph']
                    with torch.cuda.graph(g_bwd):
                        st['out'] = self._train_step_eager(st['x'], st['y'], True, graphed=True)
                except Exception as e:                      # noqa: BLE001
                    if not st['dp_in_graph']:
                        raise
                    print("m1b200: capturing the bucketed all-reduce in the CUDA graph failed (%s: %s) - one flat "
                          "all-reduce between the graphs instead" % (type(e).__name__, str(e).splitlines()[0][:200]))
                    torch.cuda.synchronize(self.device)
                    st['dp_in_graph'] = self._graph_dp_overlap = False
                    g_bwd = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g_bwd):
                        st['out'] = self._train_step_eager(st['x'], st['y'], True, graphed=True)
                with torch.cuda.graph(g_upd, pool=g_bwd.pool()):
                    self._apply_update(st['out']['l2'], 1.0 / self.world_size, graphed=True)
            except Exception:
                self._graph_failed = True
                raise
            finally:
                self.noise.step_dev, self._lr_dev = None, None
                self._graph_dp_overlap = False
            st['graph'], st['graph_update'] = g_bwd, g_upd
            st['launches'] = self.eng.launch_total() - before
            self._gs = st
        else:
            st['x'].copy_(x, non_blocking=True)
            st['y'].copy_(y, non_blocking=True)
        st['step'].fill_(self.noise.step)
        st['lr'].fill_(self.optimizer.lr_t())
        st['graph'].replay()
        if self.world_size > 1 and not st.get('dp_in_graph'):
            import torch.distributed as dist
            dist.all_reduce(self.params.g, op=dist.ReduceOp.SUM,
                            group=self.grad_sync.group if self.grad_sync is not None else None)
        st['graph_update'].replay()
        self.graph_replays = getattr(self, "graph_replays", 0) + 1
        self.optimizer.iterations += 1
        self.noise.step += 1
        return st['out']

    @property
    def launches_per_graph_step(self):
        """kernels of libm1b200 inside one replayed step (counted at capture), or None in eager mode"""
        st = getattr(self, "_gs", None)
        return st['launches'] if st else None

    def train_step(self, x, y_true, apply_update=True):
        """Forward of the 4-pass graph, focal + KL losses, full backward, (all-reduce), Adam-AMSGrad.
        x: (B,D,H,W,C) fp32, y_true: one-hot (B,D,H,W,nc). Returns dict of device scalars (in graph mode:
        views of the graph's static output buffers, overwritten by the next step)."""
        if self.eng is None:
            raise RuntimeError("M1.train_step: model not built on a GPU (m1b200 has no CPU fallback)")
        if self.optimizer is None:
            self.compile()
        x = self._to_device(x)
        y = self._to_device(y_true)
        if self._graph_ok(apply_update):
            out = self._train_step_graphed(x, y)
            if out is not None:
                return out
        return self._train_step_eager(x, y, apply_update)

    def _forward_train(self, x, input_needs_grad=False):
        """forward of the training graph: (graph dict, detection tensor [B,D,H,W,nc*heads] fp32, loss scalars)"""
        eng = self.eng
        B = x.shape[0]
        eng.noise = self.noise
        eng.begin(record=True)
        self.params.g.zero_()
        g = self._graph(eng, B, training=True, x=x, input_needs_grad=input_needs_grad)
        det = torch.empty((B,) + self.input_spatial_dims + (self.num_classes * len(g['heads']),), dtype=torch.float32,
                          device=self.device)
        scal = torch.zeros(4, dtype=torch.float32, device=self.device)   # focal, kl, l2, unused
        return g, det, scal

    def _seed_losses(self, g, y, det, scal, w_f, w_kl, with_focal=True):
        """softmax of every head into det; focal loss + its gradient seeds (with_focal), KL + its seeds.
        Loss terms of a replica are scaled 1/R (MirroredStrategy + Keras SUM_OVER_BATCH_SIZE)."""
        eng, nc = self.eng, self.num_classes
        heads = g['heads']
        inv_r = 1.0 / self.world_size
        for hi, (lg, up) in enumerate(heads):
            if not with_focal:
                self._softmax_head(eng, lg, up, det, nc * hi)
                continue
            if isinstance(lg, LazyHead):
                fresh = lg.feat.g is None
                gbuf, acc = eng.grad_buffer(lg.feat)
                if lg.feat.c == lg.feat.lc and ops.logits_softmax_focal(
                        eng.ctx, lg.feat.t, lg.w, lg.b, y, self.focal.alpha, float(self.focal.gamma), det,
                        nc * hi, 1.0 / len(heads), scal[0:1], gbuf, acc, eng.pg(lg.name + "/kernel"),
                        eng.pg(lg.name + "/bias"), w_f * inv_r):
                    continue
                if fresh:
                    lg.feat.g = None
                lg = lg.materialize(eng)
            gbuf, _ = eng.grad_buffer(lg, zero=True)
            ops.softmax_focal(eng.ctx, lg.t, y, self.focal.alpha, float(self.focal.gamma), up, det, nc * hi,
                              1.0 / len(heads), scal[0:1], gbuf, w_f * inv_r)
        for ml_q, ml_p in g['kl_pairs']:
            eng.kl(ml_q, ml_p, scal[1:2])
            eng.kl_seed_grad(ml_q, ml_p, w_kl * self.elbo.beta * inv_r)

    def _train_step_eager(self, x, y, apply_update=True, graphed=False):
        eng = self.eng
        B = x.shape[0]
        g, det, scal = self._forward_train(x)
        if self.train_flops_per_volume is None:
            self.train_flops_per_volume = 3 * eng.conv_flops // B
        inv_r = 1.0 / self.world_size
        w_f, w_kl = self.loss_weights
        eng._timed("losses", 0, lambda: self._seed_losses(g, y, det, scal, w_f, w_kl))
        if self.grad_sync is not None and (not graphed or getattr(self, "_graph_dp_overlap", False)):
            self.grad_sync.pre_fire = eng.join_side
            self.grad_sync.begin(self.params.g, eng.param_uses)
            eng.backward(self.grad_sync.param_done)
            self.grad_sync.finish()
        else:
            eng.backward()
        if apply_update and not graphed:
            eng._timed("adam+repack", 0, lambda: self._apply_update(scal[2:3], inv_r, graphed))
        if isinstance(self.noise, PhiloxNoise) and not graphed:
            self.noise.step += 1
        return dict(detection=det, focal=scal[0:1], kl=scal[1:2], l2=scal[2:3])

    def _apply_update(self, l2_out, gscale=1.0, graphed=False):
        opt, P = self.optimizer, self.params
        lr_t = opt.lr_t()
        for grp, l2 in (('kernel', self.l2_kernel), ('bias', self.l2_bias), ('plain', 0.0)):
            a, b = P.group_range[grp]
            if b > a:
                # L2 term of every rank is scaled 1/R like the data term (MirroredStrategy semantics)
                if graphed:      # step size from device memory: the captured launch is replayed every step
                    ops.adam_amsgrad_dev(self.eng.ctx, P.w[a:b], P.g[a:b], P.m[a:b], P.v[a:b], P.vhat[a:b],
                                         self._lr_dev, opt.beta_1, opt.beta_2, opt.epsilon, l2, 1.0,
                                         l2_out if l2 > 0 else None, bool(opt.amsgrad))
                else:
                    ops.adam_amsgrad(self.eng.ctx, P.w[a:b], P.g[a:b], P.m[a:b], P.v[a:b], P.vhat[a:b], lr_t,
                                     opt.beta_1, opt.beta_2, opt.epsilon, l2, 1.0, l2_out if l2 > 0 else None,
                                     bool(opt.amsgrad))
        if not graphed:
            opt.iterations += 1
        self.eng.refresh_packs()

    def total_loss(self, r):
        """loss value Keras would log: w_f * focal + w_kl * beta * KL + sum(L2)."""
        w_f, w_kl = self.loss_weights
        return w_f * r['focal'] + w_kl * self.elbo.beta * r['kl'] + r['l2']

    def fit(self, x=None, y=None, epochs=1, steps_per_epoch=None, initial_epoch=0, verbose=2, callbacks=None,
            **_):
        """unet_model.fit(x=dataset, epochs, steps_per_epoch, initial_epoch, verbose, callbacks)
        (train_model.py:253-259). `x` is an iterable of (inputs, targets) with inputs {'image': ...} and
        targets {'detection': one-hot, 'KL': ...} (train_model.py:153-164), or arrays x, y."""
        history = {'loss': [], 'detection_loss': [], 'KL_loss': []}
        if y is not None:
            data = [(x, y)]
            steps_per_epoch = steps_per_epoch or 1
        else:
            data = x
        if self.eng is not None and not isinstance(data, (list, tuple)):
            from ..prefetch import DevicePrefetcher          # train_model.py:183 .prefetch(): next batch copied ahead
            data = DevicePrefetcher(data, self.device)
        it = iter(data)
        for cb in callbacks or []:
            getattr(cb, 'on_train_begin', lambda logs=None: None)()
        for epoch in range(initial_epoch, epochs):
            t0 = time.time()
            acc = torch.zeros(3, dtype=torch.float64)
            steps = 0
            while steps_per_epoch is None or steps < steps_per_epoch:
                try:
                    inputs, targets = next(it)
                except StopIteration:
                    if steps_per_epoch is None:
                        break
                    it = iter(data)
                    inputs, targets = next(it)
                r = self.train_step(inputs, targets)
                acc += torch.stack([self.total_loss(r)[0], self.loss_weights[0] * r['focal'][0],
                                    r['kl'][0] * self.elbo.beta]).double().cpu()
                steps += 1
            logs = {k: float(v / max(steps, 1)) for k, v in zip(history, acc)}
            for k, v in logs.items():
                history[k].append(v)
            if verbose:
                print('Epoch %d/%d - %.1fs - loss: %.4f - detection_loss: %.4f - KL_loss: %.4f' % (
                    epoch + 1, epochs, time.time() - t0, logs['loss'], logs['detection_loss'], logs['KL_loss']))
            for cb in callbacks or []:
                getattr(cb, 'on_epoch_end', lambda e, logs=None: None)(epoch, logs)
            it = iter(data) if steps_per_epoch is None else it
        self.history = history
        return history

    def __call__(self, x, training=False):
        return self.predict_train_graph(x) if training else self.get_detect_model()(x)

    def predict_train_graph(self, x, y_true=None):
        """outputs of the training model: [detection (softmax of train_conv [+DS heads]), KL]."""
        eng = self.eng
        x = self._to_device(x)
        B = x.shape[0]
        eng.noise = self.noise
        eng.begin(record=False)
        g = self._graph(eng, B, training=True, x=x)
        nc = self.num_classes
        det = torch.empty((B,) + self.input_spatial_dims + (nc * len(g['heads']),), dtype=torch.float32,
                          device=self.device)
        kl = torch.zeros(1, dtype=torch.float32, device=self.device)
        for hi, (lg, up) in enumerate(g['heads']):
            self._softmax_head(eng, lg, up, det, nc * hi)
        for ml_q, ml_p in g['kl_pairs']:
            eng.kl(ml_q, ml_p, kl)
        return [det, kl]

    def _softmax_head(self, eng, lg, up, out, off):
        """softmax of one head into out[..., off:off+nc] (inference: no labels, no gradients)"""
        if isinstance(lg, LazyHead):
            if lg.feat.c == lg.feat.lc and ops.logits_softmax_focal(
                    eng.ctx, lg.feat.t, lg.w, lg.b, None, None, 0.0, out, off, 0.0, None, None, False, None, None,
                    0.0):
                return
            lg = lg.materialize(eng)
        ops.softmax_focal(eng.ctx, lg.t, None, None, 0.0, up, out, off, 0.0, None, None, 0.0)

    def get_detect_model(self):
        """R:networks.py:196-206: model reconfigured to predict segment probabilities only."""
        return DetectModel(self)

    # ---- weights / optimizer state -------------------------------------------------------------------
    def get_weights(self):
        return self.params.state_dict()

    def set_weights(self, weights, strict=True):
        self.params.load_state_dict(weights, strict)
        if self.eng is not None:
            self.eng.refresh_packs()

    def gradients(self):
        """reference-shaped gradients of the last train_step (host tensors)"""
        return self.params.grad_dict()

    def get_optimizer_state(self):
        P = self.params
        return dict(m=P.m.cpu().numpy(), v=P.v.cpu().numpy(), vhat=P.vhat.cpu().numpy(),
                    iterations=np.array([self.optimizer.iterations if self.optimizer else 0]))

    def set_optimizer_state(self, st):
        P = self.params
        for k in ('m', 'v', 'vhat'):
            getattr(P, k).copy_(torch.from_numpy(np.asarray(st[k])))
        if self.optimizer is None:
            self.compile()
        self.optimizer.iterations = int(np.asarray(st['iterations'])[0])


class DetectModel:
    """tf.keras.Model(inputs, softmax(infer_conv)) of R:networks.py:205-206: one prior pass with
    z ~ P at every level (probabilistic) or the deterministic pass, first num_classes channels."""

    def __init__(self, model):
        self.model = model
        self._pass = 0

    def __call__(self, x):
        return self.predict(x)

    GRAPH_WARMUP = 2

    def _predict_eager(self, x, pass_name):
        m = self.model
        eng = m.eng
        B = x.shape[0]
        eng.noise = m.noise
        eng.begin(record=False)
        lg = m._infer_graph(eng, B, x=x, pass_name=pass_name)
        nc = m.num_classes
        out = torch.empty((B,) + m.input_spatial_dims + (nc,), dtype=torch.float32, device=m.device)
        m._softmax_head(eng, lg, (1, 1, 1), out, 0)
        return out

    def predict(self, x, pass_name='p_sample'):
        """One (stochastic, in 'monte-carlo' mode) forward pass. Like the training step, the ~450 launches of a
        pass are captured once into a CUDA graph over a static input buffer after GRAPH_WARMUP eager calls and
        replayed; the Philox step counter lives in device memory (M1_CUDA_GRAPH=0: always eager). The returned
        tensor of a replayed pass is the graph's static output buffer, overwritten by the next call."""
        import os
        m = self.model
        if m.eng is None:
            raise RuntimeError("model not built on a GPU (m1b200 has no CPU fallback)")
        x = m._to_device(x)
        philox = isinstance(m.noise, PhiloxNoise)
        st = getattr(self, "_gs", None)
        if st is not None and (tuple(st['x'].shape) != tuple(x.shape) or st['pass'] != pass_name):
            st = self._gs = None
            self._eager_calls = 0
        if not philox or m.eng.prof is not None or os.environ.get("M1_CUDA_GRAPH", "1") == "0" \
                or getattr(self, "_graph_failed", False):
            out = self._predict_eager(x, pass_name)
        elif st is None and getattr(self, "_eager_calls", 0) < self.GRAPH_WARMUP:
            self._eager_calls = getattr(self, "_eager_calls", 0) + 1
            out = self._predict_eager(x, pass_name)
        else:
            if st is None:
                st = {'x': torch.empty_like(x), 'step': torch.zeros(1, dtype=torch.int64, device=m.device),
                      'pass': pass_name}
                st['x'].copy_(x)
                st['step'].fill_(m.noise.step)
                torch.cuda.synchronize(m.device)
                g = torch.cuda.CUDAGraph()
                m.noise.step_dev = st['step']
                before = m.eng.ctx.launch_count()
                try:
                    with torch.cuda.graph(g):
                        st['out'] = self._predict_eager(st['x'], pass_name)
                except Exception:
                    self._graph_failed = True
                    raise
                finally:
                    m.noise.step_dev = None
                st['graph'], st['launches'] = g, m.eng.ctx.launch_count() - before
                self._gs = st
            else:
                st['x'].copy_(x, non_blocking=True)
            st['step'].fill_(m.noise.step)
            st['graph'].replay()
            self.graph_replays = getattr(self, "graph_replays", 0) + 1
            out = st['out']
        if philox:
            m.noise.step += 1
        return out

    @property
    def launches_per_graph_pass(self):
        """kernels inside the most recently used replayed graph (a single pass, or a whole ensemble)"""
        st = getattr(self, "_gs_mc", None) or getattr(self, "_gs", None)
        return st['launches'] if st else None

    def _predict_mc_eager(self, x, passes, pass_name='p_sample'):
        """`passes` stochastic passes over ONE batch in one engine scope: the input activations are built once, so
        the stem + serse1 trunk (nothing stochastic before the first dropout) is computed once and shared; pass i
        draws the Philox streams of step + i (PhiloxNoise.sub), exactly those of `passes` separate predict() calls."""
        m = self.model
        eng = m.eng
        B = x.shape[0]
        eng.noise = m.noise
        eng.begin(record=False)
        img = m._inputs(eng, B, x, False)[0]
        nc = m.num_classes
        mean = torch.zeros((B,) + m.input_spatial_dims + (nc,), dtype=torch.float32, device=m.device)
        philox = isinstance(m.noise, PhiloxNoise)
        try:
            for i in range(passes):
                if philox:
                    m.noise.sub = i
                lg = m._infer_graph(eng, B, pass_name=pass_name, inputs=img)
                out = torch.empty_like(mean)
                m._softmax_head(eng, lg, (1, 1, 1), out, 0)
                ops.axpy(eng.ctx, out, 1.0 / passes, mean)
        finally:
            if philox:
                m.noise.sub = 0
        return mean

    def predict_mc(self, x, passes=20):
        """Monte-Carlo dropout ensemble (BASELINE config 4): mean softmax of `passes` stochastic prior passes with
        fresh Philox streams (the loop itself is not in the reference: the flag UNET_PROBA_ITER of train_model.py:71
        is unused there). After GRAPH_WARMUP eager calls the WHOLE ensemble - shared trunk, `passes` passes, the
        running mean - is captured once into one CUDA graph and replayed (M1_CUDA_GRAPH=0: always eager)."""
        import os
        m = self.model
        if m.eng is None:
            raise RuntimeError("model not built on a GPU (m1b200 has no CPU fallback)")
        x = m._to_device(x)
        philox = isinstance(m.noise, PhiloxNoise)
        st = getattr(self, "_gs_mc", None)
        if st is not None and (tuple(st['x'].shape) != tuple(x.shape) or st['passes'] != passes):
            st = self._gs_mc = None
            self._mc_eager_calls = 0
        if not philox or m.eng.prof is not None or os.environ.get("M1_CUDA_GRAPH", "1") == "0" \
                or getattr(self, "_graph_failed", False):
            out = self._predict_mc_eager(x, passes)
        elif st is None and getattr(self, "_mc_eager_calls", 0) < self.GRAPH_WARMUP:
            self._mc_eager_calls = getattr(self, "_mc_eager_calls", 0) + 1
            out = self._predict_mc_eager(x, passes)
        else:
            if st is None:
                st = {'x': torch.empty_like(x), 'step': torch.zeros(1, dtype=torch.int64, device=m.device),
                      'passes': passes}
                st['x'].copy_(x)
                st['step'].fill_(m.noise.step)
                torch.cuda.synchronize(m.device)
                g = torch.cuda.CUDAGraph()
                m.noise.step_dev = st['step']
                before = m.eng.launch_total()
                try:
                    with torch.cuda.graph(g):
                        st['out'] = self._predict_mc_eager(st['x'], passes)
                except Exception:
                    self._graph_failed = True
                    raise
                finally:
                    m.noise.step_dev = None
                st['graph'], st['launches'] = g, m.eng.launch_total() - before
                self._gs_mc = st
            else:
                st['x'].copy_(x, non_blocking=True)
            st['step'].fill_(m.noise.step)
            st['graph'].replay()
            self.graph_replays = getattr(self, "graph_replays", 0) + 1
            out = st['out']
        if philox:
            m.noise.step += passes
        return out


def m1(*args, **kwargs):
    """R:networks.py:232 - the mid-level wrapper is folded into M1 (one stage = one M1 object)."""
    raise NotImplementedError("use M1(...): the functional m1() wiring lives in M1._graph")
