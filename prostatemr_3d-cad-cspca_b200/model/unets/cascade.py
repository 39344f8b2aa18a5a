"""The cascaded two-stage M1 of R:networks.py:109-193 (+ decision_fusion, :209-223).

  stage 1   m1(image_1)                                                    -> softmax_1 (+ KL_1)
  stage 2   m1(concat([softmax_1[..., :nc-1], image_2]))                   -> softmax_2 (+ KL_2)
  outputs   detection_1 = [1 - p1, p1], detection_2 = [1 - j, j], j = fusion(p1, p2)   (p = class nc-1 probability)

`cascaded` is at once the switch and the fusion strategy ('identity' | 'noisy-or' | 'bayes'); True matches no branch of
the reference's decision_fusion (UnboundLocalError, Q8) and is read as 'identity', its default. For nc = 2 the slice
softmax_1[..., :nc-1] is the BACKGROUND probability - reproduced as the reference slices it.

Host logic only. The two stages are ordinary single-stage M1 objects (own parameter table and engine; stage 2 has
input_channels + nc - 1 input channels and parameter names prefixed 'stage2/'); what the cascade adds on the device is
  m1_fusion_focal      decision fusion + Focal.FL of detection_2 + d/dp1, d/dp2           (strategy != identity)
  m1_logits_prob_bwd   softmax + logits-conv backward for a gradient w.r.t. PROBABILITIES: stage 1 receives the
                       gradient of stage 2's input channel (its background probability) and of the fusion
The whole step - both stages, fusion, backward of stage 2, then of stage 1, Adam of both - is captured into CUDA graphs
and replayed like the single-stage step. Only two classes: Focal.loss over a 2-channel [1-p, p] prediction has no head
for any other class count (losses.py:43-49)."""
import os

import numpy as np
import torch

from ... import ops
from ..losses import EvidenceLowerBound, Focal
from ..optimizers import Adam
from .engine import LazyHead, PhiloxNoise
from .modelio import LoadableModel, store_config_args
from .networks import FUSION, M1, DetectModel

STAGE2 = 'stage2/'


class CascadedM1(M1):
    @store_config_args
    def __init__(self, input_spatial_dims, input_channels, num_classes, dropout_rate=0.50, dropout_mode='standard',
                 filters=(32, 64, 128, 256, 512),
                 strides=((1, 1, 1), (1, 2, 2), (1, 2, 2), (2, 2, 2), (1, 2, 2)),
                 kernel_sizes=((1, 3, 3), (1, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3)),
                 se_reduction=(8, 8, 8, 8, 8), att_sub_samp=((1, 1, 1), (1, 1, 1), (1, 1, 1)),
                 kernel_initializer=None, bias_initializer=None, kernel_regularizer=None, bias_regularizer=None,
                 cascaded='identity', dense_skip=False, deep_supervision=False, probabilistic=False,
                 prob_latent_dims=(3, 2, 1), summary=True, name='UNET-TYPE-M1',
                 *, precision='fp16', ds_in_prob='reference', seed=0, device=None, build=None, use_tcgen05=True,
                 compute_dead_branches=False):
        if cascaded not in FUSION:
            raise ValueError("cascaded must be False, 'identity', 'noisy-or', 'bayes' (or True == 'identity', Q8); "
                             "got %r" % (cascaded,))
        assert num_classes == 2, "the cascade's [1-p, p] outputs exist for two classes only (losses.py:43-49)"
        self.name = name
        self.strategy = FUSION[cascaded]
        self.input_spatial_dims = tuple(int(d) for d in input_spatial_dims)
        self.input_channels, self.num_classes = int(input_channels), int(num_classes)
        self.probabilistic, self.precision = bool(probabilistic), precision
        kw = dict(dropout_rate=dropout_rate, dropout_mode=dropout_mode, filters=filters, strides=strides,
                  kernel_sizes=kernel_sizes, se_reduction=se_reduction, att_sub_samp=att_sub_samp, cascaded=False,
                  dense_skip=dense_skip, deep_supervision=deep_supervision, probabilistic=probabilistic,
                  prob_latent_dims=prob_latent_dims, summary=summary, precision=precision, ds_in_prob=ds_in_prob,
                  device=device, build=build, use_tcgen05=use_tcgen05, compute_dead_branches=compute_dead_branches)
        for k_, v in (('kernel_initializer', kernel_initializer), ('bias_initializer', bias_initializer),
                      ('kernel_regularizer', kernel_regularizer), ('bias_regularizer', bias_regularizer)):
            if v is not None:
                kw[k_] = v
        self.stage1 = M1(input_spatial_dims, input_channels, num_classes, seed=seed, **kw)
        self.stage2 = M1(input_spatial_dims, input_channels + num_classes - 1, num_classes, seed=seed + 1, **kw)
        self.stages = (self.stage1, self.stage2)
        self.references = LoadableModel.ReferenceContainer()
        self.references.m1_stage1, self.references.m1_stage2 = self.stage1, self.stage2
        self.references.cascaded, self.references.probabilistic = cascaded, probabilistic
        self.references.num_classes = num_classes
        if summary:
            print('Cascade Prior Prediction (Softmax):-----', (None,) + self.input_spatial_dims + (2,))
            print('Cascade Follow-Up Prediction (Softmax):-', (None,) + self.input_spatial_dims + (2,))
        self.optimizer, self.history = None, None
        self.world_size, self.rank = 1, 0
        self.loss_weights4 = [1.0, 1.0, 1.0, 1.0]

    # ---- plumbing shared with M1 ------------------------------------------------------------------
    @property
    def eng(self):
        return self.stage1.eng

    @property
    def device(self):
        return self.stage1.device

    def engines(self):
        return [s.eng for s in self.stages]

    def summary(self):
        for s in self.stages:
            s.summary()

    def compile(self, optimizer=None, loss=None, loss_weights=None, **_):
        """Keras order of the four outputs: [detection_1, detection_2, KL_1, KL_2]; a 2-element loss / loss_weights
        list [Focal, ELBO] is applied to both stages."""
        self.optimizer = optimizer if optimizer is not None else Adam(1e-3, amsgrad=True)
        losses = list(loss) if isinstance(loss, (list, tuple)) else [loss]
        objs = [getattr(ls, '__self__', ls) for ls in losses if ls is not None]
        focals = [o for o in objs if isinstance(o, Focal)]
        elbos = [o for o in objs if isinstance(o, EvidenceLowerBound)]
        lw = list(loss_weights) if loss_weights is not None else [1.0] * 4
        if len(lw) == 2:
            lw = [lw[0], lw[0], lw[1], lw[1]]
        assert len(lw) == 4, "loss_weights: [detection_1, detection_2, KL_1, KL_2] or [detection, KL]"
        self.loss_weights4 = [float(v) for v in lw]
        for i, s in enumerate(self.stages):
            s.compile(optimizer=self.optimizer,
                      loss=[(focals[min(i, len(focals) - 1)] if focals else None),
                            (elbos[min(i, len(elbos) - 1)] if elbos else None)] if objs else None,
                      loss_weights=[lw[i], lw[2 + i]])
            s.optimizer = self.optimizer
        return self

    def distribute(self, bucket_bytes=32 << 20, group=None):
        for s in self.stages:
            s.distribute(bucket_bytes, group)
        self.world_size, self.rank = self.stage1.world_size, self.stage1.rank
        return self

    def set_noise(self, tensors=None, seed=None):
        """{(pass_name, site): tensor}; stage-2 keys carry the 'stage2/' prefix on the pass name (oracle.m1_cascade)"""
        if tensors is None:
            self.stage1.set_noise(None, seed)
            self.stage2.set_noise(None, None if seed is None else seed + 1)
            return
        self.stage1.set_noise({k: v for k, v in tensors.items() if not k[0].startswith(STAGE2)})
        self.stage2.set_noise({(k[0][len(STAGE2):], k[1]): v for k, v in tensors.items() if k[0].startswith(STAGE2)})

    @staticmethod
    def _pair(x):
        if isinstance(x, dict):
            return x['image_1'], x['image_2']
        if isinstance(x, (list, tuple)):
            return x[0], x[1]
        return x, x

    # ---- training step ------------------------------------------------------------------------------
    def _stage2_input(self, det1, x2):
        """concat([softmax_1[..., :nc-1], image_2]) as an fp32 NDHWC tensor (R:networks.py:129-130)"""
        s1 = self.stage1
        B, C, k = x2.shape[0], self.input_channels, self.num_classes - 1
        xin = torch.empty((B,) + self.input_spatial_dims + (C + k,), dtype=torch.float32, device=s1.device)
        ops.copy_channels(s1.eng.ctx, det1, 0, xin, 0, k)
        ops.copy_channels(s1.eng.ctx, x2, 0, xin, k, C)
        return xin

    def _step_eager(self, x1, x2, y, graphed=False):
        s1, s2 = self.stages
        w_d1, w_d2, w_k1, w_k2 = self.loss_weights4
        inv_r = 1.0 / self.world_size
        nc = self.num_classes
        ctx = s1.eng.ctx
        # ---- stage 1: forward, focal(detection_1) (detection_1 = softmax_1 for two classes), KL_1
        g1, det1, sc1 = s1._forward_train(x1)
        g1['heads'] = g1['heads'][:1]        # the cascade's outputs only see head 0: deep-supervision heads are dead
        s1._seed_losses(g1, y, det1, sc1, w_d1, w_k1)
        # ---- stage 2 on [background probability of stage 1 | image_2], differentiated w.r.t. its input
        xin = self._stage2_input(det1, x2)
        g2, det2, sc2 = s2._forward_train(xin, input_needs_grad=True)
        g2['heads'] = g2['heads'][:1]
        rows = det1.numel() // det1.shape[-1]
        dprob1 = torch.zeros((rows, nc), dtype=torch.float32, device=s1.device)
        joint = det2
        if self.strategy == 0:
            s2._seed_losses(g2, y, det2, sc2, w_d2, w_k2)                   # detection_2 = softmax_2
        else:
            s2._seed_losses(g2, y, det2, sc2, 0.0, w_k2, with_focal=False)  # softmax only + KL_2
            joint = torch.empty((det2.shape[:-1]) + (2,), dtype=torch.float32, device=s1.device)
            dp1 = torch.empty(rows, dtype=torch.float32, device=s1.device)
            dp2 = torch.empty(rows, dtype=torch.float32, device=s1.device)
            ops.fusion_focal(ctx, det1, nc - 1, det2, nc - 1, self.strategy, y, s2.focal.alpha, float(s2.focal.gamma),
                             None, joint, 1.0, sc2[0:1], dp1, dp2, w_d2 * inv_r)
            dprob2 = torch.zeros((rows, nc), dtype=torch.float32, device=s1.device)
            ops.copy_channels(ctx, dp2.view(rows, 1), 0, dprob2, nc - 1, 1)
            ops.copy_channels(ctx, dp1.view(rows, 1), 0, dprob1, nc - 1, 1)
            self._head_prob_bwd(s2, g2, det2, dprob2)
        s2.eng.backward()
        # ---- gradient of stage 2's input channel 0 (= softmax_1[..., 0]) -> stage 1
        tmp = torch.zeros((rows, nc), dtype=torch.float32, device=s1.device)
        for a in g2['inputs']:
            if a.g is None:
                continue
            ops.copy_channels(ctx, a.g, 0, tmp, 0, nc - 1)
            ops.axpy(ctx, tmp, 1.0, dprob1)
        self._head_prob_bwd(s1, g1, det1, dprob1)
        s1.eng.backward()
        return dict(detection_1=det1[..., :nc], detection_2=joint[..., :nc], focal_1=sc1[0:1], focal_2=sc2[0:1],
                    kl_1=sc1[1:2], kl_2=sc2[1:2], l2_1=sc1[2:3], l2_2=sc2[2:3])

    @staticmethod
    def _head_prob_bwd(stage, g, det, dprob):
        lg = g['heads'][0][0]
        if not isinstance(lg, LazyHead):
            raise NotImplementedError("cascade: the final logits convolution must be the fused head (32 features, 2 "
                                      "classes)")
        eng = stage.eng
        gbuf, acc = eng.grad_buffer(lg.feat)
        if not ops.logits_prob_bwd(eng.ctx, lg.feat.t, lg.w, det, 0, dprob, gbuf, acc, eng.pg(lg.name + "/kernel"),
                                   eng.pg(lg.name + "/bias")):
            raise NotImplementedError("cascade: no m1_logits_prob_bwd instantiation for this head")

    def _update(self, out, graphed=False):
        inv_r = 1.0 / self.world_size
        for s, l2 in zip(self.stages, (out['l2_1'], out['l2_2'])):
            if graphed:
                s._lr_dev = self._lr_dev
            s._apply_update(l2, inv_r, graphed)
            if not graphed:
                s.optimizer.iterations -= 1          # one shared optimizer: count the step once (below)
        if not graphed:
            self.optimizer.iterations += 1

    GRAPH_WARMUP = 2

    def train_step(self, x, y_true, apply_update=True):
        """x: [image_1, image_2] (or {'image_1':..,'image_2':..}; one array feeds both), y_true one-hot.
        Returns dict(detection_1, detection_2, focal_1/2, kl_1/2, l2_1/2) of device tensors."""
        s1, s2 = self.stages
        if s1.eng is None:
            raise RuntimeError("M1.train_step: model not built on a GPU (m1b200 has no CPU fallback)")
        if self.optimizer is None:
            self.compile()
        x1, x2 = (s1._to_device(t) for t in self._pair(x))
        y = s1._to_device(y_true)
        philox = all(isinstance(s.noise, PhiloxNoise) for s in self.stages)
        use_graph = (apply_update and philox and s1.eng.prof is None and os.environ.get("M1_CUDA_GRAPH", "1") != "0"
                     and self.world_size == 1 and not getattr(self, "_graph_failed", False))
        st = getattr(self, "_gs", None)
        if use_graph and (st is not None or getattr(self, "_eager_steps", 0) >= self.GRAPH_WARMUP):
            return self._step_graphed(x1, x2, y)
        self._eager_steps = getattr(self, "_eager_steps", 0) + 1
        out = self._step_eager(x1, x2, y)
        if self.world_size > 1:              # data parallel: one flat all-reduce per stage (no overlap in the cascade)
            import torch.distributed as dist
            for s in self.stages:
                dist.all_reduce(s.params.g, op=dist.ReduceOp.SUM,
                                group=s.grad_sync.group if s.grad_sync is not None else None)
        if apply_update:
            self._update(out)
        for s in self.stages:
            if isinstance(s.noise, PhiloxNoise):
                s.noise.step += 1
        return out

    def _step_graphed(self, x1, x2, y):
        s1, s2 = self.stages
        st = getattr(self, "_gs", None)
        if st is not None and tuple(st['x1'].shape) != tuple(x1.shape):
            st = self._gs = None
        if st is None:
            st = dict(x1=x1.clone(), x2=x2.clone(), y=y.clone(),
                      step=torch.zeros(1, dtype=torch.int64, device=s1.device),
                      lr=torch.zeros(1, dtype=torch.float32, device=s1.device))
            st['step'].fill_(s1.noise.step)
            st['lr'].fill_(self.optimizer.lr_t())
            torch.cuda.synchronize(s1.device)
            g = torch.cuda.CUDAGraph()
            for s in self.stages:
                s.noise.step_dev = st['step']
            self._lr_dev = st['lr']
            before = s1.eng.ctx.launch_count()
            try:
                with torch.cuda.graph(g):
                    st['out'] = self._step_eager(st['x1'], st['x2'], st['y'], graphed=True)
                    self._update(st['out'], graphed=True)
            except Exception:
                self._graph_failed = True
                raise
            finally:
                for s in self.stages:
                    s.noise.step_dev = None
            st['graph'], st['launches'] = g, s1.eng.ctx.launch_count() - before
            self._gs = st
        else:
            st['x1'].copy_(x1, non_blocking=True)
            st['x2'].copy_(x2, non_blocking=True)
            st['y'].copy_(y, non_blocking=True)
        st['step'].fill_(s1.noise.step)
        st['lr'].fill_(self.optimizer.lr_t())
        st['graph'].replay()
        self.graph_replays = getattr(self, "graph_replays", 0) + 1
        self.optimizer.iterations += 1
        for s in self.stages:
            s.noise.step += 1
        return st['out']

    @property
    def launches_per_graph_step(self):
        st = getattr(self, "_gs", None)
        return st['launches'] if st else None

    def total_loss(self, r):
        """w_d1 Focal(detection_1) + w_d2 Focal(detection_2) + w_k1 beta KL_1 + w_k2 beta KL_2 + sum(L2)"""
        w_d1, w_d2, w_k1, w_k2 = self.loss_weights4
        b1, b2 = self.stage1.elbo.beta, self.stage2.elbo.beta
        return (w_d1 * r['focal_1'] + w_d2 * r['focal_2'] + w_k1 * b1 * r['kl_1'] + w_k2 * b2 * r['kl_2']
                + r['l2_1'] + r['l2_2'])

    def fit(self, x=None, y=None, epochs=1, steps_per_epoch=None, initial_epoch=0, verbose=2, callbacks=None, **_):
        """as M1.fit; inputs {'image_1', 'image_2'}, targets {'detection_1', 'detection_2', ...} (same label)"""
        history = {'loss': []}
        data = [(x, y)] if y is not None else x
        steps_per_epoch = steps_per_epoch or (1 if y is not None else None)
        for epoch in range(initial_epoch, epochs):
            it, acc, steps = iter(data), 0.0, 0
            while steps_per_epoch is None or steps < steps_per_epoch:
                try:
                    inputs, targets = next(it)
                except StopIteration:
                    if steps_per_epoch is None:
                        break
                    it = iter(data)
                    inputs, targets = next(it)
                if isinstance(targets, dict):
                    targets = targets.get('detection_2', targets.get('detection_1', targets.get('detection')))
                acc += float(self.total_loss(self.train_step(inputs, targets))[0])
                steps += 1
            history['loss'].append(acc / max(steps, 1))
            if verbose:
                print('Epoch %d/%d - loss: %.4f' % (epoch + 1, epochs, history['loss'][-1]))
            for cb in callbacks or []:
                getattr(cb, 'on_epoch_end', lambda e, logs=None: None)(epoch, {'loss': history['loss'][-1]})
        self.history = history
        return history

    # ---- inference -----------------------------------------------------------------------------------
    def get_detect_model(self):
        """R:networks.py:196-201: [softmax of stage 1, softmax of stage 2] (inference graphs)"""
        return CascadedDetectModel(self)

    def __call__(self, x, training=False):
        return self.get_detect_model()(x)

    # ---- weights ----------------------------------------------------------------------------------------
    def get_weights(self):
        w = dict(self.stage1.get_weights())
        w.update({STAGE2 + k: v for k, v in self.stage2.get_weights().items()})
        return w

    def set_weights(self, weights, strict=True):
        self.stage1.set_weights({k: v for k, v in weights.items() if not k.startswith(STAGE2)}, strict)
        self.stage2.set_weights({k[len(STAGE2):]: v for k, v in weights.items() if k.startswith(STAGE2)}, strict)

    def gradients(self):
        g = dict(self.stage1.gradients())
        g.update({STAGE2 + k: v for k, v in self.stage2.gradients().items()})
        return g

    def get_optimizer_state(self):
        a, b = self.stage1.get_optimizer_state(), self.stage2.get_optimizer_state()
        st = {k: a[k] for k in ('m', 'v', 'vhat')}
        st.update({STAGE2 + k: b[k] for k in ('m', 'v', 'vhat')})
        st['iterations'] = np.array([self.optimizer.iterations if self.optimizer else 0])
        return st

    def set_optimizer_state(self, st):
        if self.optimizer is None:
            self.compile()
        it = st['iterations']
        self.stage1.set_optimizer_state({**{k: st[k] for k in ('m', 'v', 'vhat')}, 'iterations': it})
        self.stage2.set_optimizer_state({**{k: st[STAGE2 + k] for k in ('m', 'v', 'vhat')}, 'iterations': it})
        self.optimizer.iterations = int(np.asarray(it)[0])


class CascadedDetectModel:
    """tf.keras.Model(inputs, [infer_softmax_1, infer_softmax_2]) of R:networks.py:197-201"""

    def __init__(self, model):
        self.model = model
        self.d1, self.d2 = DetectModel(model.stage1), DetectModel(model.stage2)

    def __call__(self, x):
        return self.predict(x)

    def predict(self, x):
        m = self.model
        x1, x2 = (m.stage1._to_device(t) for t in m._pair(x))
        p1 = self.d1.predict(x1)
        p2 = self.d2.predict(m._stage2_input(p1, x2))
        return [p1, p2]

    def predict_mc(self, x, passes=20):
        mean = None
        for _ in range(passes):
            p1, p2 = self.predict(x)
            if mean is None:
                mean = [torch.zeros_like(p1), torch.zeros_like(p2)]
            ops.axpy(self.model.eng.ctx, p1, 1.0 / passes, mean[0])
            ops.axpy(self.model.eng.ctx, p2, 1.0 / passes, mean[1])
        return mean
