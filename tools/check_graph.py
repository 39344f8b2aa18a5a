#!/usr/bin/env python
"""CUDA-graph replay of the training step vs the eager launch sequence, same model state and Philox step:
captures at step 2, then re-runs the same steps eagerly from a snapshot and compares losses and weights.
The forward pass is deterministic (fixed-order reductions, no floating-point atomics on forward values), so the
FIRST compared step - same weights, same Philox step - must agree BIT FOR BIT between graph and eager and between
two eager runs. Later steps start from weights that carry the split-K fp32 atomics of the weight gradients and
are compared with a tolerance."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import m1b200  # noqa: E402,F401
from m1b200.model import losses, optimizers, unets  # noqa: E402
from oracle import m1_oracle as O  # noqa: E402


def main():
    strides = ((1, 1, 1), (1, 2, 2), (1, 2, 2), (2, 2, 2), (2, 2, 2))
    kernels = ((1, 3, 3), (1, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3))
    m = unets.networks.M1((8, 32, 32), 4, 2, dropout_mode='monte-carlo', filters=(32, 64, 128, 192, 256),
                          strides=strides, kernel_sizes=kernels, se_reduction=(8,) * 5,
                          att_sub_samp=((1, 1, 1),) * 4, dense_skip=True, deep_supervision=True, probabilistic=True,
                          prob_latent_dims=(3, 2, 1, 0), summary=False, precision=os.environ.get('M1_PRECISION', 'fp16'),
                          device='cuda:0', seed=0)
    sched = optimizers.CosineDecayRestarts(1e-3, 5, t_mul=2.0, m_mul=1.0, alpha=1e-3)
    m.compile(optimizer=optimizers.Adam(sched, amsgrad=True),
              loss=[losses.Focal(alpha=[0.75, 0.25], gamma=2.0).loss, losses.EvidenceLowerBound().loss],
              loss_weights=[1.0, 10.0])
    x, y = O.synthetic_batch(2, (8, 32, 32), seed=3)
    for _ in range(m.GRAPH_WARMUP):
        m.train_step(x, y)
    snap = (m.get_weights(), m.get_optimizer_state(), m.noise.step, m.optimizer.iterations)

    def run(n):
        ls = []
        for _ in range(n):
            r = m.train_step(x, y)
            ls.append([float(m.total_loss(r)), float(r['focal']), float(r['kl'])])
        return np.array(ls), m.get_weights()

    os.environ["M1_CUDA_GRAPH"] = "1"
    lg, wg = run(4)
    assert m.graph_replays == 4, m.graph_replays
    os.environ["M1_CUDA_GRAPH"] = "0"

    def restore():
        m.set_weights(snap[0]); m.set_optimizer_state(snap[1]); m.noise.step = snap[2]
        m.optimizer.iterations = snap[3]
        m.eng.refresh_packs()
    restore()
    le, we = run(4)
    restore()
    le2, we2 = run(4)          # eager twice: the run-to-run floor (fp32 atomics + bf16 rounding flips)
    rel = np.abs(lg - le) / np.abs(le)
    floor = np.abs(le2 - le) / np.abs(le)
    wd = max(float(np.abs(wg[k] - we[k]).mean()) for k in wg)
    wf = max(float(np.abs(we2[k] - we[k]).mean()) for k in wg)
    print("graph :", lg[:, 0])
    print("eager :", le[:, 0])
    print("eager2:", le2[:, 0])
    print("rel loss diff graph-vs-eager: first step %.3e max %.3e | eager-vs-eager floor: first %.3e max %.3e"
          % (rel[0].max(), rel.max(), floor[0].max(), floor.max()))
    print("worst mean |dw| graph-vs-eager %.3e | eager-vs-eager %.3e | launches per replayed step %d"
          % (wd, wf, m.launches_per_graph_step))
    # first compared step: same weights, same Philox step, deterministic forward -> identical bits
    assert np.array_equal(lg[0], le[0]), ("graph vs eager, first step", lg[0], le[0])
    assert np.array_equal(le2[0], le[0]), ("eager vs eager, first step", le2[0], le[0])
    # later steps: the weights carry fp32 atomic summation-order noise of the weight gradients, amplified by Adam
    assert rel.max() < 5e-2, (rel, floor)
    print("GRAPH OK")


if __name__ == "__main__":
    main()
