#!/usr/bin/env python
"""Dump weights + activations of the REAL reference (TF 2.5) for a fixed seed, to pin the oracle (SURVEY.md 8c
"honest limitation"): run this where `tensorflow-gpu==2.5.0`, `tensorflow_addons==0.14.0`,
`tensorflow_probability==0.13.0` and `sonnet` are installed (requirements.txt of the reference); it cannot run in
this repository's image (no TensorFlow - the script says so and exits 2), so it has never been executed here.

    python tools/dump_tf_reference.py --reference /path/to/prostateMR_3D-CAD-csPCa --out tests/golden/tf_m1_det.npz

What it writes (npz): the input volume, every Keras variable of the model under its Keras name, in creation order
(`var/<index>/<name>`), and the model outputs. The loader that maps those variables onto oracle.m1_oracle.ParamStore
(rename by layer order, R:networks.py:472-565) is NOT written yet: without a single real dump to test against it
would be guesswork. Until such a file exists the oracle stays "parity unpinned" (DESIGN.md).
Deterministic configuration only (dropout_mode='standard', training=False): the stochastic sites need TF's own
random streams, which no other implementation can reproduce."""
import argparse
import os
import sys

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="checkout of DIAGNijmegen/prostateMR_3D-CAD-csPCa")
    ap.add_argument("--out", default="tests/golden/tf_m1_det.npz")
    ap.add_argument("--dims", type=int, nargs=3, default=(8, 32, 32))
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    try:
        import tensorflow as tf
    except ImportError:
        print("dump_tf_reference: TensorFlow is not importable here - run this in the reference's environment")
        return 2
    sys.path.insert(0, os.path.join(a.reference, "tf2.5", "scripts"))
    import model.unets as unets                                     # the reference's own package

    tf.random.set_seed(a.seed)
    np.random.seed(a.seed)
    filters = (8, 16, 24, 32, 48)
    m = unets.networks.M1(input_spatial_dims=tuple(a.dims), input_channels=3, num_classes=2,
                          filters=filters, strides=((1, 1, 1), (1, 2, 2), (1, 2, 2), (2, 2, 2), (2, 2, 2)),
                          kernel_sizes=((1, 3, 3), (1, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3)),
                          se_reduction=(4,) * 5, att_sub_samp=((1, 1, 1),) * 4, dropout_rate=0.0,
                          dropout_mode='standard', dense_skip=True, deep_supervision=True, probabilistic=False,
                          cascaded=False, summary=False)
    x = np.random.RandomState(a.seed).randn(2, *a.dims, 3).astype(np.float32)
    out = m(x, training=False)
    out = out if isinstance(out, (list, tuple)) else [out]
    blob = {"x": x}
    for i, v in enumerate(m.variables):
        blob["var/%04d/%s" % (i, v.name)] = v.numpy()
    for i, o in enumerate(out):
        blob["out/%d" % i] = np.asarray(o)
    np.savez_compressed(a.out, **blob)
    print("wrote %s: %d variables, %d outputs" % (a.out, len(m.variables), len(out)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
