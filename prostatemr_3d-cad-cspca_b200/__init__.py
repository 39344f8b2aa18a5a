"""m1b200 — B200-native (sm_100a) implementation of the M1 Hierarchical Probabilistic 3D U-Net
forward/backward path of DIAGNijmegen/prostateMR_3D-CAD-csPCa.

Layout
  csrc/          hand-written CUDA kernels + the C-ABI (include/m1b200.h) -> lib/libm1b200.so
  _lib.py        ctypes binding of the C-ABI, DLPack device-pointer export
  ops.py         one thin Python wrapper per C entry point
  model/         host-side mirror of the reference API: model.unets.networks.M1, model.losses.Focal
"""
__version__ = "0.1.0"
