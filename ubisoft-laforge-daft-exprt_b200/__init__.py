"""daft_exprt_b200 — Blackwell-native (sm_100a) implementation of Daft-Exprt's mel-prediction forward/training path
behind the reference's own `DaftExprt` / `DaftExprtLoss` / `hparams` surface (reference `model.py`, `loss.py`).

Layout:  csrc/ (CUDA kernels + the C-ABI, built into libdaftexprt_b200.so)  ·  cabi.py (ctypes binding of
include/daft_exprt_b200.h)  ·  ops.py (autograd functions over the C-ABI)  ·  model.py / loss.py (drop-in modules)  ·
synthetic.py (seeded synthetic batches)  ·  ddp.py (flat-bucket gradient all-reduce over NCCL).
"""
__version__ = '0.1.0'
