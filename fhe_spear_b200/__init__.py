"""fhe_spear_b200 -- B200-native (sm_100a) CKKS BSGS diagonal matrix-vector engine.

Layout:
  csrc/        hand-written CUDA kernels + the C ABI (include/spear_b200.h) -> libspear_b200.so
  _native.py   ctypes binding of that library (mandatory; no CPU fallback)
  pyPhantom/   drop-in for the reference's `pyPhantom` pybind11 module (gpu/phantom_binding.cu)
  bsgs.py      host-side mirror of the reference's BSGS layer (scripts/bootstrap_generation.py:18-660)
  sharding.py  multi-GPU split of giant steps / projections over torch.distributed
"""
__version__ = "0.1.0"
