"""pymcz_b200 -- B200-native (sm_100a) implementation of pyMCZ's Monte-Carlo metallicity hot path.

The package holds only what that path needs:
  csrc/            hand-written CUDA kernels + the C ABI (include/mcz_b200.h) -> libmcz_b200.so
  _lib, ops        ctypes binding and torch-tensor entry points
  metallicity, metscales, mcz   py3 host-side mirror of the reference's Python surface
  synth, dist      benchmark tables and galaxy sharding over GPUs
"""
__version__ = "0.1.0"
