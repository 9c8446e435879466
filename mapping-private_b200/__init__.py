"""B200-native normals -> RSD -> GRSD hot path of cloud_algos.

Layout:
  csrc/   CUDA kernels for sm_100a + the C-ABI library (libcloudalgos_b200.so)
  host/   C++ mirror of the reference's CloudAlgo plugin surface
  cab.py  ctypes binding of the C ABI (used by tests and bench.py)
  synth.py  deterministic synthetic clouds for the BASELINE.json configs
"""
