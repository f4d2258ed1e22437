"""B200-native explicit-ALS factor update for You-Can-(Not)-Recommend (hot path only).

Layout: csrc/ (CUDA kernels + C ABI + host front end), emf_base/emf_worker/emf_master
(host-side mirror of the reference's worker interface), front_end (data side), native (ctypes).
"""
__version__ = "0.1.0"
