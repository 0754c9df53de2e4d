"""colli2de_b200 — B200-native implementation of Colli2De's per-frame collision query path.

The product is native: ``libc2d_gpu.so`` (hand-written sm_100a kernels behind the C ABI of ``include/c2d_gpu.h``)
and the C++23 header ``include/colli2de/Registry.hpp`` (drop-in ``c2d::Registry``).  This Python package only
holds the in-tree build (:mod:`colli2de_b200.build`), the synthetic scene generator used by tests and bench
(:mod:`colli2de_b200.scenes`) and thin ctypes access to a few C-ABI entry points (:mod:`colli2de_b200.capi`).
"""
