"""nk_ooc_b200 — B200-native batched function evaluation for Newton-Krylov_OOC.

Host-side mirror of the reference's operator surface (ModelStateBase /
TracerModuleStateBase, models test_problem and py_driver_2d) over the CUDA library
libnkb200.so (C ABI in include/nkb200.h).  There is no CPU fallback.
"""

__version__ = "0.1.0"
