"""noname_b200 -- B200-native particle-mesh gravity step behind NoName's operator surface.

Host side (this package) mirrors the reference's Julia modules for the PM path:
ParticleMesh (operators, ICs, evolve loops), cosmology (a(t) scalars), GridTypes, NoName
(run_noname driver), conf (AppConf files).  All arithmetic of the hot path runs in
libnnpm.so (noname_b200/csrc, C ABI in include/nnpm.h) on the GPU; there is no CPU fallback.
"""
from ._lib import NNPMError, LIB_PATH  # noqa: F401
from .device import DevicePM, pinned_empty, pinned_free  # noqa: F401
from .GridTypes import GridPatch, GridData  # noqa: F401
from . import ParticleMesh, cosmology, conf, NoName  # noqa: F401
from .NoName import run_noname  # noqa: F401
