"""pb_bss_b200 -- B200-native drop-in for the EM hot path of fgnt/pb_bss.

Host side: Python mirroring the reference's public API for this path
(``distribution.CACGMMTrainer/CACGMM/CWMMTrainer/CWMM``,
``extraction.get_power_spectral_density_matrix/get_mvdr_vector/get_gev_vector``,
``permutation_alignment.DHTVPermutationAlignment``).  All arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI of ``include/pbb.h``
(``libpbb.so``, bound with ctypes in ``_lib.py``); torch tensors are only the
device-memory containers.  There is no CPU fallback.
"""
from . import _lib  # noqa: F401  (import must not need a GPU)
from . import distribution  # noqa: F401
from . import extraction  # noqa: F401
from . import permutation_alignment  # noqa: F401
from . import initializer  # noqa: F401
from ._device import deferred_status  # noqa: F401

__all__ = ['distribution', 'extraction', 'permutation_alignment', 'initializer']
