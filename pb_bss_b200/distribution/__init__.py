"""Mixture-model trainers of the hot path (pb_bss/distribution/__init__.py)."""
from .complex_angular_central_gaussian import (  # noqa: F401
    ComplexAngularCentralGaussian,
    normalize_observation,
)
from .cacgmm import CACGMM, CACGMMTrainer  # noqa: F401
from .complex_watson import ComplexWatson, ComplexWatsonTrainer  # noqa: F401
from .cwmm import CWMM, CWMMTrainer  # noqa: F401
from .gaussian import DiagonalGaussian, SphericalGaussian  # noqa: F401
from .gcacgmm import GCACGMM, GCACGMMTrainer  # noqa: F401
from .von_mises_fisher import VonMisesFisher  # noqa: F401
from .vmfcacgmm import VMFCACGMM, VMFCACGMMTrainer  # noqa: F401
