"""Beamforming side of the hot path (pb_bss/extraction)."""
from . import linalg  # noqa: F401
from .beamformer import (  # noqa: F401
    apply_beamforming_vector,
    blind_analytic_normalization,
    get_gev_vector,
    get_mvdr_vector,
    get_mvdr_vector_souden,
    get_pca_vector,
    get_power_spectral_density_matrix,
)
from .beamformer_wrapper import get_bf_vector  # noqa: F401
