"""Import the unmodified reference from /root/reference (build container only).

The reference's top-level ``pb_bss/__init__.py`` pulls in ``paderbox`` (absent
here), so a stub package whose ``__path__`` points at the reference is
registered first, plus a ``cached_property`` stub.
Used only by ``oracle/make_golden.py`` and the optional ``live reference``
tests; it is NOT available on the GPU box (no /root/reference there).
"""
import functools
import os
import sys
import types
import warnings

REF = os.environ.get('PB_BSS_REFERENCE', '/root/reference')
if not os.path.isdir(os.path.join(REF, 'pb_bss')):
    # the GPU box has no checkout: the verbatim copy made by oracle/build_ref.py travels with the repository
    _vendored = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
    if os.path.isdir(os.path.join(_vendored, 'pb_bss')):
        REF = _vendored


def available():
    return os.path.isdir(os.path.join(REF, 'pb_bss'))


def load():
    """Returns a namespace with the reference modules of the hot path."""
    if not available():
        raise RuntimeError(f'reference checkout not found at {REF}')
    sys.dont_write_bytecode = True
    if 'pb_bss' not in sys.modules:
        pkg = types.ModuleType('pb_bss')
        pkg.__path__ = [os.path.join(REF, 'pb_bss')]
        sys.modules['pb_bss'] = pkg
    if 'cached_property' not in sys.modules:
        cp = types.ModuleType('cached_property')
        cp.cached_property = functools.cached_property
        sys.modules['cached_property'] = cp
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import pb_bss.distribution as distribution
        import pb_bss.distribution.complex_watson as complex_watson
        import pb_bss.distribution.complex_angular_central_gaussian as cacg
        import pb_bss.distribution.mixture_model_utils as mixture_model_utils
        import pb_bss.extraction.beamformer as beamformer
        import pb_bss.permutation_alignment as permutation_alignment
    return types.SimpleNamespace(
        distribution=distribution, complex_watson=complex_watson, cacg=cacg,
        mixture_model_utils=mixture_model_utils, beamformer=beamformer,
        permutation_alignment=permutation_alignment)
