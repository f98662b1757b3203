"""Mask-based beamforming on the device, signatures of pb_bss/extraction/beamformer.py.

Shapes follow the reference (beamformer.py:1-12): X (F, D, T), mask (F, K, T),
PSD (F, K, D, D); leading dims are independent.  numpy in -> numpy out, CUDA
tensors in -> CUDA tensors out.  Small matrices are complex128 on the device.
"""
import numpy as np
import torch

from .. import _device, _lib
from .linalg import eigh

__all__ = [
    'get_power_spectral_density_matrix',
    'get_pca_vector',
    'get_mvdr_vector',
    'get_mvdr_vector_souden',
    'get_gev_vector',
    'blind_analytic_normalization',
    'apply_beamforming_vector',
]


def _flat(t, inner):
    """(..., *inner dims) -> (n, *inner dims) contiguous; returns (flat, leading shape)."""
    lead = tuple(t.shape[:t.dim() - inner])
    n = int(np.prod(lead)) if lead else 1
    return t.reshape(n, *t.shape[t.dim() - inner:]).contiguous(), lead


def get_power_spectral_density_matrix(observation, mask=None, sensor_dim=-2,
                                      source_dim=-2, time_dim=-1,
                                      normalize=True):
    """Mask-weighted spatial covariance, beamformer.py:59-160.

    observation: (..., sensors, frames) by default; mask: None, (..., frames)
    or (..., sources, frames).  Returns (..., sensors, sensors) or
    (..., sources, sensors, sensors)."""
    like_numpy = not _device.is_tensor(observation)
    obs = _device.to_device(observation)
    code = _device.complex_dtype_code(obs)
    nd = obs.dim()
    sensor_dim, source_dim, time_dim = (d % nd - nd for d in (sensor_dim, source_dim, time_dim))
    order = [i for i in range(-nd, 0) if i not in (sensor_dim, time_dim)] + [sensor_dim, time_dim]
    obs = obs.permute(*[i % nd for i in order])
    obs, lead = _flat(obs, 2)
    F, D, T = obs.shape
    lib = _lib.load()
    single = False
    if mask is None:
        m, K = None, 1
        single = True
    else:
        m = _device.to_device(mask)
        m = m.to(torch.float64)
        if m.dim() + 1 == nd:
            m = m.unsqueeze(-2)
            single = True
        else:
            morder = [i for i in range(-nd, 0) if i not in (source_dim, time_dim)] + [source_dim, time_dim]
            m = m.permute(*[i % nd for i in morder])
        K = m.shape[-2]
        m = m.expand(*lead, K, T).reshape(F, K, T).contiguous()
    psd = _device.empty((F, K, D, D), torch.complex128)
    nbytes = lib.pbb_psd_workspace_bytes(F, T, D, K)
    ws = _device.workspace(nbytes)
    _lib.check(lib.pbb_power_spectral_density(
        _device.ptr(obs), code, F, D, T, _device.ptr(m), K, int(bool(normalize)),
        _device.ptr(psd), _device.ptr(ws), nbytes, _device.stream_ptr()),
        'pbb_power_spectral_density')
    if single:
        out = psd.reshape(*lead, D, D)
    else:
        out = psd.reshape(*lead, K, D, D)
        if source_dim < -2:
            # PSD shape (sources, ..., sensors, sensors), beamformer.py:156-158
            out = out.movedim(-3, source_dim % nd)
    return _device.to_host(out.contiguous(), like_numpy)


def get_pca_vector(target_psd_matrix, scaling=None):
    """Principal eigenvector of the target PSD, beamformer.py:197-224."""
    like_numpy = not _device.is_tensor(target_psd_matrix)
    psd = _device.to_device(target_psd_matrix, torch.complex128)
    w, v = eigh(psd)
    vec, val = v[..., -1], w[..., -1]
    if scaling is None:
        pass
    elif scaling == 'trace':
        tr = torch.diagonal(psd, dim1=-2, dim2=-1).sum(-1)
        vec = vec * (torch.sqrt(tr) / torch.linalg.vector_norm(vec, dim=-1))[..., None]
    elif scaling == 'eigenvalue':
        vec = vec * (val / torch.linalg.vector_norm(vec, dim=-1))[..., None]
    else:
        raise ValueError(scaling)
    return _device.to_host(vec.contiguous(), like_numpy)


def get_mvdr_vector(atf_vector, noise_psd_matrix):
    """w = N^-1 a / (a^H N^-1 a), beamformer.py:230-260.
    atf_vector (..., bins, sensors), noise_psd_matrix (bins, sensors, sensors)."""
    assert noise_psd_matrix is not None
    like_numpy = not _device.is_tensor(atf_vector)
    atf = _device.to_device(atf_vector, torch.complex128)
    noise = _device.to_device(noise_psd_matrix, torch.complex128)
    while atf.dim() > noise.dim() - 1:
        noise = noise.unsqueeze(0)
    D = atf.shape[-1]
    lead = torch.broadcast_shapes(atf.shape[:-1], noise.shape[:-2])
    atf_f = atf.expand(*lead, D).reshape(-1, D).contiguous()
    noise_f = noise.expand(*lead, D, D).reshape(-1, D, D).contiguous()
    n = atf_f.shape[0]
    w = _device.empty((n, D), torch.complex128)
    scratch = _device.empty((n, D), torch.complex128)
    status = _device.empty((1,), torch.int32)
    status.zero_()
    lib = _lib.load()
    _lib.check(lib.pbb_mvdr(_device.ptr(atf_f), _device.ptr(noise_f), n, D, _device.ptr(w),
                            _device.ptr(scratch), _device.ptr(status), _device.stream_ptr()), 'pbb_mvdr')
    # a singular noise PSD matrix takes the reference's np.linalg.lstsq fallback (beamformer.py:251-256) on the
    # device (minimum-norm solution); the status word is only set where that fallback does not exist (D > 40)
    def on_error(s):
        raise np.linalg.LinAlgError(f'get_mvdr_vector: singular noise PSD matrix {s - 1} (D > 40: no lstsq fallback)')
    _device.check_status(status, on_error)
    return _device.to_host(w.reshape(*lead, D), like_numpy)


def get_gev_vector(target_psd_matrix, noise_psd_matrix, force_cython=False,
                   use_eig=False):
    """Generalised-eigenvalue beamformer, beamformer.py:292-364: eigenvector of
    the largest eigenvalue of (target, noise), LAPACK ``zhegvd`` normalisation.
    ``force_cython`` is accepted for signature parity (there is only the device
    path here); ``use_eig`` (non-Hermitian ``zggev``) is not implemented."""
    assert noise_psd_matrix is not None
    if use_eig:
        raise NotImplementedError('use_eig=True (general eig) is outside the hot path (SURVEY.md section 2)')
    like_numpy = not _device.is_tensor(target_psd_matrix)
    a = _device.to_device(target_psd_matrix, torch.complex128)
    b = _device.to_device(noise_psd_matrix, torch.complex128)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert a.shape[-1] == a.shape[-2], a.shape
    D = a.shape[-1]
    af, lead = _flat(a, 2)
    bf, _ = _flat(b, 2)
    n = af.shape[0]
    w = _device.empty((n, D), torch.complex128)
    status = _device.empty((1,), torch.int32)
    status.zero_()
    lib = _lib.load()
    _lib.check(lib.pbb_gev_batched(_device.ptr(af), _device.ptr(bf), n, D, _device.ptr(w),
                                   _device.ptr(status), _device.stream_ptr()), 'pbb_gev_batched')
    def on_error(s):
        # get_gev_vector.pyx:130-147 / beamformer.py:398-408
        raise ValueError(f'Error for frequency {s - 1}: noise PSD not positive definite or non-finite input')
    _device.check_status(status, on_error)
    return _device.to_host(w.reshape(*lead, D), like_numpy)


def get_mvdr_vector_souden(target_psd_matrix, noise_psd_matrix, ref_channel=None,
                           eps=None, return_ref_channel=False):
    """Souden MVDR, beamformer.py:627-698 (+ get_optimal_reference_channel :601-624)."""
    assert noise_psd_matrix is not None
    like_numpy = not _device.is_tensor(target_psd_matrix)
    if isinstance(target_psd_matrix, (list, tuple)):
        target_psd_matrix = np.asarray(target_psd_matrix)
    if isinstance(noise_psd_matrix, (list, tuple)):
        noise_psd_matrix = np.asarray(noise_psd_matrix)
    # real PSD matrices give a real vector in the reference (NumPy keeps the dtype); the device path is complex
    real_in = like_numpy and not np.iscomplexobj(target_psd_matrix) and not np.iscomplexobj(noise_psd_matrix)
    t = _device.to_device(np.asarray(target_psd_matrix) if isinstance(target_psd_matrix, (list, tuple))
                          else target_psd_matrix, torch.complex128)
    nz = _device.to_device(np.asarray(noise_psd_matrix) if isinstance(noise_psd_matrix, (list, tuple))
                           else noise_psd_matrix, torch.complex128)
    D = t.shape[-1]
    tf, lead = _flat(t, 2)
    nf, _ = _flat(nz.expand_as(t), 2)
    n = tf.shape[0]
    if eps is None:
        eps = np.finfo(np.float64).tiny
    lib = _lib.load()
    phi = _device.empty((n, D, D), torch.complex128)
    status = _device.empty((1,), torch.int32)
    status.zero_()
    _lib.check(lib.pbb_solve_batched(_device.ptr(nf), _device.ptr(tf), n, D, D, 0, _device.ptr(phi),
                                     _device.ptr(status), _device.stream_ptr()), 'pbb_solve_batched')
    # stable_solve (math/solve.py:95-114): singular systems get the minimum-norm (lstsq) solution on the device
    def on_error(s):
        raise np.linalg.LinAlgError(f'get_mvdr_vector_souden: singular noise PSD matrix {s - 1} (D > 40: no lstsq fallback)')
    _device.check_status(status, on_error)
    mat = _device.empty((n, D, D), torch.complex128)
    num = _device.empty((n, D), torch.complex128)
    den = _device.empty((n, D), torch.complex128)
    nsum = _device.empty((D,), torch.complex128)
    dsum = _device.empty((D,), torch.complex128)
    _lib.check(lib.pbb_souden(_device.ptr(phi), _device.ptr(tf), _device.ptr(nf), n, D, float(eps),
                              _device.ptr(mat), _device.ptr(num), _device.ptr(den), _device.ptr(nsum),
                              _device.ptr(dsum), _device.stream_ptr()), 'pbb_souden')
    if ref_channel is None:
        if len(lead) != 1:
            raise ValueError(
                'Estimating the ref_channel expects currently that the input '
                'has 3 ndims (frequency x sensors x sensors). '
                'Considering an independent dim in the SNR estimate is not unique.')
        ns, ds = nsum.cpu().numpy(), dsum.cpu().numpy()
        snr = ns / np.maximum(ds, eps)
        assert np.all(np.isfinite(snr)), snr
        ref_channel = int(np.argmax(snr.real))
    assert np.isscalar(ref_channel), ref_channel
    beamformer = _device.to_host(mat[..., ref_channel].reshape(*lead, D).contiguous(), like_numpy)
    if real_in:
        beamformer = np.ascontiguousarray(beamformer.real)
    return (beamformer, ref_channel) if return_ref_channel else beamformer


def blind_analytic_normalization(vector, noise_psd_matrix):
    """beamformer.py:459-488."""
    like_numpy = not _device.is_tensor(vector)
    v = _device.to_device(vector, torch.complex128)
    nz = _device.to_device(noise_psd_matrix, torch.complex128)
    D = v.shape[-1]
    lead = torch.broadcast_shapes(v.shape[:-1], nz.shape[:-2])
    vf = v.expand(*lead, D).reshape(-1, D).contiguous()
    nf = nz.expand(*lead, D, D).reshape(-1, D, D).contiguous()
    out = _device.empty(vf.shape, torch.complex128)
    lib = _lib.load()
    _lib.check(lib.pbb_blind_analytic_normalization(_device.ptr(vf), _device.ptr(nf), vf.shape[0], D,
                                                    _device.ptr(out), _device.stream_ptr()),
               'pbb_blind_analytic_normalization')
    return _device.to_host(out.reshape(*lead, D), like_numpy)


def apply_beamforming_vector(vector, mix):
    """out[..., t] = sum_a conj(vector[..., a]) mix[..., a, t], beamformer.py:572-583."""
    like_numpy = not _device.is_tensor(mix)
    v = _device.to_device(vector, torch.complex128)
    y = _device.to_device(mix)
    code = _device.complex_dtype_code(y)
    assert v.shape[-1] < 30, (v.shape, y.shape)
    D, T = y.shape[-2], y.shape[-1]
    lead = tuple(torch.broadcast_shapes(v.shape[:-1], y.shape[:-2]))
    ye = y.expand(*lead, D, T)
    # leading dims over which the mix is only broadcast (K beamformers applied to ONE STFT): the kernel reads the one
    # mix for every index instead of a materialised copy per index
    bdims = [i for i in range(len(lead)) if lead[i] > 1 and ye.stride(i) == 0]
    if bdims:
        nb = len(bdims)
        others = [i for i in range(len(lead)) if i not in bdims]
        bshape, oshape = [lead[i] for i in bdims], [lead[i] for i in others]
        B, F = int(np.prod(bshape)), int(np.prod(oshape)) if others else 1
        ve = v.expand(*lead, D).permute(*bdims, *others, len(lead)).reshape(B, F, D).contiguous()
        ysmall = ye[tuple(0 if i in bdims else slice(None) for i in range(len(lead)))].reshape(F, D, T).contiguous()
        out = _device.empty((B, F, T), torch.complex128)
        lib = _lib.load()
        _lib.check(lib.pbb_apply_beamforming_vector_shared(_device.ptr(ve), _device.ptr(ysmall), code, B, F, D, T,
                                                           _device.ptr(out), _device.stream_ptr()),
                   'pbb_apply_beamforming_vector_shared')
        out = out.reshape(*bshape, *oshape, T)
        if bdims != list(range(nb)):
            out = out.permute(*np.argsort(bdims + others).tolist(), len(lead)).contiguous()
        return _device.to_host(out, like_numpy)
    vf = v.expand(*lead, D).reshape(-1, D).contiguous()
    yf = ye.reshape(-1, D, T).contiguous()
    F = vf.shape[0]
    out = _device.empty((F, T), torch.complex128)
    lib = _lib.load()
    _lib.check(lib.pbb_apply_beamforming_vector(_device.ptr(vf), _device.ptr(yf), code, F, D, T,
                                                _device.ptr(out), _device.stream_ptr()),
               'pbb_apply_beamforming_vector')
    return _device.to_host(out.reshape(*lead, T), like_numpy)
