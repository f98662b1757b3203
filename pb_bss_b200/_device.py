"""torch tensors as device-memory containers for the C ABI.

Nothing here computes: tensors are allocated, copied host<->device and handed
to the library as raw pointers on torch's current CUDA stream.
"""
import threading

import numpy as np
import torch

from . import _lib

_scope = threading.local()


class deferred_status:
    """``with deferred_status():`` -- inside the block the device status words of the library calls (non-finite
    covariances, GEV / eigensolver failures, singular systems) are not read back after every call, because each read
    synchronises the stream and exposes the launch overhead of everything that follows; they are all read when the
    block ends, in call order, and the first failing call raises what it would have raised on the spot.  Later calls of
    the block may then have run on that call's (NaN) output.  For pipelines of CUDA tensors (parallel.py)."""

    def __enter__(self):
        self.items = []
        self.prev = getattr(_scope, 'cur', None)
        _scope.cur = self
        return self

    def __exit__(self, exc_type, exc, tb):
        _scope.cur = self.prev
        if exc_type is None:
            self.check()
        return False

    def check(self):
        items, self.items = self.items, []
        for status, on_error in items:
            s = int(status.item())
            if s:
                on_error(s)


def check_status(status, on_error):
    """Read a device status word now (synchronises) or, inside ``deferred_status``, at the end of the block;
    ``on_error(s)`` raises the call's exception for a non-zero status ``s``."""
    cur = getattr(_scope, 'cur', None)
    if cur is not None:
        cur.items.append((status, on_error))
        return
    s = int(status.item())
    if s:
        on_error(s)


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError(
            'pb_bss_b200 needs a CUDA device (B200, sm_100a); there is no CPU '
            'fallback. torch.cuda.is_available() is False.')


def device():
    require_cuda()
    return torch.device('cuda', torch.cuda.current_device())


def is_tensor(x):
    return isinstance(x, torch.Tensor)


def to_device(x, dtype=None, keep_pinned=False):
    """numpy array / torch tensor -> contiguous CUDA tensor (copy only if needed).

    ``keep_pinned``: a contiguous CPU tensor in pinned (page-locked) memory is returned as
    it is -- the library reads such buffers in place over PCIe (include/pbb.h,
    pbb_cacgmm_fit), which overlaps the upload with the computation.
    """
    dev = device()
    if (keep_pinned and is_tensor(x) and x.device.type == 'cpu' and x.is_pinned() and x.is_contiguous()
            and (dtype is None or x.dtype == dtype)):
        return x
    if not is_tensor(x):
        x = np.asarray(x)
        if not x.flags.c_contiguous:
            x = np.ascontiguousarray(x)
        if not x.flags.writeable:
            x = x.copy()
        x = torch.from_numpy(x)
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    if x.device != dev:
        x = x.to(dev, non_blocking=True)
    return x.contiguous()


def complex_dtype_code(t):
    if t.dtype == torch.complex128:
        return _lib.PBB_C128
    if t.dtype == torch.complex64:
        return _lib.PBB_C64
    raise AssertionError(f'complex input required, got {t.dtype}')


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def empty(shape, dtype):
    return torch.empty(shape, dtype=dtype, device=device())


def to_host(t, like_numpy):
    """Returns numpy if the caller passed numpy, else the tensor itself."""
    if like_numpy:
        return t.cpu().numpy()
    return t


_workspaces = {}


def workspace(nbytes):
    """A cached uint8 scratch tensor of at least nbytes on the current device
    (the caller of the C ABI owns all memory, include/pbb.h)."""
    dev = device()
    key = (dev.index, torch.cuda.current_stream().cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        _workspaces[key] = ws
    return ws
