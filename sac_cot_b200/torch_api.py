"""Tensor-native front of the SAC-COT hot path: torch tensors in, torch tensors out, no new
arithmetic — every call lands in `sac_cot_register_packed` of the C ABI (include/sac_cot.h).

  * CUDA tensors: the work is enqueued on torch's CURRENT stream of the tensors' device
    (SAC_COT_LOC_DEVICE: no host synchronisation, results are ready when the stream is); device and stream are
    inferred from the inputs, one library context is kept per (device, stream).
  * CPU tensors: the host-buffer path (H2D, pipeline, D2H and one stream sync inside the call) on `device`.

The reference has no binding to mirror (/root/reference/README.md:1-2 is the whole repository).  torch is used for
device memory and streams only; the product is the CUDA library, and without it (or without a B200) every call
raises — there is no eager/PyTorch fallback.
"""
from __future__ import annotations

import threading

import numpy as np

from . import _abi
from .api import Registrar, SacCotError, load_library

_lock = threading.Lock()
_registrars: dict[tuple[int, int], Registrar] = {}


def _registrar(device_index: int, stream_handle: int) -> Registrar:
    key = (device_index, stream_handle)
    with _lock:
        reg = _registrars.get(key)
        if reg is None:
            # handle < 0: CPU-tensor calls, a private stream of the library
            reg = Registrar(lib=load_library(), device=device_index, stream=None if stream_handle < 0 else stream_handle)
            _registrars[key] = reg
        return reg


def release_contexts() -> None:
    """Destroys the cached library contexts (their workspaces go back to the driver)."""
    with _lock:
        for reg in _registrars.values():
            reg.close()
        _registrars.clear()


def _pack(src, dst):
    import torch

    if isinstance(src, torch.Tensor) and isinstance(dst, torch.Tensor):
        if src.shape != dst.shape or src.shape[-1] != 3 or src.dim() not in (2, 3):
            raise ValueError("src and dst must both be (N, 3) or (B, N, 3)")
        batched = src.dim() == 3
        B, N = (src.shape[0], src.shape[1]) if batched else (1, src.shape[0])
        offsets = np.arange(B + 1, dtype=np.int64) * N
        s, d = src.reshape(-1, 3), dst.reshape(-1, 3)
    else:  # sequences of (N_b, 3) tensors: ragged batch
        src, dst = list(src), list(dst)
        if len(src) != len(dst) or any(a.shape != b.shape or a.dim() != 2 or a.shape[1] != 3 for a, b in zip(src, dst)):
            raise ValueError("src and dst must be equally long sequences of (N_b, 3) tensors with matching shapes")
        batched = True
        offsets = np.zeros(len(src) + 1, dtype=np.int64)
        np.cumsum([a.shape[0] for a in src], out=offsets[1:])
        s = torch.cat(src) if src else torch.empty((0, 3))
        d = torch.cat(dst) if dst else torch.empty((0, 3))
    if s.device != d.device:
        raise ValueError("src and dst must live on the same device")
    s = s.to(torch.float32).contiguous()
    d = d.to(torch.float32).contiguous()
    return s, d, offsets, batched


def register(src, dst, *, tau_compat: float = 0.1, tau_inlier: float | None = None, num_edges: int = 1024,
             apex_per_edge: int = 4, score_mode: int = 0, refit: bool = True, device: int = 0):
    """SAC-COT registration of one pair (N, 3), a batch (B, N, 3) or a ragged batch (sequences of (N_b, 3)).

    Returns (R, t, inliers): (3, 3) / (3,) / 0-d int32 for a single pair, (B, 3, 3) / (B, 3) / (B,) for a batch, on
    the inputs' device.  dst ~= R @ src + t.  For CUDA inputs the call only enqueues work on the current stream;
    `last_status(tensor)` reports, after a synchronisation, whether a workspace overflow voided it (then simply call
    again: the workspace has been grown)."""
    import torch

    s, d, offsets, batched = _pack(src, dst)
    B = len(offsets) - 1
    on_gpu = s.is_cuda
    dev_index = s.device.index if on_gpu else device
    if on_gpu:
        stream = torch.cuda.current_stream(s.device).cuda_stream
        reg = _registrar(dev_index, stream)
    else:
        reg = _registrar(dev_index, -1)
    p = reg.params
    p.tau_compat = float(tau_compat)
    p.tau_inlier = float(tau_compat if tau_inlier is None else tau_inlier)
    p.num_edges, p.apex_per_edge, p.score_mode, p.refit = int(num_edges), int(apex_per_edge), int(score_mode), int(bool(refit))
    R = torch.empty((B, 3, 3), dtype=torch.float32, device=s.device)
    t = torch.empty((B, 3), dtype=torch.float32, device=s.device)
    inl = torch.empty(B, dtype=torch.int32, device=s.device)
    if B:
        reg.register_packed_ptr(s.data_ptr(), d.data_ptr(), offsets, R.data_ptr(), t.data_ptr(), inl.data_ptr(),
                                _abi.LOC_DEVICE if on_gpu else _abi.LOC_HOST)
        if on_gpu:  # the inputs must outlive the enqueued kernels even if the caller drops them right away
            s.record_stream(torch.cuda.current_stream(s.device))
            d.record_stream(torch.cuda.current_stream(s.device))
    if batched:
        return R, t, inl
    return R[0], t[0], inl[0]


def last_status(like=None, device: int = 0) -> int:
    """Deferred status of the CUDA-tensor calls made on the current stream of `like`'s device (or of `device`) since
    the previous query; synchronises with them.  0 = fine, SAC_COT_E_NOMEM (-8) = a key-pool overflow voided a call."""
    import torch

    if like is not None and like.is_cuda:
        dev = like.device
    else:
        dev = torch.device("cuda", device)
    reg = _registrar(dev.index, torch.cuda.current_stream(dev).cuda_stream)
    return reg.get("last_status")


__all__ = ["register", "last_status", "release_contexts", "SacCotError"]
