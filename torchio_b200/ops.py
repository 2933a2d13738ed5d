"""Functional, tensor-level API over the C-ABI kernels.

Each function is the plain-tensor replacement of one reference helper
(cited per function; TorchIO 2.0.0a2, paths relative to
src/torchio/transforms/).  Inputs are CUDA tensors; parameter tables are
uploaded with one pinned staging copy per call (`upload`).  Launches go to the
caller's current CUDA stream on the tensor's device.  No CPU fallback.
"""

from __future__ import annotations

import os
import threading

import numpy as np
import torch
from torch import Tensor

from . import _native

DTYPE_CODES = {
    torch.float32: 0, torch.uint8: 1, torch.int8: 2,
    torch.int16: 3, torch.int32: 4, torch.int64: 5,
}
NEAREST, LINEAR, LABEL_PV = 0, 1, 2
FLAG_PASSTHROUGH, FLAG_ELASTIC = 1, 2

_counter = threading.local()


def launches() -> int:
    """Number of kernels this thread has launched through the library."""
    return getattr(_counter, "n", 0)


def _count(n: int) -> None:
    _counter.n = getattr(_counter, "n", 0) + n


def _stream(t: Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t: Tensor | None):
    return None if t is None else t.data_ptr()


def _require_cuda(t: Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"torchio_b200.ops.{name}: expected a CUDA tensor (got {t.device});"
            " the kernels have no CPU fallback"
        )
    if t.requires_grad:
        raise NotImplementedError(
            f"torchio_b200.ops.{name}: kernels are forward-only; detach() the input"
        )


class _StagingRing:
    """Fixed ring of pinned staging buffers for the per-call parameter tables.

    Allocating pinned memory per call (``torch.empty(pin_memory=True)``) looked
    free but is not: while the host runs ahead of the GPU the caching host
    allocator cannot recycle blocks whose copies are still queued, so every call
    ends in ``cudaHostAlloc`` — measured at ~1.2 ms of GPU stall per upload on
    B200.  The ring allocates once; a slot is reused only after the event
    recorded behind its last copy has completed."""

    SLOTS = 64
    SLOT_BYTES = 1 << 20

    def __init__(self) -> None:
        self.buffers = [torch.empty(self.SLOT_BYTES, dtype=torch.uint8, pin_memory=True)
                        for _ in range(self.SLOTS)]
        self.events: list = [None] * self.SLOTS
        self.index = 0
        self.lock = threading.Lock()

    def take(self):
        with self.lock:
            i = self.index
            self.index = (i + 1) % self.SLOTS
            pending, self.events[i] = self.events[i], None
        if pending is not None:  # the copy queued behind this slot's previous use
            pending.synchronize()
        return i, self.buffers[i]

    def release(self, i: int, device: torch.device) -> None:
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        with self.lock:
            self.events[i] = ev


_ring: _StagingRing | None = None
_init_lock = threading.Lock()


def upload(device: torch.device, *arrays):
    """Pack host arrays into one pinned buffer, copy once, return device views.

    Each array is a numpy array or CPU tensor (or None -> None).  16-byte
    aligned segments so float4/TMA consumers can read them directly.
    """
    global _ring
    specs, offset = [], 0
    for a in arrays:
        if a is None:
            specs.append(None)
            continue
        t = torch.as_tensor(a)
        t = t.contiguous()
        nbytes = t.numel() * t.element_size()
        specs.append((t, offset, nbytes))
        offset += (nbytes + 15) // 16 * 16
    if offset == 0:
        return [None] * len(arrays)
    slot = None
    if torch.cuda.is_available() and offset <= _StagingRing.SLOT_BYTES:
        if _ring is None:
            with _init_lock:
                if _ring is None:
                    _ring = _StagingRing()
        slot, stage = _ring.take()
        stage = stage[:offset]
    else:
        stage = torch.empty(offset, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    # plain memcpy through numpy views: torch's CPU copy_ fans tensors above 32 KiB out to
    # the intra-op thread pool, and waking 100+ OpenMP threads on a busy host was measured
    # to stall this call for tens of milliseconds (tools/host_stalls3.py)
    stage_np = stage.numpy()
    for s in specs:
        if s is None:
            continue
        t, off, nbytes = s
        stage_np[off:off + nbytes] = t.reshape(-1).numpy().view(np.uint8)
    if slot is not None:
        # SM copy kernel instead of cudaMemcpyAsync: see tio_upload in include/tio_b200.h
        dev = torch.empty(offset, dtype=torch.uint8, device=device)
        _native.call("tio_upload", stage.data_ptr(), dev.data_ptr(), offset,
                     torch.cuda.current_stream(dev.device).cuda_stream)
        _count(1)
        _ring.release(slot, torch.device(device))
    else:
        dev = stage.to(device, non_blocking=True)
    out = []
    for s in specs:
        if s is None:
            out.append(None)
            continue
        t, off, nbytes = s
        out.append(dev[off:off + nbytes].view(t.dtype).reshape(t.shape))
    return out


EXACT_COORDS = 0x100  # TIO_EXACT_COORDS
_exact_default = os.environ.get("TIO_B200_EXACT_COORDS", "0") not in ("", "0")


def set_exact_coords(enabled: bool) -> bool:
    """Process-wide default of ``resample(exact_coords=None)``; returns the previous value.
    Also settable with ``TIO_B200_EXACT_COORDS=1`` before import."""
    global _exact_default
    previous, _exact_default = _exact_default, bool(enabled)
    return previous


def exact_coords_default() -> bool:
    return _exact_default


def resample(
    src: Tensor, mat: Tensor, cp: Tensor | None, flags: Tensor | None,
    spacing_in, spacing_out, *, affine_first: bool, mode: int,
    fill: Tensor | None, out_shape=None, box_hint: int = 0, exact_coords: bool | None = None,
) -> Tensor:
    """K1.  Replaces _build_sampling_grid + _sample_batch[_per_sample]
    (spatial/spatial.py:1504-1579,1651-1857).

    src (B,C,I,J,K) fp32 or integer label dtype; mat (B,12) fp32 cuda;
    cp (B,ni,nj,nk,3) fp32 cuda or None; flags (B,) uint8 cuda or None;
    fill (C,) fp32 cuda or None (= no mask step).

    exact_coords: keep the reference's fp32 rounding sequence of the sampling coordinates on
    every voxel (TIO_EXACT_COORDS).  Default (None -> ``exact_coords_default()``): fp32
    trilinear voxels whose taps all lie inside the volume use one fma per axis, which differs
    from the reference by its own coordinate noise (<= ~2e-5 voxel); label maps, padding and
    fill decisions are exact either way.
    """
    _require_cuda(src, "resample")
    if src.dtype not in DTYPE_CODES:
        raise TypeError(f"resample: unsupported dtype {src.dtype}")
    src = src.contiguous()
    b, c, i, j, k = src.shape
    oi, oj, ok = (i, j, k) if out_shape is None else out_shape
    dst = torch.empty((b, c, oi, oj, ok), dtype=src.dtype, device=src.device)
    ni = nj = nk = 0
    if cp is not None:
        ni, nj, nk = cp.shape[1:4]
    sp_in = np.asarray(spacing_in, dtype=np.float32)
    sp_out = np.asarray(spacing_out, dtype=np.float32)
    workspace, ws_bytes = None, 0
    tiled = (src.dtype == torch.float32 and mode == LINEAR) or (
        mode in (NEAREST, LABEL_PV) and src.dtype in (torch.uint8, torch.int16, torch.int32))
    if box_hint >= 0 and tiled:
        ws_bytes = _native.lib().tio_resample_workspace_bytes(b, oi, oj, ok)
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=src.device)
    with torch.cuda.device(src.device):
        _native.call(
            "tio_resample", _ptr(src), _ptr(dst), DTYPE_CODES[src.dtype],
            b, c, i, j, k, oi, oj, ok, _ptr(mat), _ptr(cp), _ptr(flags), ni, nj, nk,
            sp_in.ctypes.data, sp_out.ctypes.data, int(bool(affine_first)),
            int(mode) | (EXACT_COORDS if (_exact_default if exact_coords is None else exact_coords) else 0),
            _ptr(fill), int(box_hint), _ptr(workspace), ws_bytes, _stream(src),
        )
    _count(2 if workspace is not None else 1)
    return dst


def _label_table(labels: Tensor, dtype: torch.dtype) -> Tensor:
    return labels.to(torch.float32 if dtype == torch.float32 else torch.int64).contiguous()


def onehot(src: Tensor, labels: Tensor) -> Tensor:
    """(B,1,I,J,K) label batch -> (B,n,I,J,K) fp32 one-hot channels, ``labels`` = the distinct
    values in ascending order (spatial/spatial.py:1362-1365)."""
    _require_cuda(src, "onehot")
    if src.dtype not in DTYPE_CODES:
        raise TypeError(f"onehot: unsupported dtype {src.dtype}")
    src = src.contiguous()
    b, n = src.shape[0], int(labels.numel())
    vox = src[0].numel()
    table = _label_table(labels, src.dtype)
    dst = torch.empty((b, n, *src.shape[2:]), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        _native.call("tio_onehot", _ptr(src), DTYPE_CODES[src.dtype], b, vox, _ptr(table), n, _ptr(dst),
                     _stream(src))
    _count(1)
    return dst


def label_argmax(sampled: Tensor, labels: Tensor, pad_label: float, dtype: torch.dtype) -> Tensor:
    """(B,n,I,J,K) sampled one-hot channels -> (B,1,I,J,K) labels of ``dtype``: first maximum over
    the channels, ``pad_label`` where their sum is not > 0.5 (spatial/spatial.py:1378-1389)."""
    _require_cuda(sampled, "label_argmax")
    if dtype not in DTYPE_CODES:
        raise TypeError(f"label_argmax: unsupported dtype {dtype}")
    sampled = sampled.contiguous()
    b, n = sampled.shape[:2]
    vox = sampled[0, 0].numel()
    table = _label_table(labels, dtype)
    dst = torch.empty((b, 1, *sampled.shape[2:]), dtype=dtype, device=sampled.device)
    with torch.cuda.device(sampled.device):
        _native.call("tio_label_argmax", _ptr(sampled), b, n, vox, _ptr(table), float(pad_label), _ptr(dst),
                     DTYPE_CODES[dtype], _stream(sampled))
    _count(1)
    return dst


def min_sample0(src: Tensor) -> Tensor:
    """Per-channel min of batch element 0, on device, no sync
    (spatial/spatial.py:2054-2060,2094-2095)."""
    _require_cuda(src, "min_sample0")
    src = src.contiguous()
    c = src.shape[1]
    n = src[0, 0].numel()
    fill = torch.empty(c, dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        _native.call("tio_min_sample0", _ptr(src), c, n, _ptr(fill), _stream(src))
    _count(2)
    return fill


def crop_patches(volume: Tensor, corners, size, out: Tensor | None = None) -> Tensor:
    """Gather ``n`` patches of ``size`` at voxel ``corners`` (n,3) from one volume
    (C,I,J,K) into a dense (n,C,*size) block in a single launch
    (data/sampler.py:54-67 + loader.py:15-24).  ``out``: a contiguous (n,C,*size) block to
    write into (e.g. consecutive slots of a patch ring) instead of a fresh tensor."""
    _require_cuda(volume, "crop_patches")
    if volume.ndim != 4:
        raise ValueError(f"crop_patches expects a (C, I, J, K) volume, got {tuple(volume.shape)}")
    volume = volume.contiguous()
    corners = np.ascontiguousarray(np.asarray(corners, dtype=np.int32).reshape(-1, 3))
    n = corners.shape[0]
    c, i, j, k = (int(v) for v in volume.shape)
    pi, pj, pk = (int(v) for v in size)
    if n == 0:
        return volume.new_empty((0, c, pi, pj, pk))
    if corners.min() < 0 or np.any(corners + np.asarray([pi, pj, pk]) > np.asarray([i, j, k])):
        raise ValueError("crop_patches: a patch extends beyond the volume")
    (corners_d,) = upload(volume.device, corners)
    if out is None:
        dst = torch.empty((n, c, pi, pj, pk), dtype=volume.dtype, device=volume.device)
    else:
        dst = out
        if (tuple(dst.shape) != (n, c, pi, pj, pk) or dst.dtype != volume.dtype or dst.device != volume.device
                or not dst.is_contiguous()):
            raise ValueError("crop_patches: `out` must be a contiguous (n, C, *size) block of the volume's dtype/device")
    with torch.cuda.device(volume.device):
        _native.call(
            "tio_crop_patches", _ptr(volume), _ptr(dst), volume.element_size(), c, i, j, k, n,
            _ptr(corners_d), pi, pj, pk, _stream(volume),
        )
    _count(1)
    return dst


PAD_MODES = {"constant": 0, "replicate": 1, "reflect": 2, "circular": 3}


def remap(src: Tensor, out_shape, offsets, *, mode: str = "constant", fill=0, flip: Tensor | None = None,
          out: Tensor | None = None) -> Tensor:
    """Flip / Crop / Pad in one pass: ``out[..., o] = src[..., o - offset]`` per spatial
    axis with F.pad's out-of-range rules, then per-element axis reversal (``flip``: (B,)
    uint8 cuda, bit 0 I, 1 J, 2 K).  flip.py:233-263, crop.py:84-101, _padding.py:73-104."""
    _require_cuda(src, "remap")
    if src.ndim != 5:
        raise ValueError(f"remap expects (B, C, I, J, K), got {tuple(src.shape)}")
    src = src.contiguous()
    b, c, i, j, k = (int(v) for v in src.shape)
    oi, oj, ok = (int(v) for v in out_shape)
    if out is None:
        dst = torch.empty((b, c, oi, oj, ok), dtype=src.dtype, device=src.device)
    else:
        dst = out
        if (tuple(dst.shape) != (b, c, oi, oj, ok) or dst.dtype != src.dtype or dst.device != src.device
                or not dst.is_contiguous()):
            raise ValueError("remap: `out` must be a contiguous (B, C, *out_shape) block of the source's dtype/device")
    fill_host = torch.tensor([fill]).to(src.dtype)  # F.pad casts the value to the tensor's dtype
    with torch.cuda.device(src.device):
        _native.call(
            "tio_remap", _ptr(src), _ptr(dst), src.element_size(), b, c, i, j, k, oi, oj, ok,
            int(offsets[0]), int(offsets[1]), int(offsets[2]), PAD_MODES[mode],
            fill_host.data_ptr(), _ptr(flip), _stream(src),
        )
    _count(1)
    return dst


def bias_field(src: Tensor, coarse: Tensor, identity: Tensor | None, *, divide=False,
               out: Tensor | None = None) -> Tensor:
    """K2 (intensity/bias_field.py:201-255,296-341)."""
    _require_cuda(src, "bias_field")
    src = src.contiguous()
    b, c, i, j, k = src.shape
    dst = torch.empty_like(src) if out is None else out
    si, sj, sk = coarse.shape[2:]
    with torch.cuda.device(src.device):
        _native.call(
            "tio_bias_field", _ptr(src), _ptr(dst), b, c, i, j, k, _ptr(coarse), si, sj, sk,
            _ptr(identity), int(bool(divide)), _stream(src),
        )
    _count(1)
    return dst


def blur(src: Tensor, taps: Tensor, radius: Tensor, big_r: int, axes_mask: int,
         identity: Tensor | None) -> Tensor:
    """K3 (intensity/blur.py:129-252)."""
    _require_cuda(src, "blur")
    src = src.contiguous()
    b, c, i, j, k = src.shape
    dst = torch.empty_like(src)
    scratch = torch.empty_like(src) if (axes_mask & 6) else None
    with torch.cuda.device(src.device):
        _native.call(
            "tio_blur", _ptr(src), _ptr(dst), _ptr(scratch), b, c, i, j, k, _ptr(taps),
            _ptr(radius), int(big_r), int(axes_mask), _ptr(identity), _stream(src),
        )
    _count(_fused_launches(axes_mask, False))
    return dst


def _fused_launches(axes_mask: int, has_bias: bool) -> int:
    jk = bool(axes_mask & 6)
    march = (not jk) or bool(axes_mask & 1) or has_bias
    return int(jk) + int(march)


def noise(src: Tensor, mean: Tensor, std: Tensor, keep: Tensor | None, z: Tensor,
          z2: Tensor | None = None) -> Tensor:
    """K4 with caller-provided normals (intensity/noise.py:98-178)."""
    _require_cuda(src, "noise")
    src = src.contiguous()
    dst = torch.empty_like(src)
    with torch.cuda.device(src.device):
        _native.call(
            "tio_noise", _ptr(src), _ptr(dst), src.shape[0], src[0].numel(), _ptr(mean),
            _ptr(std), _ptr(keep), _ptr(z), _ptr(z2), _stream(src),
        )
    _count(1)
    return dst


def noise_philox(src: Tensor, mean: Tensor, std: Tensor, keep: Tensor | None, seed: int,
                 rician: bool = False) -> Tensor:
    """K4b: in-register Philox normals (not the reference stream)."""
    _require_cuda(src, "noise_philox")
    src = src.contiguous()
    dst = torch.empty_like(src)
    with torch.cuda.device(src.device):
        _native.call(
            "tio_noise_philox", _ptr(src), _ptr(dst), src.shape[0], src[0].numel(),
            _ptr(mean), _ptr(std), _ptr(keep), int(seed), int(bool(rician)), _stream(src),
        )
    _count(1)
    return dst


def gamma(src: Tensor, gam: Tensor) -> Tensor:
    """K5 (intensity/gamma.py:88-90)."""
    _require_cuda(src, "gamma")
    src = src.contiguous()
    dst = torch.empty_like(src)
    with torch.cuda.device(src.device):
        _native.call(
            "tio_gamma", _ptr(src), _ptr(dst), src.shape[0], src[0].numel(), _ptr(gam),
            _stream(src),
        )
    _count(1)
    return dst


def moments(values: Tensor, mask: Tensor | None = None) -> tuple[float, float, float]:
    """(sum, sum of squares, count) of the selected values of a contiguous fp32 CUDA tensor, fp64
    accumulation on the device, one small D2H read (the reference calls ``.item()`` here too:
    standardize.py:76-77)."""
    _require_cuda(values, "moments")
    values = values.contiguous()
    m8 = None if mask is None else mask.expand_as(values).contiguous().to(torch.uint8)
    out = torch.empty(3, dtype=torch.float64, device=values.device)
    with torch.cuda.device(values.device):
        _native.call("tio_moments", _ptr(values), _ptr(m8), values.numel(), _ptr(out), _stream(values))
    _count(1)
    s, ss, n = out.tolist()
    return s, ss, n


def quantile_neighbours(values: Tensor, qs, mask: Tensor | None = None):
    """For each q in ``qs`` (at most two): the order statistics torch.kthvalue(lower + 1) and
    kthvalue(lower + 2) return, and ``index - lower`` (transforms/_statistics.py:37-45), found by
    an exact radix select on the device.  Returns (values[2m], weights[m], count)."""
    _require_cuda(values, "quantile_neighbours")
    values = values.contiguous()
    m8 = None if mask is None else mask.expand_as(values).contiguous().to(torch.uint8)
    qs = np.ascontiguousarray(np.asarray(qs, dtype=np.float64).reshape(-1))
    m = int(qs.shape[0])
    dev = values.device
    vals = torch.empty(2 * m, dtype=torch.float32, device=dev)
    out = torch.empty(m + 1, dtype=torch.float64, device=dev)
    ws_bytes = _native.lib().tio_quantiles_workspace_bytes()
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _native.call("tio_quantiles", _ptr(values), _ptr(m8), values.numel(), qs.ctypes.data, m, _ptr(vals),
                     _ptr(out), out[m:].data_ptr(), _ptr(ws), ws_bytes, _stream(values))
    _count(8)
    host = out.tolist()
    return vals.tolist(), host[:m], int(host[m])


def rescale(src: Tensor, *, lo: float | None = None, hi: float | None = None, sub=None, div=None, mul=None,
            add=None, keep=None) -> Tensor:
    """``((clamp(src, lo, hi) - sub[b]) / div[b]) * mul[b] + add[b]`` over a (B, ...) fp32 batch, each
    step rounded like the reference's separate elementwise ops; omitted steps are skipped.
    ``sub/div/mul/add``: scalars or length-B sequences; ``keep``: length-B, 0 = copy the row."""
    _require_cuda(src, "rescale")
    src = src.contiguous()
    b = src.shape[0]
    flags = (1 if lo is not None else 0)
    tabs = []
    for bit, table in ((2, sub), (4, div), (8, mul), (16, add)):
        if table is None:
            tabs.append(None)
            continue
        flags |= bit
        arr = np.asarray(table, dtype=np.float32).reshape(-1)
        tabs.append(np.ascontiguousarray(np.broadcast_to(arr, (b,)) if arr.size == 1 else arr))
    keep_np = None if keep is None else np.ascontiguousarray(np.asarray(keep, dtype=np.uint8))
    sub_d, div_d, mul_d, add_d, keep_d = upload(src.device, *tabs, keep_np)
    dst = torch.empty_like(src)
    with torch.cuda.device(src.device):
        _native.call("tio_rescale", _ptr(src), _ptr(dst), b, src[0].numel(),
                     float(lo if lo is not None else 0.0), float(hi if hi is not None else 0.0),
                     _ptr(sub_d), _ptr(div_d), _ptr(mul_d), _ptr(add_d), _ptr(keep_d), flags, _stream(src))
    _count(2)
    return dst


def intensity_fused(
    src: Tensor, *, coarse: Tensor | None = None, bias_identity: Tensor | None = None,
    bias_divide: bool = False, taps: Tensor | None = None, radius: Tensor | None = None,
    big_r: int = 0, axes_mask: int = 0, mean: Tensor | None = None, std: Tensor | None = None,
    keep: Tensor | None = None, z: Tensor | None = None, z2: Tensor | None = None,
    philox_seed: int = 0, noise_mode: int = 0, rician: bool = False,
    gamma: Tensor | None = None,
) -> Tensor:
    """Fused bias -> blur -> noise -> gamma (two HBM passes); any stage optional.

    Compose-level fusion of consecutive intensity transforms; equal to running
    `bias_field`, `blur`, `noise`, `gamma` in sequence up to fp32 summation order.
    """
    _require_cuda(src, "intensity_fused")
    src = src.contiguous()
    b, c, i, j, k = src.shape
    dst = torch.empty_like(src)
    jk = taps is not None and (axes_mask & 6)
    scratch = torch.empty_like(src) if jk else None
    si = sj = sk = 0
    if coarse is not None:
        si, sj, sk = coarse.shape[2:]
    with torch.cuda.device(src.device):
        _native.call(
            "tio_intensity_fused", _ptr(src), _ptr(dst), _ptr(scratch), b, c, i, j, k,
            _ptr(coarse), si, sj, sk, _ptr(bias_identity), int(bool(bias_divide)),
            _ptr(taps), _ptr(radius), int(big_r), int(axes_mask),
            _ptr(mean), _ptr(std), _ptr(keep), _ptr(z), _ptr(z2),
            int(philox_seed) & (2**64 - 1), int(noise_mode), int(bool(rician)),
            _ptr(gamma), _stream(src),
        )
    _count(_fused_launches(axes_mask if taps is not None else 0, coarse is not None))
    return dst


# ---- exact replay of torch's CPU randn stream (K4a) ---------------------------

_MT_TABLE_FILE = _native.LIB_PATH.parent / "mt19937_jump.bin"
_mt_host_table: Tensor | None = None
_mt_device_tables: dict = {}
_mt_lock = threading.Lock()
MT_MAX_WORDS = 1 << 31  # stream positions the two-level jump table reaches


def mt19937_host_table() -> Tensor:
    """Jump-ahead table (constants of MT19937): loaded from the file written at
    build time, else computed on the host (~2 s) and cached."""
    global _mt_host_table
    with _mt_lock:
        if _mt_host_table is None:
            lib = _native.lib()
            nbytes = lib.tio_mt19937_table_bytes()
            blob = None
            if _MT_TABLE_FILE.exists() and _MT_TABLE_FILE.stat().st_size == nbytes:
                blob = torch.from_numpy(np.fromfile(_MT_TABLE_FILE, dtype=np.uint8))
            if blob is None:
                blob = torch.zeros(nbytes, dtype=torch.uint8)
                _native.call("tio_mt19937_build_table", blob.data_ptr(), nbytes)
                try:
                    blob.numpy().tofile(_MT_TABLE_FILE)
                except OSError:
                    pass
            _mt_host_table = blob
        return _mt_host_table


def _mt_table(device: torch.device) -> Tensor:
    key = (device.type, device.index)
    table = _mt_device_tables.get(key)
    if table is None:
        with _init_lock:
            table = _mt_device_tables.get(key)
            if table is None:
                table = mt19937_host_table().to(device)
                torch.cuda.synchronize(device)  # other threads' streams may read it right away
                _mt_device_tables[key] = table
    return table


def randn_mt19937(seed: int, offset: int, n: int, device, out: Tensor | None = None) -> Tensor:
    """Elements [offset, offset+n) of ``torch.randn(N, generator=CPU mt19937(seed))``
    computed on ``device`` (to ~1 ulp of the host's libm).  Needs offset, n
    multiples of 16 and offset + n <= 2**31; callers keep ragged tails on the host."""
    device = torch.device(device)
    if n < 16 or n % 16 or offset % 16 or offset + n > MT_MAX_WORDS:
        raise ValueError("randn_mt19937: offset and n must be multiples of 16, n >= 16,"
                         " offset + n <= 2**31")
    z = torch.empty(n, dtype=torch.float32, device=device) if out is None else out
    lib = _native.lib()
    ws_bytes = lib.tio_randn_mt19937_workspace_bytes(offset, n)
    workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    table = _mt_table(device)
    with torch.cuda.device(device):
        _native.call("tio_randn_mt19937", int(seed) & 0xFFFFFFFF, offset, n, _ptr(z), _ptr(table),
                     _ptr(workspace), ws_bytes, torch.cuda.current_stream(device).cuda_stream)
    _count(4)
    return z
