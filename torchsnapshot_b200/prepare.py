"""``_custom_tensor_prepare_func`` helpers (T:snapshot.py:120-122, 591-595; T:io_preparers/tensor.py:59-81).

The reference traces the hook at plan time to fill ``TensorEntry.dtype`` but then stages the *unprocessed* tensor
(T:io_preparers/tensor.py:240-258), so a casting hook yields an unreadable snapshot there.  Here the processed tensor
is what gets persisted — and when the hook is one of the recognisable casts below, no processed tensor is ever
materialised: the conversion is fused into the pack kernel (``tsnap_copy_desc.dst_dtype != src_dtype``), which reads
the live tensor once and writes the narrower wire image straight into the staging arena."""
from __future__ import annotations

import fnmatch
import functools
from typing import Callable, Optional

import torch

_FLOATS = (torch.float16, torch.bfloat16, torch.float32, torch.float64)


class _Traced:
    """What a tracing call returns when materialising the processed tensor would be wasteful or impossible
    (quantised meta tensors do not exist): dtype and shape, the two things T:io_preparers/tensor.py:59-81 reads."""

    def __init__(self, shape, dtype) -> None:
        self.shape, self.dtype = torch.Size(shape), dtype


def cast_on_save(dtype: torch.dtype, only: Optional[str] = None) -> Callable[[str, torch.Tensor, bool], torch.Tensor]:
    """A ``_custom_tensor_prepare_func`` that stores floating-point tensors as ``dtype`` (e.g. fp32 master weights as
    bf16).  ``only``: glob on the logical path; other tensors are stored unchanged.

        Snapshot.take(path, app_state, _custom_tensor_prepare_func=cast_on_save(torch.bfloat16, only="model/*"))
    """
    if dtype not in _FLOATS:
        raise ValueError(f"cast_on_save supports {_FLOATS}, not {dtype}")

    def applies(logical_path: Optional[str], tensor: torch.Tensor) -> Optional[torch.dtype]:
        if tensor.dtype not in _FLOATS or tensor.dtype == dtype or tensor.is_quantized:
            return None
        if only is not None and (logical_path is None or not fnmatch.fnmatch(logical_path, only)):
            return None
        return dtype

    def prepare(logical_path: str, tensor: torch.Tensor, tracing: bool) -> torch.Tensor:
        target = applies(logical_path, tensor)
        if target is None:
            return tensor
        if tracing:  # the planner only looks at dtype and shape
            return torch.empty(tensor.shape, dtype=target, device="meta")
        return tensor.to(target)

    prepare.tsnap_cast = applies  # type: ignore[attr-defined]  # recognised by native_plan: fused into the pack kernel
    return prepare


def fused_cast_of(func, tensor: torch.Tensor) -> Optional[torch.dtype]:
    """Wire dtype when ``func`` (possibly wrapped in functools.partial(func, logical_path), as Snapshot does,
    T:snapshot.py:591-595) is a recognised pure cast for ``tensor``; None otherwise."""
    args = []
    inner = func
    while isinstance(inner, functools.partial):
        args = list(inner.args) + args
        inner = inner.func
    spec = getattr(inner, "tsnap_cast", None)
    if spec is None:
        return None
    return spec(args[0] if args else None, tensor)


PER_TENSOR_QTENSOR = "per_tensor_qtensor"  # TensorEntry.serializer of the reference's per-tensor format (T:serialization.py:278-310)


def quantize_on_save(dtype: torch.dtype = torch.qint8, qparams: Optional[Callable[[str, torch.Tensor], "tuple[float, int]"]] = None,
                     only: Optional[str] = None) -> Callable[[str, torch.Tensor, bool], torch.Tensor]:
    """A ``_custom_tensor_prepare_func`` that stores floating-point tensors per-tensor affine quantised
    (``torch.qint8`` / ``torch.quint8``) in the reference's self-describing per-tensor format — int_repr bytes followed
    by ``[q_scale: double][q_zero_point: int64]`` (T:serialization.py:278-310, specified and unit-tested there but
    never wired into a preparer).  The quantisation runs inside the pack kernel (``TSNAP_QINT8``/``TSNAP_QUINT8``
    wire dtype); restoring into a floating-point tensor dequantises (``tensor_copy`` semantics,
    T:io_preparers/tensor.py:385-409).

    ``qparams(logical_path, tensor) -> (scale, zero_point)``; default: symmetric abs-max scaling for qint8
    (zero_point 0), min/max affine for quint8."""
    if dtype not in (torch.qint8, torch.quint8):
        raise ValueError("quantize_on_save supports torch.qint8 and torch.quint8")
    cache: dict = {}

    def default_qparams(_path: str, t: torch.Tensor):
        tf = t.detach().float()
        if dtype == torch.qint8:
            amax = float(tf.abs().max().item()) if tf.numel() else 0.0
            return (amax / 127.0) or 1.0, 0
        lo, hi = (float(tf.min().item()), float(tf.max().item())) if tf.numel() else (0.0, 0.0)
        lo, hi = min(lo, 0.0), max(hi, 0.0)
        scale = ((hi - lo) / 255.0) or 1.0
        return scale, int(round(-lo / scale))

    def applies(logical_path: Optional[str], tensor: torch.Tensor):
        if tensor.dtype not in _FLOATS or tensor.is_quantized or tensor.numel() == 0:
            return None
        if only is not None and (logical_path is None or not fnmatch.fnmatch(logical_path, only)):
            return None
        key = (logical_path, tensor.data_ptr(), tuple(tensor.shape), tensor._version)
        if key not in cache:
            cache.clear()  # planning and staging ask about the same tensor back to back: one entry is enough
            scale, zp = (qparams or default_qparams)(logical_path or "", tensor)
            cache[key] = (dtype, float(scale), int(zp))
        return cache[key]

    def prepare(logical_path: str, tensor: torch.Tensor, tracing: bool) -> torch.Tensor:
        spec = applies(logical_path, tensor)
        if spec is None:
            return tensor
        _, scale, zp = spec
        if tracing:  # the planners (this package's and the reference's) only look at .dtype and .shape
            return _Traced(tensor.shape, dtype)
        return torch.quantize_per_tensor(tensor.detach().float(), scale, zp, dtype)

    prepare.tsnap_quant = applies  # type: ignore[attr-defined]
    return prepare


def fused_quant_of(func, tensor: torch.Tensor):
    """(qdtype, scale, zero_point) when ``func`` is a recognised quantise-on-save hook for ``tensor``, else None."""
    args = []
    inner = func
    while isinstance(inner, functools.partial):
        args = list(inner.args) + args
        inner = inner.func
    spec = getattr(inner, "tsnap_quant", None)
    if spec is None:
        return None
    return spec(args[0] if args else None, tensor)
