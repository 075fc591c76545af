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
