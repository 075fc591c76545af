"""Wire-format vocabulary: dtype names, serializer tags and the raw ("buffer_protocol") codec.

The raw payload of a tensor is ``numel * element_size`` bytes: the C-contiguous, native-endian image
of the *logical* view, no header (T:serialization.py:177-204, 254-265).  On the device path these
bytes are produced by the pack kernel; the helpers below are the host-side equivalents used for CPU
objects that do not go through the engine and by the tests."""
from __future__ import annotations

import io
import warnings
from enum import Enum
from typing import Dict, List

import torch

# name <-> dtype: exactly the strings the reference persists (T:serialization.py:72-88)
_NAMES: Dict[torch.dtype, str] = {
    dt: str(dt)
    for dt in (
        torch.float64,
        torch.float32,
        torch.float16,
        torch.bfloat16,
        torch.complex128,
        torch.complex64,
        torch.int64,
        torch.int32,
        torch.int16,
        torch.int8,
        torch.uint8,
        torch.bool,
        torch.qint32,
        torch.qint8,
        torch.quint8,
    )
}
_BY_NAME: Dict[str, torch.dtype] = {v: k for k, v in _NAMES.items()}
_ELEMENT_SIZE: Dict[torch.dtype, int] = {
    torch.float64: 8, torch.float32: 4, torch.float16: 2, torch.bfloat16: 2,
    torch.complex128: 16, torch.complex64: 8,
    torch.int64: 8, torch.int32: 4, torch.int16: 2, torch.int8: 1, torch.uint8: 1, torch.bool: 1,
    torch.qint32: 4, torch.qint8: 1, torch.quint8: 1,
}  # fmt: skip

ALL_SUPPORTED_DTYPES: List[torch.dtype] = list(_NAMES)
SUPPORTED_QUANTIZED_DTYPES: List[torch.dtype] = [torch.qint32, torch.qint8, torch.quint8]
# dtypes whose payload is the raw image (T:serialization.py:162-173); everything else uses torch.save
BUFFER_PROTOCOL_SUPPORTED_DTYPES: List[torch.dtype] = [
    torch.float64, torch.float32, torch.float16, torch.bfloat16,
    torch.int64, torch.int32, torch.int16, torch.int8, torch.uint8, torch.bool,
]  # fmt: skip


class Serializer(Enum):
    TORCH_SAVE = "torch_save"
    BUFFER_PROTOCOL = "buffer_protocol"
    PER_TENSOR_QTENSOR = "per_tensor_qtensor"
    PER_CHANNEL_QTENSOR = "per_channel_qtensor"


def _unsupported(what) -> ValueError:
    return ValueError(f"Unsupported dtype {what}. (Supported dtypes are: {ALL_SUPPORTED_DTYPES})")


def dtype_to_string(dtype: torch.dtype) -> str:
    try:
        return _NAMES[dtype]
    except KeyError:
        raise _unsupported(dtype) from None


def string_to_dtype(s: str) -> torch.dtype:
    try:
        return _BY_NAME[s]
    except KeyError:
        raise _unsupported(s) from None


def dtype_to_element_size(dtype: torch.dtype) -> int:
    try:
        return _ELEMENT_SIZE[dtype]
    except KeyError:
        raise _unsupported(dtype) from None


def tensor_as_memoryview(tensor: torch.Tensor) -> memoryview:
    """Zero-copy raw image of a CPU tensor (a copy is made only if the view is not dense)."""
    if tensor.dtype not in BUFFER_PROTOCOL_SUPPORTED_DTYPES:
        raise ValueError(f"tensor_as_memoryview() doesn't support the dtype {tensor.dtype}.")
    if tensor.device.type != "cpu":
        raise ValueError("tensor_as_memoryview() only accepts CPU tensors.")
    flat = tensor.detach().contiguous().reshape(-1)
    if flat.numel() == 0:
        return memoryview(b"")
    # reinterpret as bytes: works for every raw dtype incl. bfloat16, which numpy cannot name
    return memoryview(flat.view(torch.uint8).numpy()).cast("b")


def tensor_from_memoryview(mv, dtype: torch.dtype, shape: List[int]) -> torch.Tensor:
    numel = 1
    for s in shape:
        numel *= s
    if numel == 0:
        return torch.empty(shape, dtype=dtype)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # read-only buffers are fine: the result is a temporary
        return torch.frombuffer(mv, dtype=dtype).reshape(shape)


def torch_save_as_bytes(obj) -> bytes:
    buf = io.BytesIO()
    torch.save(obj, buf)
    return buf.getvalue()


def torch_load_from_bytes(buf):
    return torch.load(io.BytesIO(buf), weights_only=False)
