from __future__ import annotations

from collections import UserDict
from typing import Any, Dict, Protocol, runtime_checkable

import torch


@runtime_checkable
class Stateful(Protocol):
    def state_dict(self) -> Dict[str, Any]: ...

    def load_state_dict(self, state_dict: Dict[str, Any]) -> None: ...


AppState = Dict[str, Stateful]


class StateDict(UserDict):
    """A dict that is its own state dict: captures loose tensors / primitives in the app state."""

    def state_dict(self) -> Dict[str, Any]:
        return self.data

    def load_state_dict(self, state_dict: Dict[str, Any]) -> None:
        self.data.update(state_dict)


class RNGState:
    """Global torch RNG state; restored last so that take() and restore() leave identical RNG streams."""

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {"rng_state": torch.get_rng_state()}

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor]) -> None:
        torch.set_rng_state(state_dict["rng_state"])
