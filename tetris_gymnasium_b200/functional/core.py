"""EnvConfig / State of the functional env (reference functional/core.py:11-42), with torch tensors."""
from dataclasses import dataclass, replace
from typing import NamedTuple

import torch


class EnvConfig(NamedTuple):
    """Same fields and defaults as the reference (functional/core.py:11-25)."""

    width: int
    height: int
    padding: int
    queue_size: int
    gravity_enabled: bool = True


@dataclass
class State:
    """Reference functional/core.py:28-42.  Batched: every field has a leading env axis (size 1 for the
    un-batched reset/step).  `score` is float32 like the reference (jnp.float32(0) at reset)."""

    rng_key: torch.Tensor          # u32-valued int64 [B, 2]
    board: torch.Tensor            # int8 [B, H_pad, W_pad]
    active_tetromino: torch.Tensor  # int32 [B]
    rotation: torch.Tensor         # int32 [B]
    x: torch.Tensor                # int32 [B]
    y: torch.Tensor                # int32 [B]
    queue: torch.Tensor            # int32 [B, queue_size]
    queue_index: torch.Tensor      # int32 [B]
    game_over: torch.Tensor        # bool [B]
    score: torch.Tensor            # float32 [B]

    def replace(self, **kw):
        return replace(self, **kw)
