"""Env registration -- mirrors tetris_gymnasium/envs/__init__.py:10-14 (id "tetris_gymnasium/Tetris").

With gymnasium installed, `gym.make("tetris_gymnasium_b200/Tetris", num_envs=..., **reference_kwargs)`
returns the batched CUDA env; the id "tetris_gymnasium/Tetris" is registered too when the reference
package has not claimed it, so existing `gym.make("tetris_gymnasium/Tetris", ...)` call sites keep working.
"""
from .tetris import Tetris  # noqa: F401

try:  # gymnasium is optional: the env does not depend on it
    import gymnasium as _gym

    for _id in ("tetris_gymnasium_b200/Tetris", "tetris_gymnasium/Tetris"):
        if _id not in getattr(_gym, "registry", {}):
            _gym.register(id=_id, entry_point="tetris_gymnasium_b200.envs:Tetris", disable_env_checker=True,
                          order_enforce=False)
except Exception:  # pragma: no cover
    pass
