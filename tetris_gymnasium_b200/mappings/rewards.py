"""RewardsMapping -- same fields and defaults as the reference (tetris_gymnasium/mappings/rewards.py:12-15)."""
from dataclasses import dataclass


@dataclass
class RewardsMapping:
    alife: float = 1
    clear_line: float = 1
    game_over: float = 0
    invalid_action: float = -0.1
