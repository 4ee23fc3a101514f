"""Randomizer descriptors (reference components/tetromino_randomizer.py).  The draws happen on the device:
BagRandomizer -> TG_RANDOMIZER_BAG (7-bag, in-place reshuffle, :49-102), TrueRandomizer -> TG_RANDOMIZER_TRUE
(`rng.integers(0, size)` per draw, :105-136)."""


class Randomizer:
    kind = None

    def __init__(self, size: int = 7):
        if not 1 <= size <= 7:
            raise ValueError("tetromino sets hold 1..7 pieces")
        self.size = size    # informational: the device draws from the env's tetromino set


class BagRandomizer(Randomizer):
    kind = "bag"


class TrueRandomizer(Randomizer):
    kind = "true"
