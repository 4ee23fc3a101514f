"""TetrominoHolder descriptor (reference components/tetromino_holder.py:14-21).  Only the reference default of one slot
is built into the device record."""


class TetrominoHolder:
    def __init__(self, size: int = 1):
        self.size = int(size)
