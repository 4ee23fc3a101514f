"""TetrominoHolder descriptor (reference components/tetromino_holder.py:14-21): `Tetris(holder=TetrominoHolder(size))`.
The held pieces live in the env's hot record on the device: one (piece, rotation) pair for the default size 1, a FIFO of up
to four pairs for size 2..4 (swap stores the active piece and hands back the oldest one once the holder is full, :31-49)."""


class TetrominoHolder:
    def __init__(self, size: int = 1):
        self.size = int(size)
        if not 1 <= self.size <= 4:
            raise ValueError("holder size must be 1..4")
