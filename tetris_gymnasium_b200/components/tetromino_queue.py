"""TetrominoQueue descriptor (reference components/tetromino_queue.py:15-22): randomizer + visible queue length."""


class TetrominoQueue:
    def __init__(self, randomizer=None, size: int = 4):
        self.randomizer = randomizer
        self.size = int(size)
