"""Import the UNMODIFIED reference NumPy env (oracle tooling and the bench's reference arm only).

Looks for the reference in /root/reference (the build container) and then in baseline/_ref (the pip --target install made
by oracle/install_reference.py; git-ignored, travels to the GPU box with the snapshot).  oracle/gymnasium_shim goes on sys.path
only when the real gymnasium is absent.  Where neither copy exists `available()` is False and everything that needs the live
reference is skipped.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(_HERE, "gymnasium_shim")
_CANDIDATES = [os.environ.get("TETRIS_REFERENCE", "/root/reference"), os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]


def _find():
    for c in _CANDIDATES:
        if c and os.path.isdir(os.path.join(c, "tetris_gymnasium")):
            return c
    return None


REF = _find() or _CANDIDATES[0]


def available():
    return _find() is not None


def where():
    """'source tree' (/root/reference) or 'baseline/_ref' (installed copy) -- recorded by the bench's reference arm."""
    f = _find()
    return None if f is None else ("baseline/_ref" if f.endswith(os.path.join("baseline", "_ref")) else f)


def load():
    if not available():
        raise RuntimeError("reference not present")
    try:
        import gymnasium  # noqa: F401
    except ImportError:
        sys.path.insert(0, SHIM)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import tetris_gymnasium.envs  # noqa: F401  (registers the env id)
    from tetris_gymnasium.components.tetromino import Tetromino
    from tetris_gymnasium.components.tetromino_holder import TetrominoHolder
    from tetris_gymnasium.components.tetromino_queue import TetrominoQueue
    from tetris_gymnasium.components.tetromino_randomizer import Randomizer, TrueRandomizer
    from tetris_gymnasium.envs.tetris import Tetris
    from tetris_gymnasium.wrappers.grouped import GroupedActionsObservations
    from tetris_gymnasium.wrappers.observation import FeatureVectorObservation, RgbObservation

    class Scripted(Randomizer):
        """Injected piece stream: the same hook SURVEY 8c describes (post-construction patch)."""

        def __init__(self, seq):
            super().__init__(7)
            self.seq = [int(v) for v in seq]
            self.cur = 0

        def get_next_tetromino(self):
            v = self.seq[self.cur % len(self.seq)]
            self.cur += 1
            return v

        def reset(self, seed=None):
            pass

    def make(width=10, height=20, gravity=True, queue_size=4, seq=None, true_random=False, **kw):
        env = Tetris(width=width, height=height, gravity=gravity, **kw)
        if true_random:
            # Tetris(randomizer=...) leaves self.randomizer unset (envs/tetris.py:138-141 only assigns the defaults),
            # so the randomizer is swapped in after construction like the scripted one
            env.randomizer = TrueRandomizer(len(env.tetrominoes))
            env.queue = TetrominoQueue(env.randomizer, size=queue_size)
        elif seq is not None:
            env.randomizer = Scripted(seq)
            env.queue = TetrominoQueue(env.randomizer, size=queue_size)
        elif queue_size != 4:
            env.queue = TetrominoQueue(env.randomizer, size=queue_size)
        return env

    return dict(Tetris=Tetris, make=make, TetrominoHolder=TetrominoHolder, Tetromino=Tetromino, Scripted=Scripted, TrueRandomizer=TrueRandomizer, TetrominoQueue=TetrominoQueue,
                GroupedActionsObservations=GroupedActionsObservations,
                FeatureVectorObservation=FeatureVectorObservation, RgbObservation=RgbObservation)
