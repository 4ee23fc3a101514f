"""`RandomNumberGenerator` is an alias of numpy's Generator in real gymnasium as well."""
import numpy as np

RandomNumberGenerator = np.random.Generator


def np_random(seed=None):
    ss = np.random.SeedSequence(seed)
    return np.random.Generator(np.random.PCG64(ss)), ss.entropy
