"""gym.utils.seeding stand-in.  The env's spawn RNG is Python's global `random`
seeded with the value returned here (atc_gym.py:124-125), so return it unchanged."""
import numpy as np


def np_random(seed=None):
    rng = np.random.RandomState()
    rng.seed(seed if seed is None else int(seed) % (2 ** 32))
    return rng, seed
