"""Seed handling with the reference's convention (utils/random.py:16-71): None -> numpy's global RandomState, an int -> a
fresh RandomState, a RandomState -> itself.  ``pipeline.countsketch_table`` draws from it in the reference's order."""
from typing import Union

import numpy as np

RandomStateLike = Union[None, int, np.random.RandomState]


def check_random_state(seed: RandomStateLike) -> np.random.RandomState:
    if seed is None or seed is np.random:
        return np.random.mtrand._rand
    if isinstance(seed, (int, np.integer)):
        return np.random.RandomState(seed)
    if isinstance(seed, np.random.RandomState):
        return seed
    raise ValueError(f"'{seed}' cannot be used to seed a numpy.random.RandomState instance. "
                     f"Expected None, int, or np.random.RandomState, got {type(seed)}.")
