"""atc-reinforcement-learning_b200 — B200-native batched ATC approach-control environment.

Drop-in for the hot path of fvalka/atc-reinforcement-learning (AtcGym.step / reset): the gym surface with a leading
batch dimension, hand-written sm_100a CUDA kernels behind a C ABI (include/atc_b200.h), PyTorch owning the tensors.
Importing the env classes requires the built CUDA library; there is no CPU fallback.
"""
from .model import SimParameters, EntryPoint, MinimumVectoringAltitude, Runway  # noqa: F401
from .scenarios import Scenario, LOWW, SimpleScenario, load_scenario  # noqa: F401
from .sector import CompiledSector  # noqa: F401

__all__ = ['SimParameters', 'EntryPoint', 'MinimumVectoringAltitude', 'Runway', 'Scenario', 'LOWW', 'SimpleScenario',
           'load_scenario', 'CompiledSector', 'BatchedAtcEnv', 'AtcGym', 'make']


def __getattr__(name):          # the env needs torch + the native library: import lazily
    if name in ('BatchedAtcEnv', 'AtcGym', 'make', 'TERM_NAMES'):
        from . import env
        return getattr(env, name)
    raise AttributeError(name)
