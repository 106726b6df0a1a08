"""Stand-in for gym 0.15.4 — only the names envs/atc/atc_gym.py:4-8,22 touches."""
from . import spaces  # noqa: F401
from . import envs  # noqa: F401
from . import utils  # noqa: F401


class Env(object):
    metadata = {'render.modes': []}
    reward_range = (-float('inf'), float('inf'))
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode='human'):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return


def make(env_id, **kwargs):
    from .envs.registration import make as _make
    return _make(env_id, **kwargs)
