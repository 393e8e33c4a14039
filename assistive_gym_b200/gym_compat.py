"""`gym` is not installed in the build container.  Use it when importable,
otherwise a minimal stand-in with the three things the reference uses: `gym.Env`, `spaces.Box`,
`utils.seeding.np_random` (envs/env.py:3-5,79-80)."""
import numpy as np

try:                                    # pragma: no cover - depends on the box
    import gym                          # noqa: F401
    from gym import spaces              # noqa: F401
    from gym.utils import seeding       # noqa: F401
    HAVE_GYM = True
except Exception:                       # ImportError or a broken install
    HAVE_GYM = False

    class Env:
        metadata = {}
        reward_range = (-float('inf'), float('inf'))
        action_space = None
        observation_space = None

        def step(self, action):
            raise NotImplementedError

        def reset(self):
            raise NotImplementedError

        def render(self, mode='human'):
            return None

        def close(self):
            return None

        def seed(self, seed=None):
            return [seed]

    class _Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low = np.asarray(low, dtype=dtype) if shape is None else np.full(shape, low, dtype=dtype)
            self.high = np.asarray(high, dtype=dtype) if shape is None else np.full(shape, high, dtype=dtype)
            self.shape = self.low.shape
            self.dtype = np.dtype(dtype)
            self._rng = np.random.RandomState()

        def seed(self, seed=None):
            self._rng = np.random.RandomState(seed)
            return [seed]

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high)

    class _Spaces:
        Box = _Box

    class _Seeding:
        @staticmethod
        def np_random(seed=None):
            if seed is None:
                seed = int(np.random.SeedSequence().entropy % (2 ** 31))
            return np.random.RandomState(seed), seed

    class _Gym:
        Env = Env

    gym = _Gym()
    spaces = _Spaces()
    seeding = _Seeding()
