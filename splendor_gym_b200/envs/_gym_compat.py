"""gymnasium is optional: use it when importable, otherwise the few classes SplendorEnv touches
(envs/splendor_env.py:3-4,23,33-36,42 of the reference)."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the image
    import gymnasium as gym
    from gymnasium import spaces

    Env = gym.Env
    Wrapper = gym.Wrapper
    HAVE_GYMNASIUM = True
except Exception:  # noqa: BLE001
    HAVE_GYMNASIUM = False

    class Env:
        metadata: dict = {}
        _np_random = None

        def reset(self, *, seed=None, options=None):
            if seed is not None:  # gymnasium.utils.seeding.np_random(seed)
                self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))

        @property
        def np_random(self):
            if self._np_random is None:
                self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence()))
            return self._np_random

        def close(self):
            pass

    class Wrapper:
        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)
            return getattr(self.env, name)

        def reset(self, **kw):
            return self.env.reset(**kw)

        def step(self, action):
            return self.env.step(action)

    class _Discrete:
        def __init__(self, n):
            self.n = int(n)
            self.shape = ()
            self.dtype = np.int64

        def contains(self, x):
            return 0 <= int(x) < self.n

        def sample(self):
            return int(np.random.randint(self.n))

    class _Box:
        def __init__(self, low, high, shape, dtype):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool((x >= self.low).all())

    class spaces:  # noqa: N801
        Discrete = _Discrete
        Box = _Box
