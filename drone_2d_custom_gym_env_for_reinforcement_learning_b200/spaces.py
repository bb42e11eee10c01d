"""Box space: gymnasium's / gym's if importable, else a minimal stand-in with
the attributes SB3 and the reference scripts read (shape, dtype, low, high,
sample, contains)."""
import numpy as np

try:  # pragma: no cover - not installed in the build container
    from gymnasium.spaces import Box  # type: ignore
except Exception:  # noqa: BLE001
    try:  # pragma: no cover
        from gym.spaces import Box  # type: ignore
    except Exception:  # noqa: BLE001
        class Box:  # type: ignore
            def __init__(self, low, high, shape=None, dtype=np.float32):
                low = np.asarray(low, dtype=dtype)
                high = np.asarray(high, dtype=dtype)
                if shape is not None:
                    low = np.broadcast_to(low, shape).copy()
                    high = np.broadcast_to(high, shape).copy()
                self.low, self.high = low, high
                self.shape = low.shape
                self.dtype = np.dtype(dtype)
                self._rng = np.random.default_rng()

            def seed(self, seed=None):
                self._rng = np.random.default_rng(seed)
                return [seed]

            def sample(self):
                return self._rng.uniform(self.low, self.high).astype(self.dtype)

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

            def __repr__(self):
                return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)
