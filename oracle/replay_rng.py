"""TEST INFRASTRUCTURE ONLY — replay recorded draws through NumPy's global-RNG entry points.

The reference's `sensor` class draws from `np.random.normal` / `np.random.random` (environment/quadrotor_env.py:600-662).
Inside `with ReplayRNG(normals, uniforms):` those two functions return the recorded STANDARD draws instead
(`normal(loc, scale, size)` = loc + scale * z, exactly what NumPy's legacy generator computes from its own gaussians), so
the reference class, the oracle and the CUDA-backed drop-in `sensor` can all be driven by one and the same stream."""
import numpy as np


class ReplayRNG:
    def __init__(self, normals, uniforms=()):
        self.z = np.asarray(normals, dtype=np.float64).ravel()
        self.u = np.asarray(uniforms, dtype=np.float64).ravel()
        self.kz = self.ku = 0

    def _take(self, buf, k, n, what):
        if k + n > buf.size:
            raise RuntimeError("ReplayRNG: the recorded %s stream is exhausted" % what)
        return buf[k:k + n]

    def normal(self, loc=0.0, scale=1.0, size=None):
        shape = np.broadcast(np.asarray(loc), np.asarray(scale)).shape if size is None else tuple(np.atleast_1d(size))
        n = int(np.prod(shape, dtype=np.int64))
        z = self._take(self.z, self.kz, n, "normal").reshape(shape)
        self.kz += n
        out = loc + scale * z
        return float(out) if shape == () else out

    def random(self, size=None):
        shape = () if size is None else tuple(np.atleast_1d(size))
        n = int(np.prod(shape, dtype=np.int64))
        u = self._take(self.u, self.ku, n, "uniform").reshape(shape)
        self.ku += n
        return float(u) if shape == () else u.copy()

    def __enter__(self):
        self._saved = (np.random.normal, np.random.random)
        np.random.normal, np.random.random = self.normal, self.random
        return self

    def __exit__(self, *exc):
        np.random.normal, np.random.random = self._saved
        return False
