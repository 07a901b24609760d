"""Multi-GPU helpers above the C ABI (additive; the reference is single-node shared-memory only).

* Horizon: ``horizon.horizon_gridded(..., devices=N)`` (``hzb_horizon_gridded_multi``): the 4-row blocks
  of the inner domain dealt out to the GPUs -- the reference's row partition (``horizon_comp.cpp:739-744``).
* Shadow / sw_dir_cor: the reference computes one sun position per call (``shadow_comp.cpp:386-605``) and a
  map is only 13 MB for a 3601 x 3601 domain, so sharding the ROWS of one position would be latency-bound;
  ``MultiTerrain`` replicates the terrain (DEM + BVH + per-cell inputs) on every GPU and deals the SUN
  POSITIONS of a batch out to them, one host thread per GPU.
"""
import threading

import numpy as np

from . import resident, shadow


def _set_device(dev):
    resident._check(resident.lib().hzb_set_device(int(dev)))


class MultiTerrain:
    """``shadow.Terrain`` replicated on several GPUs; batches of sun positions are split between them.

    ``devices``: ``0`` = all visible GPUs, ``n`` = the first n, or an explicit list of device indices
    (repeats allowed: two replicas on one GPU, as the single-GPU test does)."""

    def __init__(self, devices=0):
        if isinstance(devices, int):
            n = resident.lib().hzb_device_count()
            if n <= 0:
                raise RuntimeError("horayzon_b200: no CUDA device available (no CPU fallback)")
            devices = list(range(n if devices <= 0 else min(devices, n)))
        self.devices = [int(d) for d in devices]
        self.terrains = [None] * len(self.devices)

    def _each(self, fn):
        errs = [None] * len(self.devices)

        def work(r):
            try:
                _set_device(self.devices[r])
                fn(r)
            except Exception as exc:  # re-raised in the caller's thread
                errs[r] = exc
        th = [threading.Thread(target=work, args=(r,)) for r in range(1, len(self.devices))]
        for t in th:
            t.start()
        work(0)
        for t in th:
            t.join()
        for e in errs:
            if e is not None:
                raise e

    def initialise(self, *args, **kwargs):
        """Arguments of ``shadow.Terrain.initialise`` (``shadow.pyx:27-86``); every GPU gets a copy."""
        def init(r):
            t = shadow.Terrain()
            t.initialise(*args, **kwargs)
            self.terrains[r] = t
        self._each(init)

    def _batch(self, method, sun_positions, dtype):
        sun_positions = np.ascontiguousarray(sun_positions, np.float32)
        if sun_positions.ndim != 2 or sun_positions.shape[1] != 3:
            raise ValueError("array 'sun_positions' has incorrect shape")
        n = len(sun_positions)
        bounds = np.linspace(0, n, len(self.devices) + 1).astype(int)
        parts = [None] * len(self.devices)

        def run(r):
            b, e = bounds[r], bounds[r + 1]
            if e > b:
                parts[r] = getattr(self.terrains[r], method)(sun_positions[b:e])
        self._each(run)
        parts = [p for p in parts if p is not None]
        if not parts:
            t = self.terrains[0]
            return np.empty((0,) + tuple(getattr(t, "shape", (0, 0))), dtype)
        return np.concatenate(parts, axis=0)

    def shadow_batch(self, sun_positions):
        """uint8 (n, y, x): ``Terrain.shadow`` for every row of ``sun_positions`` (n, 3)."""
        return self._batch("shadow_batch", sun_positions, np.uint8)

    def sw_dir_cor_batch(self, sun_positions):
        """float32 (n, y, x): ``Terrain.sw_dir_cor`` for every row of ``sun_positions`` (n, 3)."""
        return self._batch("sw_dir_cor_batch", sun_positions, np.float32)
