"""horayzon_b200 -- B200-native drop-in for HORAYZON's horizon / shadow / SVF path.

``horizon``, ``shadow``, ``topo_param``, ``transform`` and ``direction`` mirror the
reference modules of the same names (``horayzon/__init__.py:1-12``); ``resident`` is the additive
device-pointer tier used for benchmarking and multi-GPU sharding; ``multi`` / ``sharding`` hold the multi-GPU helpers; ``synthetic``
holds the synthetic DEM recipes and the vertex-buffer wire format.
"""
try:
    from . import horizon, shadow, topo_param, transform, direction  # noqa: F401  (compiled Cython wrappers)
except ImportError as exc:  # fail loudly: there is no Python/CPU fallback
    raise ImportError(
        "horayzon_b200 extension modules are not built (" + str(exc) + "); run "
        "`python horayzon_b200/_build.py` or `__graft_entry__.build()`") from exc
from . import synthetic, resident, multi, sharding  # noqa: F401
from .synthetic import rearrange_pad_buffer, pad_buffer  # noqa: F401

__version__ = "0.1"
