"""rusterize_b200 — B200-native (sm_100a) implementation of rusterize's polygon / line / point
burn path, behind the reference's `rusterize()` surface.  See DESIGN.md."""
from .core import Geoms, group_keys, raster_info, rasterize_dense  # noqa: F401

__version__ = "0.1.0"
