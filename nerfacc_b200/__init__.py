"""nerfacc_b200 -- Blackwell-native sampling + volume-rendering hot path with nerfacc's API.

The public names below are the subset of /root/reference/nerfacc/__init__.py that
lies on the hot path; `import nerfacc` resolves to this
package through the alias package at the repository root.
"""
from ._lib import defer_until_wait
from .data_specs import RayIntervals, RaySamples
from .estimators.occ_grid import OccGridEstimator
from .estimators.prop_net import PropNetEstimator
from .grid import ray_aabb_intersect, traverse_grids
from .losses import distortion
from .pack import pack_info
from .pdf import importance_sampling, searchsorted
from .scan import exclusive_prod, exclusive_sum, inclusive_prod, inclusive_sum
from .version import __version__
from .volrend import (
    accumulate_along_rays,
    accumulate_along_rays_,
    render_transmittance_from_alpha,
    render_transmittance_from_density,
    render_visibility_from_alpha,
    render_visibility_from_density,
    render_weight_from_alpha,
    render_weight_from_density,
    rendering,
)

__all__ = [
    "__version__",
    "inclusive_prod",
    "exclusive_prod",
    "inclusive_sum",
    "exclusive_sum",
    "pack_info",
    "render_visibility_from_alpha",
    "render_visibility_from_density",
    "render_weight_from_alpha",
    "render_weight_from_density",
    "render_transmittance_from_alpha",
    "render_transmittance_from_density",
    "accumulate_along_rays",
    "accumulate_along_rays_",
    "rendering",
    "RayIntervals",
    "RaySamples",
    "ray_aabb_intersect",
    "traverse_grids",
    "OccGridEstimator",
    "PropNetEstimator",
    "importance_sampling",
    "searchsorted",
    "distortion",
    "defer_until_wait",
]
