"""RaySamples / RayIntervals containers.

Same fields and meaning as /root/reference/nerfacc/data_specs.py:12-180; the
``_to_cpp/_from_cpp`` pybind plumbing is gone because the native boundary here
takes raw pointers (include/nerfacc_b200.h).
"""
from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class RaySamples:
    """Samples along rays: batched ``(n_rays, n_samples)`` or flattened ``(all_samples,)``.

    For flattened data give ``packed_info`` (n_rays, 2) = (start, count) and/or
    ``ray_indices`` (all_samples,).  ``is_valid`` marks used slots of an
    over-allocated buffer (traverse_grids(..., over_allocate=True)).
    """

    vals: torch.Tensor
    packed_info: Optional[torch.Tensor] = None
    ray_indices: Optional[torch.Tensor] = None
    is_valid: Optional[torch.Tensor] = None

    @property
    def device(self) -> torch.device:
        return self.vals.device


@dataclass
class RayIntervals:
    """Interval edges along rays.

    ``vals`` holds edges; ``is_left`` / ``is_right`` say whether an edge opens /
    closes a sample, so contiguous and broken chains share one representation:
    ``t_starts = vals[is_left]``, ``t_ends = vals[is_right]``.
    """

    vals: torch.Tensor
    packed_info: Optional[torch.Tensor] = None
    ray_indices: Optional[torch.Tensor] = None
    is_left: Optional[torch.Tensor] = None
    is_right: Optional[torch.Tensor] = None

    @property
    def device(self) -> torch.device:
        return self.vals.device
