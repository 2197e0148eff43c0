"""Regularisers built from the packed scans (reference nerfacc/losses.py:7-41)."""
from torch import Tensor

from .scan import exclusive_sum
from .volrend import accumulate_along_rays


def distortion(weights: Tensor, t_starts: Tensor, t_ends: Tensor, ray_indices: Tensor, n_rays: int) -> Tensor:
    """Mip-NeRF 360 distortion loss per ray, shape ``(n_rays, 1)``.

    ``sum_i w_i^2 (e_i - s_i) / 3  +  2 sum_i w_i (m_i W_i - M_i)`` with ``m`` the interval midpoints and
    ``W`` / ``M`` the exclusive per-ray prefix sums of ``w`` / ``w m`` (the pairwise term in O(N)).
    All inputs are flattened ``(n_samples,)`` tensors addressed by ``ray_indices``.
    """
    if not (weights.shape == t_starts.shape == t_ends.shape == ray_indices.shape):
        raise AssertionError(
            f"the shape of the inputs are not the same: weights {weights.shape}, t_starts {t_starts.shape}, "
            f"t_ends {t_ends.shape}, ray_indices {ray_indices.shape}")
    mids = 0.5 * (t_starts + t_ends)
    widths = t_ends - t_starts
    within = (1 / 3) * (widths * weights.pow(2))
    w_before = exclusive_sum(weights, indices=ray_indices)
    wm_before = exclusive_sum(weights * mids, indices=ray_indices)
    across = 2 * (weights * mids * w_before - weights * wm_before)
    return accumulate_along_rays(within + across, None, ray_indices, n_rays)
