"""Synthetic scenes for parity tests and bench.py.

numpy-only and seeded, so the same inputs can be rebuilt on any box.
"""
import numpy as np


def ball_grid(res: int = 128, radius: float = 0.5) -> np.ndarray:
    """[1, res, res, res] bool: cells whose centre lies in the ball |x| <= radius of roi [-1, 1]^3."""
    c = (np.arange(res, dtype=np.float64) + 0.5) / res * 2.0 - 1.0
    x, y, z = np.meshgrid(c, c, c, indexing="ij")
    return (x * x + y * y + z * z <= radius * radius)[None]


def ball_rays(n_rays: int, seed: int = 42, dist: float = 4.0, radius: float = 0.5):
    """Rays from a sphere of radius `dist` aimed at uniformly drawn points of the disk of
    radius `radius` through the origin, perpendicular to the view direction (all rays hit)."""
    rng = np.random.default_rng(seed)
    u = rng.standard_normal((n_rays, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    a = rng.standard_normal((n_rays, 3))
    a -= (a * u).sum(1, keepdims=True) * u
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = np.cross(u, a)
    rad = radius * np.sqrt(rng.random(n_rays))
    ang = 2.0 * np.pi * rng.random(n_rays)
    target = (rad * np.cos(ang))[:, None] * a + (rad * np.sin(ang))[:, None] * b
    o = dist * u
    d = target - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o.astype(np.float32), d.astype(np.float32)


ROI_AABB = np.array([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0], dtype=np.float32)

# render_step_size calibrated once on the oracle so that the ball scene gives ~128 samples / ray
# (mean samples/ray in [122, 134]); measured 129.1 on 4096 rays.
BALL_STEP = 5.2e-3


def nested_aabbs(levels: int) -> np.ndarray:
    c, e = (ROI_AABB[:3] + ROI_AABB[3:]) / 2, (ROI_AABB[3:] - ROI_AABB[:3]) / 2
    return np.stack([np.concatenate([c - e * 2 ** i, c + e * 2 ** i]) for i in range(levels)]).astype(np.float32)
