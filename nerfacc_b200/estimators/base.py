"""What every sampling estimator has in common (the role of nerfacc/estimators/base.py:7-22).

An estimator decides where along each ray samples are placed; concrete classes provide
``sampling`` and ``update_every_n_steps``.
"""
from typing import Any

import torch


class AbstractEstimator(torch.nn.Module):
    """Base class of :class:`OccGridEstimator` and :class:`PropNetEstimator`."""

    def __init__(self) -> None:
        super().__init__()
        # `.device` must follow `.to()` / `.cuda()`; an empty non-persistent buffer does that without
        # showing up in state_dict()
        self.register_buffer("_dummy", torch.empty(0), persistent=False)

    def update_every_n_steps(self, *args, **kwargs) -> None:
        """Refresh whatever the estimator learns during training; subclasses implement it."""
        raise NotImplementedError(f"{type(self).__name__} does not implement update_every_n_steps()")

    def sampling(self, *args, **kwargs) -> Any:
        """Place samples along the rays; subclasses implement it."""
        raise NotImplementedError(f"{type(self).__name__} does not implement sampling()")

    @property
    def device(self) -> torch.device:
        """Device the estimator's buffers live on."""
        return self._dummy.device
