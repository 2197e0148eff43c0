"""Common base of the sampling estimators (reference nerfacc/estimators/base.py:7-22)."""
from typing import Any

import torch
from torch import nn


class AbstractEstimator(nn.Module):
    """A transmittance estimator: decides where along each ray samples are placed."""

    def __init__(self) -> None:
        super().__init__()
        # zero-size, non-persistent: only there so `.device` follows `.to()`
        self.register_buffer("_dummy", torch.empty(0), persistent=False)

    @property
    def device(self) -> torch.device:
        return self._dummy.device

    def sampling(self, *args, **kwargs) -> Any:
        raise NotImplementedError

    def update_every_n_steps(self, *args, **kwargs) -> None:
        raise NotImplementedError
