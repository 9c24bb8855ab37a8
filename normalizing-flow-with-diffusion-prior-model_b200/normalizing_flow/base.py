"""Abstract contracts of the flow package (mirror of the reference's normalizing_flow/base.py:9-83).

Only the interface is shared with the reference: ``Transform.transform(x, log_det_jac, logp)`` returning
``(y, log_det_jac, logp)``, ``Transform.invert(y)``, ``Prior.sample`` / ``Prior.compute_log_prob`` and the
``device`` attribute callers read (run_baseline_experiment.py:44-45).
"""
from abc import ABC, abstractmethod

import torch
import torch.nn as nn


def _default_device() -> torch.device:
    # reference: base.py:18 / :58 — CUDA when present.  The kernels need CUDA; a CPU-only process can still
    # construct modules (state_dict work, tests of the host logic) but cannot run them.
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


class Transform(nn.Module, ABC):
    def __init__(self):
        super().__init__()
        self.device = _default_device()

    @abstractmethod
    def transform(self, x, log_det_jac, logp):
        """x -> (f(x), log_det_jac, logp); both accumulators are updated IN PLACE and returned."""

    @abstractmethod
    def invert(self, y):
        """y -> f^{-1}(y)."""


class Prior(nn.Module, ABC):
    def __init__(self):
        super().__init__()
        self.device = _default_device()

    @abstractmethod
    def sample(self, shape):
        ...

    @abstractmethod
    def compute_log_prob(self, x):
        ...
