"""panogrf_b200 — B200-native (sm_100a) implementation of PanoGRF's render-time hot path.

Host-side mirror of the reference's module API over a C-ABI CUDA library; see DESIGN.md.
"""
from . import _lib  # noqa: F401
from .spherical_cost_volume import calculate_cost_volume_erp, calculate_cost_volume_erp_multiview  # noqa: F401

from .renderer import NeuralRayBaseRenderer, NeuralRayGenRenderer, name2network  # noqa: F401
from . import render_ops  # noqa: F401

__all__ = ["calculate_cost_volume_erp", "calculate_cost_volume_erp_multiview", "NeuralRayBaseRenderer", "NeuralRayGenRenderer", "name2network"]
