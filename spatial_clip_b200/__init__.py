"""spatial_clip_b200 -- B200-native (sm_100a) implementation of Spatial-Clip's contrastive-loss hot path.

Public API mirrors the reference's loss modules (see losses.py); the compute is CUDA-only behind the
C ABI declared in include/scl_b200.h.
"""
from .losses import (ClipLoss, GlobalMappingMultiPositiveClipLoss, SpatialLoss, SpatialLossFromColumns,  # noqa: F401
                     release_cuda_graphs)

__all__ = ["SpatialLoss", "ClipLoss", "GlobalMappingMultiPositiveClipLoss", "SpatialLossFromColumns",
           "release_cuda_graphs"]
__version__ = "0.1.0"
