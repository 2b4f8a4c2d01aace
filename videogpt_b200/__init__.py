"""videogpt_b200 -- B200-native next-clip denoising path of Video-GPT.

Drop-in surface (same names as the reference's ``LVM`` package, ``LVM/__init__.py``):
``LVMProcessor``, ``LVM``, ``LVMScheduler``, ``LVMPipeline``; plus ``replace_attention`` (the
operator seam of ``LVM/transform/sdpa_transform.py``).  The CUDA library is loaded lazily on the
first kernel call and its absence is an error -- there is no CPU or PyTorch fallback.
"""
from .processor import LVMProcessor, LVMCollator, FrameGeometry  # noqa: F401
from .scheduler import LVMScheduler  # noqa: F401
from .model import LVM  # noqa: F401
from .pipeline import LVMPipeline  # noqa: F401
from .transform import replace_attention  # noqa: F401
from .rollout import LatentRollout  # noqa: F401

__all__ = ["LVMProcessor", "LVMCollator", "FrameGeometry", "LVMScheduler", "LVM", "LVMPipeline",
           "replace_attention", "LatentRollout"]
