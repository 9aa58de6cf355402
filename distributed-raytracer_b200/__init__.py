"""distributed-raytracer_b200: B200 (sm_100a) GPU backend for the render hot path of
MWindels/distributed-raytracer (primary rays -> nearest-hit ray/triangle -> shadow rays + Phong ->
8-bit region framebuffer), behind a C ABI (include/rt_b200.h) that a cgo shim binds.

The directory name carries a hyphen (it mirrors the reference's repository name); import it as
`distributed_raytracer_b200` through the shim module at the repo root.
"""
from . import _lib
from ._lib import (FP32_FAST, FP64_EXACT, LANE, FLAG_BRUTE_FORCE, FLAG_COUNT, FLAG_PERSISTENT, FLAG_SIMPLE, FLAG_FUSED, FLAG_PIPELINE, FLAG_NO_CULL, FLAG_SM_AFFINE, LAYOUT_COLMAJOR, LAYOUT_ROWMAJOR,
                   QUANT_RGB255, QUANT_RGBA16_SHR8, RtError)
from .gpu_scene import DeviceBuffer, FrameState, GpuGroup, GpuScene, HostSharedBuffer, make_camera, make_scene_desc, tiles_per_rank

__all__ = ["GpuScene", "GpuGroup", "DeviceBuffer", "HostSharedBuffer", "FrameState", "RtError", "FP32_FAST", "FP64_EXACT", "LANE", "FLAG_BRUTE_FORCE", "FLAG_COUNT", "FLAG_SIMPLE", "FLAG_PERSISTENT", "FLAG_FUSED", "FLAG_PIPELINE", "FLAG_NO_CULL", "FLAG_SM_AFFINE",
           "LAYOUT_COLMAJOR", "LAYOUT_ROWMAJOR", "QUANT_RGB255", "QUANT_RGBA16_SHR8", "make_camera", "make_scene_desc",
           "tiles_per_rank", "_lib"]
