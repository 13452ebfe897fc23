"""flux_b200 — B200-native per-pixel render loop of jtdaugherty/flux.

Host-side mirror of the reference's interface for the render path
(scene data model, Scene/Camera/GpuWorker) over the C-ABI of
``include/fluxb200.h`` (``flux_b200/lib/libfluxb200.so``, hand-written sm_100a
CUDA).  There is no CPU fallback: rendering without the built library or
without a CUDA device raises.
"""
from .scene import (CameraData, CameraSettings, Emissive, GlossyReflective, JobConfiguration, Matte,
                    MeshData, OutputSettings, PlaneData, Reflective, SceneData, SphereData,
                    TriangleData, RectangleData, BoxData, WorkUnit, WorkUnitResult, work_units)

__all__ = ["CameraData", "CameraSettings", "Emissive", "GlossyReflective", "JobConfiguration", "Matte",
           "MeshData", "OutputSettings", "PlaneData", "Reflective", "SceneData", "SphereData",
           "TriangleData", "RectangleData", "BoxData", "WorkUnit", "WorkUnitResult", "work_units"]
__version__ = "0.1.0"
