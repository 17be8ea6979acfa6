"""B200-native multi-view geometry hot path of SmartEdgeSensor3DHumanPose (association, DLT triangulation with
unscented covariance, plausibility/merge, reprojection) behind the C ABI of include/ses3d.h.

    from smartedgesensor3dhumanpose_b200 import GeometryPipeline, Skeleton3D, PoseReprojection

`api` mirrors the reference's entry points, `layouts` the person_msgs PODs, `rigs` / `synth` provide camera rigs and
the synthetic frame generator, `assembler` / `wire` the live-replay helpers, `sharding` the multi-GPU frame split.
The CUDA library (libses3d.so) is built with `python -m smartedgesensor3dhumanpose_b200.build`; there is no CPU path.
"""
from .layouts import default_params  # noqa: F401

__version__ = "0.2.0"
__all__ = ["GeometryPipeline", "Skeleton3D", "PoseReprojection", "PosePrior", "PriorTracker", "default_params"]


def __getattr__(name):  # the API classes load the shared library on first use, not at import
    if name in ("GeometryPipeline", "Skeleton3D", "PoseReprojection", "PosePrior", "PriorTracker"):
        from . import api
        return getattr(api, name)
    raise AttributeError(name)
