"""3dgs_rigidbody_b200 -- B200-native (sm_100a) animate -> project -> tile-sort -> composite hot path of
JTStephens18/3DGS_rigidbody behind the reference's own operator / API boundary.

The directory name starts with a digit, so import it with
    rs = importlib.import_module("3dgs_rigidbody_b200")
(or `import rigidsplat`, the alias module at the repo root).

Public surface (mirrors the reference):
    rasterization                      gsplat/rendering.py:33-770   (+ cluster_ids / body_* kwargs)
    fully_fused_projection, isect_tiles, isect_offset_encode, rasterize_to_pixels   gsplat/cuda/_wrapper.py
    _C                                 stand-in for the pybind module gsplat/cuda/ext.cpp (same names / positional args)
    RigidPoses, FrameRenderer          rigid-pose table; sync-free fused per-frame renderer (animation loop)
    cgc_contrastive_clustering_loss    examples/utils.py:828-904 (identity-feature training step)
    SegmentationHead                   examples/simple_trainer.py:442-446 (fused 16 -> 64 -> 16 head of the same step)
    load_cluster_groups, body_properties, PoseStream   clustering / physics data contract (rigid.py)
"""
from . import _C  # noqa: F401
from ._C import RigidPoses, RigidSplatError  # noqa: F401
from .animation import FramePipeline, FrameRenderer  # noqa: F401
from .rendering import rasterization  # noqa: F401
from .identity import (  # noqa: F401
    SegmentationHead,
    cgc_contrastive_clustering_loss,
    cluster_tables,
    segmentation_head_forward,
)
from .io import load_ply  # noqa: F401
from .rigid import (  # noqa: F401
    PoseStream,
    body_centers,
    body_properties,
    cluster_ids_from_groups,
    load_cluster_groups,
    make_rigid,
)
from .sh import spherical_harmonics  # noqa: F401


def __getattr__(name):  # torch.distributed-dependent classes are imported on first use
    if name in ("ShardedFrameRenderer", "PeerSplatExchange", "GaussianShardExchange"):
        from . import distributed as _d

        return getattr(_d, name)
    raise AttributeError(name)
from .wrapper import (  # noqa: F401
    fully_fused_projection,
    isect_offset_encode,
    isect_tiles,
    rasterize_to_pixels,
)

__version__ = "0.1.0"
