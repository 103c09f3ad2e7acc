"""octree-slam_b200: B200-native (sm_100a) hot path of Octree-SLAM -- depth -> SVO integration and SVO raycast.

The directory name contains a hyphen (the layout the task prescribes); import it through
`__graft_entry__.load_package()` which registers it as `octree_slam_b200`.
"""
from . import capi  # noqa: F401
from .capi import Counters, OslError, RaycastParams, RaycastStats, lib  # noqa: F401
from .world import (SVO, BoundingBox, Octree, Scene, computeKeys, computePointCloudBoundingBox,  # noqa: F401
                    coneTraceSVO, generateVertexMap, meshToVoxelGrid, meshToVoxelGridThin, transformVertexMap)
from . import sensor  # noqa: F401
from .sensor import RGBDCamera  # noqa: F401
from . import synth  # noqa: F401
from . import shard  # noqa: F401
