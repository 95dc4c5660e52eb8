"""The reference's parameters for this path, read from the same keys its nodes read from the ROS parameter server
(config/spot.yaml + launch/spot.launch:4-6).  load() accepts the reference's own spot.yaml unchanged (extra keys are kept
in `raw`, missing keys take the reference's nh.param defaults) and, optionally, its spot.launch for the three launch
params."""
from __future__ import annotations

import dataclasses
import os
import re

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_YAML = os.path.join(ROOT, "config", "spot.yaml")


@dataclasses.dataclass
class PipelineConfig:
    image_width: int = 1024            # /intensity_feature_tracker/image_width    mapOptimization.cpp:522
    image_height: int = 64             # /intensity_feature_tracker/image_height   scanRegistration.cpp:692 (N_SCANS)
    minimum_range: float = 0.3         # /map_optimization_parameters/remove_radius scanRegistration.cpp:695
    mapping_line_resolution: float = 0.4   # spot.launch:4, laserMapping.cpp:1181
    mapping_plane_resolution: float = 0.8  # spot.launch:5, laserMapping.cpp:1183
    mapping_skip_frame: int = 1        # spot.launch:6, laserOdometry.cpp:265
    sliding_window_size: int = 0       # mapOptimization.cpp:538
    ground_plane_window_size: int = 2  # mapOptimization.cpp:541
    cloud_topic: str = "/os_cloud_node/points"
    # hard-coded in the reference (not parameters): kept here so that one object describes the path
    voxel_leaf: float = 0.8            # mapOptimization.cpp:578 voxel_grid_ leaf
    downsample_size: float = 0.4       # mapOptimization.cpp:504 ikd-Tree box
    raw: dict = dataclasses.field(default_factory=dict, repr=False)

    def validate(self):
        if self.image_height != 64:
            raise ValueError("only the 64-ring branch of scanRegistration.cpp:308-316 is implemented (image_height must be 64)")
        if self.image_width <= 0 or not (self.mapping_line_resolution > 0 and self.mapping_plane_resolution > 0):
            raise ValueError("image_width and the mapping resolutions must be positive")
        if not self.minimum_range >= 0:
            raise ValueError("remove_radius must be >= 0")
        return self


def _launch_params(path):
    """<param name="..." value="..."/> entries of a roslaunch file."""
    txt = open(path).read()
    return {m.group(1): m.group(2) for m in re.finditer(r'<param\s+name="([^"]+)"[^>]*?\svalue="([^"]*)"', txt)}


def load(yaml_path: str | None = None, launch_path: str | None = None) -> PipelineConfig:
    y = yaml.safe_load(open(yaml_path or DEFAULT_YAML)) or {}
    ift = y.get("intensity_feature_tracker", {}) or {}
    mo = y.get("map_optimization_parameters", {}) or {}
    c = PipelineConfig(raw=y)
    c.image_width = int(ift.get("image_width", c.image_width))
    c.image_height = int(ift.get("image_height", c.image_height))
    c.cloud_topic = str(ift.get("cloud_topic", c.cloud_topic))
    c.minimum_range = float(mo.get("remove_radius", c.minimum_range))
    c.sliding_window_size = int(mo.get("sliding_window_size", c.sliding_window_size))
    c.ground_plane_window_size = int(mo.get("ground_plane_window_size", c.ground_plane_window_size))
    top = dict(y)
    if launch_path:
        top.update(_launch_params(launch_path))
    c.mapping_line_resolution = float(top.get("mapping_line_resolution", c.mapping_line_resolution))
    c.mapping_plane_resolution = float(top.get("mapping_plane_resolution", c.mapping_plane_resolution))
    c.mapping_skip_frame = int(top.get("mapping_skip_frame", c.mapping_skip_frame))
    return c.validate()


def make_slam(ctx, cfg: PipelineConfig, mapping: str = "laserMapping", cube_capacity: int = 0, pipelined: bool = False):
    """The launched pipeline configured from the reference's parameters."""
    from .binding import Slam
    return Slam(ctx, cfg.mapping_line_resolution, cfg.mapping_plane_resolution, cfg.minimum_range, cube_capacity, mapping=mapping,
                voxel_leaf=cfg.voxel_leaf, downsample_size=cfg.downsample_size, pipelined=pipelined)
