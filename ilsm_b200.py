"""Import alias: the package directory name contains a hyphen (it mirrors the reference repository's name), so
`import ilsm_b200` is the spelling Python code uses."""
import importlib
import sys

import os

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
_pkg = importlib.import_module("intensity_based_lidar_slam_for_me-_b200")
sys.modules[__name__] = _pkg
