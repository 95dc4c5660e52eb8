"""The reference's parameter keys (config/spot.yaml + launch/spot.launch:4-6) -> PipelineConfig."""
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_default_config_matches_the_reference_defaults(ilsm):
    c = ilsm.config.load()
    assert (c.image_width, c.image_height, c.minimum_range) == (1024, 64, 0.3)
    assert (c.mapping_line_resolution, c.mapping_plane_resolution, c.mapping_skip_frame) == (0.4, 0.8, 1)
    assert (c.sliding_window_size, c.ground_plane_window_size) == (0, 2)
    assert c.cloud_topic == "/os_cloud_node/points"


def test_overrides_and_launch_params(ilsm, tmp_path):
    y = tmp_path / "spot.yaml"
    y.write_text("intensity_feature_tracker:\n  image_width: 2048  # wide\n  num_threads: 16\n"
                 "map_optimization_parameters:\n  remove_radius: 0.5\n")
    l = tmp_path / "spot.launch"
    l.write_text('<launch>\n  <param name="mapping_line_resolution" type="double" value="0.2"/>\n'
                 '  <param name="mapping_skip_frame" type="int" value="2" />\n</launch>\n')
    c = ilsm.config.load(str(y), str(l))
    assert c.image_width == 2048 and c.minimum_range == 0.5 and c.mapping_line_resolution == 0.2 and c.mapping_skip_frame == 2
    assert c.mapping_plane_resolution == 0.8  # nh.param default (laserMapping.cpp:1183)
    assert c.raw["intensity_feature_tracker"]["num_threads"] == 16  # keys the path does not use are kept, not rejected


def test_unsupported_values_are_rejected(ilsm, tmp_path):
    y = tmp_path / "bad.yaml"
    y.write_text("intensity_feature_tracker:\n  image_height: 128\n")
    with pytest.raises(ValueError):
        ilsm.config.load(str(y))


@pytest.mark.gpu
def test_pipeline_from_config(ctx, ilsm):
    cfg = ilsm.config.load()
    slam = ilsm.config.make_slam(ctx, cfg, cube_capacity=2048)
    c = ilsm.synth.config1(n_map=20_000)
    qo, to, qm, tm, st = slam.frame(c["cloud"])
    assert st.n_cloud > 10000 and st.n_less_flat > 500
    slam.close()
