"""The BASELINE.json workloads as parameter sets + stream recipes."""
from .capi import Params

_KITTI = dict(  # examples/kitti/vo_config.yaml:3-21 + examples/kitti/calib/00.yml:7,9
    near_plane_distance=0.01, far_plane_distance=500.0, triangulation_ratio_test_threshold=0.60,
    tracking_ratio_test_threshold=0.80, descriptor_matching_threshold=30.0, min_num_matches_for_tracking=10,
    tracking_radius=25, agast_threshold=25, detection_cell_size=250, max_keypoints_per_cell=150,
    untracked_threshold=10, staged_threshold=2, enable_logging=0, enable_visualization=0,
    triangulation_policy=1, viewer_camera_size=0.6, viewer_point_size=5,
    fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, baseline=0.53716571886, img_width=1242, img_height=375)

_EUROC = dict(  # examples/euroc/vo_config_euroc.yaml + examples/euroc/euroc_example.cpp:109-113
    near_plane_distance=0.01, far_plane_distance=500.0, triangulation_ratio_test_threshold=0.60,
    tracking_ratio_test_threshold=0.70, descriptor_matching_threshold=30.0, min_num_matches_for_tracking=10,
    tracking_radius=25, agast_threshold=20, detection_cell_size=250, max_keypoints_per_cell=100,
    untracked_threshold=10, staged_threshold=0, enable_logging=0, enable_visualization=0,
    triangulation_policy=1, viewer_camera_size=0.2, viewer_point_size=2,
    fx=435.2046959714599, fy=435.2046959714599, cx=367.4517211914062, cy=252.2008514404297,
    baseline=0.110077842, img_width=752, img_height=480)

_TUM3 = dict(  # examples/tum_rgbd/config_tum3.yaml
    fx=535.4, fy=539.2, cx=320.1, cy=247.6, img_width=640, img_height=480, near_plane_distance=0.1,
    far_plane_distance=5.0, triangulation_ratio_test_threshold=0.60, tracking_ratio_test_threshold=0.70,
    descriptor_matching_threshold=30.0, min_num_matches_for_tracking=10, tracking_radius=30, agast_threshold=18,
    detection_cell_size=2000, max_keypoints_per_cell=1000, untracked_threshold=10, staged_threshold=0,
    enable_logging=0, enable_visualization=0, triangulation_policy=2, viewer_camera_size=0.06, viewer_point_size=2)

CONFIGS = {
    # config 2 / 4: 1242x375 synthetic stereo, ~2000 keypoints / frame
    "kitti_synth": dict(sensor=1, params=dict(_KITTI, max_keypoints_per_cell=250),
                        stream=dict(W=1242, H=375, disparity=20, step=16, density=180, noise=2.0)),
    # the stock KITTI parameters (k = 150)
    "kitti_stock": dict(sensor=1, params=dict(_KITTI),
                        stream=dict(W=1242, H=375, disparity=20, step=16, density=180, noise=2.0)),
    # config 5: 752x480 EuRoC shape, ~5000 keypoints / frame
    "euroc_synth": dict(sensor=1, params=dict(_EUROC, max_keypoints_per_cell=1020, agast_threshold=8),
                        stream=dict(W=752, H=480, disparity=12, step=12, density=60, noise=6.0)),
    # config 3: 640x480 RGB-D TUM shape, ~1500 keypoints / frame
    "tum_synth": dict(sensor=2, params=dict(_TUM3, max_keypoints_per_cell=1860),
                      stream=dict(W=640, H=480, step=8, density=180, noise=2.0)),
}


def make_params(name, **overrides):
    p = Params()
    for k, v in dict(CONFIGS[name]["params"], **overrides).items():
        setattr(p, k, v)
    return p
