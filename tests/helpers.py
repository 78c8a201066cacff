import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from lvt_b200 import capi, configs, synth  # noqa: E402


def kp_array(k):
    return np.stack([k["x"], k["y"], k["response"]], 1).astype(np.float32).reshape(-1, 3)


def kp_equal(a, b):
    return len(a) == len(b) and np.array_equal(kp_array(a), kp_array(b))


def make_stream(name, n_frames, seed=0):
    cfg = configs.CONFIGS[name]
    if cfg["sensor"] == 1:
        return synth.StereoStream(n_frames=n_frames, seed=seed, **cfg["stream"])
    return synth.RgbdStream(n_frames=n_frames, seed=seed, **cfg["stream"])


def track(vo, sensor, a, b):
    return vo.track(a, b) if sensor == 1 else vo.track_rgbd(a, b)


def knn_params(lib, **kw):
    """parameters under which find_match_index returns the raw best match (tests of top-2)"""
    p = configs.make_params("kitti_synth", tracking_ratio_test_threshold=1e9, descriptor_matching_threshold=1e9, **kw)
    return p


def knn_via_match(ctx, p, train, query, mask):
    """drive masked top-2 through lvtk_match_projected: all features at one pixel, mask as marks"""
    n = len(train)
    kps = np.zeros(n, capi.KP_DTYPE)
    kps["x"] = 100.0
    kps["y"] = 100.0
    Z = 10.0
    pt = np.array([[(100.0 - p.cx) / p.fx * Z, (100.0 - p.cy) / p.fy * Z, Z]])
    r = ctx.match_projected(pt, query.reshape(1, 32), [1, 0, 0, 0], [0, 0, 0], kps, train, matched=1 - mask, retry_below=0)
    return int(r["idx"][0]), float(r["d1"][0]), float(r["d2"][0])
