#!/usr/bin/env python
"""Generate tests/golden/rectify_cv2.npz: outputs of the GENUINE OpenCV code (Python cv2) for the
two calls of the EuRoC driver that sit in front of track() -- cv2.initUndistortRectifyMap(...,
CV_32FC1) and cv2.remap(..., INTER_LINEAR) (examples/euroc/euroc_example.cpp:106-107,142-143).
Small cases are stored in full; the full-size EuRoC case as SHA-1 digests of the map / image bytes.
Run in the build container (cv2 4.13.0); the fixtures travel with the repo, cv2 does not have to."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def cases():
    """(name, K, D, R, P, (w, h), image seed)"""
    from lvt_b200 import euroc_calib as ec
    out = []
    for side, (K, D, R, P) in zip(("left", "right"), ec.rectify_args(0.25)):
        out.append(("euroc_quarter_" + side, K, D, R, P, (188, 120), 11))
    # strong radial + tangential + k3, principal point shifted: many taps fall outside the raw image
    K = np.array([[90.0, 0, 47.5], [0, 88.0, 41.0], [0, 0, 1]])
    D = np.array([-0.35, 0.12, 0.004, -0.003, -0.02])
    a = 0.05
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]) @ np.array(
        [[1, 0, 0], [0, np.cos(0.03), -np.sin(0.03)], [0, np.sin(0.03), np.cos(0.03)]])
    P = np.array([[70.0, 0, 40.0], [0, 70.0, 44.0], [0, 0, 1]])
    out.append(("strong", K, D, R, P, (96, 80), 12))
    # identity: the maps are the pixel grid, remap must reproduce the image
    K = np.array([[50.0, 0, 20.0], [0, 50.0, 15.0], [0, 0, 1]])
    out.append(("identity", K, np.zeros(5), np.eye(3), K, (40, 30), 13))
    return out


def test_image(seed, w, h):
    import cv2
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w)).astype(np.uint8)
    return cv2.GaussianBlur(img, (0, 0), 1.2)


def main():
    import cv2
    from lvt_b200 import euroc_calib as ec
    data = {"cv2_version": np.array(cv2.__version__)}
    names = []
    for name, K, D, R, P, (w, h), seed in cases():
        m1, m2 = cv2.initUndistortRectifyMap(K, D, R, P, (w, h), cv2.CV_32FC1)
        img = test_image(seed, w, h)
        data[name + "_K"], data[name + "_D"], data[name + "_R"], data[name + "_P"] = K, D, R, P
        data[name + "_map_x"], data[name + "_map_y"] = m1, m2
        data[name + "_img"] = img
        data[name + "_out"] = cv2.remap(img, m1, m2, cv2.INTER_LINEAR)
        names.append(name)
    data["names"] = np.array(names)
    # full-size EuRoC: digests only
    w, h = ec.IMG_SIZE
    for side, (K, D, R, P) in zip(("left", "right"), ec.rectify_args()):
        m1, m2 = cv2.initUndistortRectifyMap(K, D, R, P, (w, h), cv2.CV_32FC1)
        img = test_image(21, w, h)
        out = cv2.remap(img, m1, m2, cv2.INTER_LINEAR)
        data["euroc_full_%s_sha1" % side] = np.array([hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()
                                                      for a in (m1, m2, out)])
    data["euroc_full_img_seed"] = np.int32(21)
    np.savez_compressed(os.path.join(OUT, "rectify_cv2.npz"), **data)
    print("written", os.path.join(OUT, "rectify_cv2.npz"))





TUM_FR1_LIKE = dict(k1=0.2312, k2=-0.7849, p1=-0.0033, p2=-0.0001, k3=0.9172)  # examples/tum_rgbd/config_tum1.yaml:7-11 shape


def undistort_fixture():
    """tests/golden/undistort_cv2.npz: cv2.undistortPoints(kps, K, dist, None, None, K) -- the call of the RGB-D
    path (lvt/src/lvt_image_features_handler.cpp:268-294) -- on the corners the oracle extracts from frame 0 of
    the synthetic RGB-D stream (seed 2), with a TUM fr1-like distortion.  K and dist are the float parameters of
    lvt_parameters widened to double, as the reference passes them."""
    import subprocess
    import cv2
    from lvt_b200 import capi, configs, synth
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    orc = capi.Library(os.path.join(ROOT, "oracle", "_build", "liblvt_oracle.so"))
    p = configs.make_params("tum_synth", max_keypoints_per_cell=400, **TUM_FR1_LIKE)
    st = synth.RgbdStream(n_frames=1, seed=2, **configs.CONFIGS["tum_synth"]["stream"])
    gray, _ = st.frame(0)
    k, _ = orc.context(p).extract(gray)
    K = np.array([[p.fx, 0, p.cx], [0, p.fy, p.cy], [0, 0, 1]], np.float64)
    dist = np.array([p.k1, p.k2, p.p1, p.p2, p.k3], np.float64)
    src = np.stack([k["x"], k["y"]], 1).astype(np.float32)
    corners = np.array([[0, 0], [p.img_width, 0], [0, p.img_height], [p.img_width, p.img_height]], np.float32)
    und = cv2.undistortPoints(src.reshape(-1, 1, 2), K, dist, None, None, K).reshape(-1, 2)
    und_c = cv2.undistortPoints(corners.reshape(-1, 1, 2), K, dist, None, None, K).reshape(-1, 2)
    np.savez_compressed(os.path.join(OUT, "undistort_cv2.npz"), detected=src, undistorted=und, corners=corners,
                        corners_undistorted=und_c, cv2_version=np.array(cv2.__version__))
    print("written", os.path.join(OUT, "undistort_cv2.npz"), len(src), "points")


if __name__ == "__main__":
    if "--undistort" in sys.argv:
        undistort_fixture()
    else:
        main()
