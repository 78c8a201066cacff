"""Dataset drivers and trajectory writers: the three example executables of the reference
(examples/kitti/kitti_example.cpp, examples/euroc/euroc_example.cpp,
examples/tum_rgbd/tum_rgbd_example.cpp) over the C ABI, for whichever library is bound (the CUDA
product or, in tests, the CPU oracle).  File decoding stays on the host as in the reference
(cv::imread -> Python cv2, imported lazily); everything from the raw pixels on runs behind
lvt_track / lvt_track_rgbd, including EuRoC's rectification (lvt_set_rectification)."""
import os
import re

import numpy as np

from . import capi, euroc_calib


def _cv2():
    import cv2
    return cv2


def _gray(img):
    """cv::cvtColor(img, COLOR_BGR2GRAY) when the file is not single-channel (kitti_example.cpp:119-126,
    tum_rgbd_example.cpp:128-129)."""
    if img.ndim == 3:
        img = _cv2().cvtColor(img, _cv2().COLOR_BGR2GRAY)
    return np.ascontiguousarray(img, np.uint8)


def _imread(path):
    img = _cv2().imread(path, _cv2().IMREAD_UNCHANGED)
    if img is None:
        raise IOError("failed to load image %s" % path)
    return img


# ---- trajectory writers --------------------------------------------------------------------------
def quat_from_matrix(R):
    """Eigen::Quaterniond(Matrix3d) (used at euroc_example.cpp:153-154): returns (x, y, z, w)."""
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0:
        t = np.sqrt(t + 1.0)
        w = 0.5 * t
        t = 0.5 / t
        return ((R[2, 1] - R[1, 2]) * t, (R[0, 2] - R[2, 0]) * t, (R[1, 0] - R[0, 1]) * t, w)
    i = 0
    if R[1, 1] > R[0, 0]:
        i = 1
    if R[2, 2] > R[i, i]:
        i = 2
    j, k = (i + 1) % 3, (i + 2) % 3
    t = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
    q = [0.0, 0.0, 0.0]
    q[i] = 0.5 * t
    t = 0.5 / t
    w = (R[k, j] - R[j, k]) * t
    q[j] = (R[j, i] + R[i, j]) * t
    q[k] = (R[k, i] + R[i, k]) * t
    return (q[0], q[1], q[2], w)


def dump_kitti_trajectory(path, poses):
    """12 numbers per line, %.9f: rows of [R | t] (kitti_example.cpp:33-47)."""
    with open(path, "w") as f:
        for R, t in poses:
            f.write(" ".join("%.9f" % v for v in (R[0, 0], R[0, 1], R[0, 2], t[0], R[1, 0], R[1, 1], R[1, 2], t[1],
                                                  R[2, 0], R[2, 1], R[2, 2], t[2])) + "\n")


def dump_tum_trajectory(path, poses, stamps):
    """timestamp %.6f, position and quaternion (x y z w) %.7f (euroc_example.cpp:34-47,
    tum_rgbd_example.cpp:34-47)."""
    with open(path, "w") as f:
        for (R, t), ts in zip(poses, stamps):
            q = quat_from_matrix(np.asarray(R))
            f.write("%.6f %.7f %.7f %.7f %.7f %.7f %.7f %.7f\n" % (ts, t[0], t[1], t[2], q[0], q[1], q[2], q[3]))


def _identity():
    return np.eye(3), np.zeros(3)


# ---- KITTI odometry (kitti_example.cpp:49-160) -----------------------------------------------------
def read_kitti_calib(path):
    """camera_matrix (3x3) and baseline of examples/kitti/calib/NN.yml"""
    txt = open(path).read()
    m = re.search(r"camera_matrix:.*?data:\s*\[([^\]]*)\]", txt, re.S)
    b = re.search(r"baseline:\s*([-+0-9.eE]+)", txt)
    if not m or not b:
        raise ValueError("no camera_matrix / baseline in %s" % path)
    K = np.array([float(v) for v in m.group(1).replace("\n", " ").split(",")]).reshape(3, 3)
    return K, float(b.group(1))


def run_kitti(lib, sequences_dir, seq, config_yaml, calib_yml, out_path=None, max_frames=None):
    seq_dir = os.path.join(sequences_dir, "%02d" % int(seq))
    left = sorted(f for f in os.listdir(os.path.join(seq_dir, "image_0")) if f.endswith(".png"))
    if max_frames:
        left = left[:max_frames]
    params = lib.params_from_file(config_yaml)
    K, baseline = read_kitti_calib(calib_yml)
    first = _gray(_imread(os.path.join(seq_dir, "image_0", left[0])))
    params.fx, params.fy, params.cx, params.cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    params.baseline = baseline
    params.img_height, params.img_width = first.shape
    vo = lib.create(params, capi.SENSOR_STEREO)
    poses = [_identity() for _ in left]
    for i, name in enumerate(left):
        a = _gray(_imread(os.path.join(seq_dir, "image_0", name)))
        b = _gray(_imread(os.path.join(seq_dir, "image_1", name)))
        poses[i] = vo.track(a, b)
        if vo.get_state() == capi.STATE_LOST:
            break
    vo.destroy()
    if out_path:
        dump_kitti_trajectory(out_path, poses)
    return poses


# ---- EuRoC MAV (euroc_example.cpp:49-160) ----------------------------------------------------------
def run_euroc(lib, root_dir, stamps_dir, dataset_name, config_yaml, out_path=None, max_frames=None):
    seq_dir = os.path.join(root_dir, dataset_name, "mav0")
    titles, stamps = [], []
    for line in open(os.path.join(stamps_dir, dataset_name + ".txt")):
        line = line.strip()
        if not line:
            continue
        titles.append(line + ".png")
        stamps.append(float(line) / 1e9)
    if max_frames:
        titles, stamps = titles[:max_frames], stamps[:max_frames]
    params = lib.params_from_file(config_yaml)
    params.fx, params.fy, params.cx, params.cy = euroc_calib.FX, euroc_calib.FY, euroc_calib.CX, euroc_calib.CY
    params.baseline = euroc_calib.BASELINE
    vo = None
    poses = [_identity() for _ in titles]
    for i, name in enumerate(titles):
        a = _gray(_imread(os.path.join(seq_dir, "cam0", "data", name)))
        b = _gray(_imread(os.path.join(seq_dir, "cam1", "data", name)))
        if vo is None:
            params.img_height, params.img_width = a.shape
            vo = lib.create(params, capi.SENSOR_STEREO)
            # the raw frames go in; initUndistortRectifyMap + remap happen behind lvt_track
            scale = a.shape[1] / float(euroc_calib.IMG_SIZE[0])
            (Kl, Dl, Rl, Pl), (Kr, Dr, Rr, Pr) = euroc_calib.rectify_args(scale)
            vo.set_rectification(capi.Rectify.make(Kl, Dl, Rl, Pl), capi.Rectify.make(Kr, Dr, Rr, Pr))
        R, t = vo.track(a, b)
        cam = np.eye(4)
        cam[:3, :3], cam[:3, 3] = R, t
        body = euroc_calib.T_BS @ cam  # euroc_example.cpp:150-154
        poses[i] = (body[:3, :3].copy(), body[:3, 3].copy())
        if vo.get_state() == capi.STATE_LOST:
            break
    if vo is not None:
        vo.destroy()
    if out_path:
        dump_tum_trajectory(out_path, poses, stamps)
    return poses, stamps


# ---- TUM RGB-D (tum_rgbd_example.cpp:49-145) -------------------------------------------------------
def run_tum_rgbd(lib, root_dir, associations_dir, dataset_name, config_yaml, out_path=None, max_frames=None):
    stamps, rgb, depth = [], [], []
    for line in open(os.path.join(associations_dir, dataset_name + ".txt")):
        f = line.split()
        if len(f) < 4:
            continue
        stamps.append(float(f[0]))
        rgb.append(f[1])
        depth.append(f[3])
    if max_frames:
        stamps, rgb, depth = stamps[:max_frames], rgb[:max_frames], depth[:max_frames]
    if not rgb:
        raise ValueError("image associations were not read correctly")
    params = lib.params_from_file(config_yaml)
    vo = lib.create(params, capi.SENSOR_RGBD)
    poses = [_identity() for _ in rgb]
    depth_scale = np.float32(1.0 / 5000.0)  # tum_rgbd_example.cpp:111
    for i in range(len(rgb)):
        g = _gray(_imread(os.path.join(root_dir, dataset_name, rgb[i])))
        d = _imread(os.path.join(root_dir, dataset_name, depth[i])).astype(np.float32) * depth_scale
        poses[i] = vo.track_rgbd(g, d)
        if vo.get_state() == capi.STATE_LOST:
            break
    vo.destroy()
    if out_path:
        dump_tum_trajectory(out_path, poses, stamps)
    return poses, stamps


def main(argv=None):
    import argparse
    import lvt_b200
    ap = argparse.ArgumentParser(description="run a dataset through the B200 track() path")
    sub = ap.add_subparsers(dest="cmd", required=True)
    k = sub.add_parser("kitti")
    k.add_argument("sequences_dir"), k.add_argument("seq"), k.add_argument("config"), k.add_argument("calib")
    e = sub.add_parser("euroc")
    e.add_argument("root_dir"), e.add_argument("stamps_dir"), e.add_argument("dataset"), e.add_argument("config")
    t = sub.add_parser("tum")
    t.add_argument("root_dir"), t.add_argument("associations_dir"), t.add_argument("dataset"), t.add_argument("config")
    for p in (k, e, t):
        p.add_argument("--out", default=None)
        p.add_argument("--max-frames", type=int, default=None)
    a = ap.parse_args(argv)
    lib = lvt_b200.load()
    if a.cmd == "kitti":
        run_kitti(lib, a.sequences_dir, a.seq, a.config, a.calib, a.out or "%02d.txt" % int(a.seq), a.max_frames)
    elif a.cmd == "euroc":
        run_euroc(lib, a.root_dir, a.stamps_dir, a.dataset, a.config, a.out or a.dataset + ".txt", a.max_frames)
    else:
        run_tum_rgbd(lib, a.root_dir, a.associations_dir, a.dataset, a.config, a.out or a.dataset + ".txt", a.max_frames)


if __name__ == "__main__":
    main()
