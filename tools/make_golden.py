#!/usr/bin/env python
"""Generate tests/golden/*.npz: outputs of the GENUINE OpenCV code (Python cv2) for the two
third-party calls on the hot path that cv2 exposes -- AgastFeatureDetector (OAST_9_16) and
BFMatcher(NORM_HAMMING).knnMatch(k=2, mask) -- on small seeded inputs.  Run in the build
container (cv2 4.13.0); the fixtures travel with the repo, cv2 does not have to.

Also records the CPU oracle's own trajectory on a short synthetic stereo stream so that both the
oracle (regression) and the CUDA path (parity) can be checked against a committed file.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def agast_cases():
    rng = np.random.default_rng(2018)
    cases = []
    for i, (h, w, th, kind) in enumerate([(60, 80, 20, "noise"), (72, 96, 12, "blur"), (64, 64, 25, "quant"),
                                          (50, 120, 35, "rects"), (125, 242, 25, "rects"), (7, 7, 5, "noise"),
                                          (250, 2, 10, "noise"), (40, 40, 8, "quant")]):
        if kind == "noise":
            img = rng.integers(0, 256, (h, w)).astype(np.uint8)
        elif kind == "quant":
            img = (rng.integers(0, 8, (h, w)) * 32).astype(np.uint8)
        else:
            from lvt_b200 import synth
            img = synth.canvas(100 + i, h, w, density=60 if kind == "blur" else 180, noise=3.0)
        cases.append((img, th))
    return cases


def main():
    import cv2
    os.makedirs(OUT, exist_ok=True)
    # ---- AGAST ------------------------------------------------------------------------------
    data = {}
    for i, (img, th) in enumerate(agast_cases()):
        for nms in (0, 1):
            det = cv2.AgastFeatureDetector_create(th, bool(nms), cv2.AgastFeatureDetector_OAST_9_16)
            k = det.detect(img)
            data["img%d" % i] = img
            data["th%d" % i] = np.int32(th)
            data["kp%d_nms%d" % (i, nms)] = np.array([(q.pt[0], q.pt[1], q.response) for q in k], np.float32).reshape(-1, 3)
    data["n_cases"] = np.int32(len(agast_cases()))
    data["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(OUT, "agast_cv2.npz"), **data)

    # ---- masked Hamming top-2 ----------------------------------------------------------------
    rng = np.random.default_rng(7)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING, False)
    train = rng.integers(0, 256, (300, 32)).astype(np.uint8)
    train[10] = train[3]      # duplicates -> distance ties, lower index must win
    train[11] = train[3]
    queries, masks, res = [], [], []
    for c in range(40):
        q = rng.integers(0, 256, (1, 32)).astype(np.uint8)
        if c % 5 == 0:
            q = train[3:4].copy()
            q[0, 0] ^= 1
        if c == 7:
            q = train[3:4].copy()   # d0 = d1 = 0 -> 0/0 = NaN
        m = np.zeros((1, 300), np.uint8)
        n_c = [0, 1, 2, 3, 25, 300][c % 6]
        m[0, rng.permutation(300)[:n_c]] = 1
        if c % 5 == 0 or c == 7:
            m[0, [3, 10, 11]] = 1
        out = bf.knnMatch(q, train, 2, m)[0]
        r = np.full((2, 2), -1.0)
        for j, d in enumerate(out):
            r[j] = (d.trainIdx, d.distance)
        queries.append(q[0])
        masks.append(m[0])
        res.append(r)
    np.savez_compressed(os.path.join(OUT, "knn_cv2.npz"), train=train, queries=np.array(queries), masks=np.array(masks),
                        results=np.array(res), cv2_version=np.array(cv2.__version__))

    # ---- oracle trajectory ---------------------------------------------------------------------
    import subprocess
    from lvt_b200 import capi, configs, synth
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    orc = capi.Library(os.path.join(ROOT, "oracle", "_build", "liblvt_oracle.so"))
    for name, nf in (("kitti_synth", 10), ("euroc_synth", 4)):
        cfg = configs.CONFIGS[name]
        p = configs.make_params(name)
        st = synth.StereoStream(n_frames=nf, seed=0, **cfg["stream"])
        vo = orc.create(p, 1)
        poses, infos, n_desc = [], [], []
        for t in range(nf):
            a, b = st.frame(t)
            R, tt = vo.track(a, b)
            poses.append(np.concatenate([R.ravel(), tt]))
            fi = vo.frame_info()
            infos.append([fi[k] for k in sorted(fi)])
            xy, desc = vo.features(0)
            n_desc.append([len(xy), int(desc.astype(np.uint64).sum()), int(xy.astype(np.float64).sum())])
        np.savez_compressed(os.path.join(OUT, "track_%s.npz" % name), poses=np.array(poses), infos=np.array(infos),
                            info_keys=np.array(sorted(fi)), feature_sums=np.array(n_desc))
        vo.destroy()
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
