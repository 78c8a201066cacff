"""Pin the CPU oracle against the genuine OpenCV code: committed cv2 fixtures (always) and, when
cv2 is importable, freshly computed cv2 outputs on more inputs."""
import numpy as np
import pytest

from helpers import GOLDEN, kp_array, knn_params, knn_via_match, configs, synth, capi


def test_agast_matches_cv2_fixtures(oracle):
    g = np.load(GOLDEN + "/agast_cv2.npz")
    ctx = oracle.context(configs.make_params("kitti_synth"))
    for i in range(int(g["n_cases"])):
        img, th = g["img%d" % i], int(g["th%d" % i])
        for nms in (0, 1):
            got = kp_array(ctx.agast(img, th, bool(nms)))
            assert np.array_equal(got, g["kp%d_nms%d" % (i, nms)]), (i, nms)


def test_knn_matches_cv2_fixtures(oracle):
    g = np.load(GOLDEN + "/knn_cv2.npz")
    p = knn_params(oracle)
    ctx = oracle.context(p)
    for q, m, r in zip(g["queries"], g["masks"], g["results"]):
        idx, d0, d1 = knn_via_match(ctx, p, g["train"], q, m)
        n = int((r[:, 0] >= 0).sum())
        if n == 0:
            assert idx == -1
        elif n == 1:
            assert idx == int(r[0, 0]) and d0 == r[0, 1] and d1 == -1.0
        elif r[0, 1] == 0 and r[1, 1] == 0:
            assert idx == -1  # 0/0 = NaN fails the ratio test
        else:
            assert idx == int(r[0, 0]) and d0 == r[0, 1] and d1 == r[1, 1]


def test_agast_matches_live_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    ctx = oracle.context(configs.make_params("kitti_synth"))
    rng = np.random.default_rng(99)
    imgs = [rng.integers(0, 256, (90, 110)).astype(np.uint8), (rng.integers(0, 6, (80, 80)) * 40).astype(np.uint8),
            synth.canvas(3, 250, 250), synth.canvas(4, 125, 242), synth.canvas(5, 230, 250, density=60, noise=6.0)]
    for img in imgs:
        for th in (8, 20, 38):
            for nms in (False, True):
                det = cv2.AgastFeatureDetector_create(th, nms, cv2.AgastFeatureDetector_OAST_9_16)
                ref = np.array([(k.pt[0], k.pt[1], k.response) for k in det.detect(img)], np.float32).reshape(-1, 3)
                assert np.array_equal(kp_array(ctx.agast(img, th, nms)), ref)


def test_sliver_and_tiny_tiles(oracle):
    ctx = oracle.context(configs.make_params("kitti_synth"))
    rng = np.random.default_rng(1)
    for shape in ((250, 2), (2, 250), (6, 50), (2, 2)):
        assert len(ctx.agast(rng.integers(0, 256, shape).astype(np.uint8), 5, True)) == 0
