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


def test_yaml_parameters_equal_cv2_filestorage(oracle, tmp_path):
    """lvt_parameters::init_from_file reads its YAML with cv::FileStorage (lvt/src/lvt_parameters.cpp:54-93); the
    restated reader must give every field the value the genuine FileStorage gives (floats as float(node), ints as
    int(node), absent keys 0)"""
    import ctypes as C
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    text = ("%YAML:1.0\n\n# a comment line\nfx: 517.306408\nfy: 5.16469215e+02\ncx: 318.643040\ncy: 255.313989\n"
            "k1: 0.262383\nk2: -0.953104\np1: -0.005358\np2: 2.628e-3\nk3: 1.163314\nbaseline: .5372\n"
            "img_width: 640\nimg_height: 480\nnear_plane_distance: 0.1\nfar_plane_distance: 5.0\n"
            "triangulation_ratio_test_threshold: 0.60\ntracking_ratio_test_threshold: 0.70\n"
            "descriptor_matching_threshold: 30\nmin_num_matches_for_tracking: 10\ntracking_radius: 30\n"
            "agast_threshold: 18\ndetection_cell_size: 2000\nmax_keypoints_per_cell: 1000   # trailing comment\n"
            "untracked_threshold: 10\nstaged_threshold: 0\nenable_logging: 0\nenable_visualization: 1\n"
            "triangulation_policy: 2\nviewer_camera_size: 0.06\nviewer_point_size: 2\nunknown_key: 7\n")
    files = [text]
    for _ in range(3):  # random values in the same layout
        lines = ["%YAML:1.0", ""]
        for k, t in capi.Params._fields_:
            v = int(rng.integers(0, 3000)) if t is C.c_int else float(np.round(rng.uniform(-2, 900), int(rng.integers(0, 7))))
            lines.append("%s: %r" % (k, v))
        files.append("\n".join(lines) + "\n")
    for i, body in enumerate(files):
        f = tmp_path / ("cfg%d.yaml" % i)
        f.write_text(body)
        fs = cv2.FileStorage(str(f), cv2.FILE_STORAGE_READ)
        p = oracle.params_from_file(str(f))
        for k, t in capi.Params._fields_:
            node = fs.getNode(k)
            want = 0.0 if node.empty() else node.real()
            got = getattr(p, k)
            if k in ("enable_logging", "enable_visualization"):  # bool members: (int)node converted to bool
                assert got == int(int(want) != 0), (i, k, got, want)
            elif t is C.c_int:
                assert got == int(want), (i, k, got, want)
            else:
                assert np.float32(got) == np.float32(want), (i, k, got, want)
        fs.release()
