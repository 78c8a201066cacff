"""ctypes binding of the C ABI declared in include/lvt_c.h and include/lvt_kernels.h.

The binding is library-agnostic: it is given the path of a shared object exporting that ABI.
The product (`lvt_b200.load()`) binds lvt_b200/lib/liblvt_b200.so; tests bind the CPU oracle
with the same class so that both are driven through identical calls.

Python mirror of the reference interface: `Library.create()` returns a `System` whose methods
carry the names of lvt_system's (lvt/src/lvt_system.h:57-70): track, track_with_external_corners,
reset, get_state.
"""
import ctypes as C
import os

import numpy as np

c_u8p = C.POINTER(C.c_uint8)
c_i32p = C.POINTER(C.c_int)
c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)


class Params(C.Structure):
    """lvt_params_c == struct lvt_parameters (lvt/src/lvt_parameters.h:29-64)."""

    _fields_ = [
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("baseline", C.c_float),
        ("img_width", C.c_int), ("img_height", C.c_int),
        ("k1", C.c_float), ("k2", C.c_float), ("p1", C.c_float), ("p2", C.c_float), ("k3", C.c_float),
        ("near_plane_distance", C.c_float), ("far_plane_distance", C.c_float),
        ("triangulation_ratio_test_threshold", C.c_float),
        ("tracking_ratio_test_threshold", C.c_float),
        ("descriptor_matching_threshold", C.c_float),
        ("min_num_matches_for_tracking", C.c_int),
        ("tracking_radius", C.c_int),
        ("detection_cell_size", C.c_int),
        ("max_keypoints_per_cell", C.c_int),
        ("agast_threshold", C.c_int),
        ("untracked_threshold", C.c_int),
        ("staged_threshold", C.c_int),
        ("enable_logging", C.c_int),
        ("enable_visualization", C.c_int),
        ("triangulation_policy", C.c_int),
        ("viewer_camera_size", C.c_float),
        ("viewer_point_size", C.c_int),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}

    def to_array(self):
        """Pack as float64 (every field is exactly representable) -- the broadcast payload."""
        return np.array([float(getattr(self, k)) for k, _ in self._fields_], dtype=np.float64)

    @classmethod
    def from_array(cls, a):
        p = cls()
        for (k, t), v in zip(cls._fields_, a):
            setattr(p, k, int(round(float(v))) if t is C.c_int else float(v))
        return p


class FrameInfo(C.Structure):
    _fields_ = [(k, C.c_int) for k in (
        "frame_number", "state", "n_features_left", "n_features_right", "map_points_before",
        "staged_before", "tracked", "inliers", "map_points_after", "staged_after", "triangulated",
        "new_points", "retried_matching")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Rectify(C.Structure):
    """lvt_rectify_c: one camera's arguments of cv::initUndistortRectifyMap
    (examples/euroc/euroc_example.cpp:96-107): K, D = k1 k2 p1 p2 k3, R, P(0:3,0:3), row-major."""

    _fields_ = [("K", C.c_double * 9), ("D", C.c_double * 5), ("R", C.c_double * 9), ("P", C.c_double * 9)]

    @classmethod
    def make(cls, K, D, R, P):
        r = cls()
        r.K[:] = [float(v) for v in np.asarray(K, np.float64).reshape(9)]
        r.D[:] = [float(v) for v in np.asarray(D, np.float64).reshape(5)]
        r.R[:] = [float(v) for v in np.asarray(R, np.float64).reshape(9)]
        r.P[:] = [float(v) for v in np.asarray(P, np.float64)[:3, :3].reshape(9)]
        return r


KP_DTYPE = np.dtype([("x", np.float32), ("y", np.float32), ("response", np.float32)])

STATE_NOT_INITIALIZED, STATE_TRACKING, STATE_LOST = 1, 2, 3
SENSOR_STEREO, SENSOR_RGBD = 1, 2


def _u8(a):
    return a.ctypes.data_as(c_u8p)


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class LvtError(RuntimeError):
    pass


_OPTIONAL = ("lvt_track_device", "lvt_track_stream_device", "lvt_track_stream_host", "lvt_set_device",
             "lvtk_last_error", "lvt_get_stage_times")


class Library:
    def __init__(self, path):
        if not os.path.exists(path):
            raise LvtError("shared library not found: %s (run `python -c 'import __graft_entry__ as g; g.build()'`)" % path)
        self.path = path
        self.lib = lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        vp = C.c_void_p
        lib.lvt_create.restype = vp
        lib.lvt_create.argtypes = [C.c_char_p, C.c_int]
        lib.lvt_create_from_params.restype = vp
        lib.lvt_create_from_params.argtypes = [C.POINTER(Params), C.c_int]
        lib.lvt_destroy.argtypes = [vp]
        lib.lvt_reset.argtypes = [vp]
        lib.lvt_track.argtypes = [vp, c_u8p, c_u8p, C.c_int, C.c_int, c_f64p, c_f64p]
        lib.lvt_track_rgbd.argtypes = [vp, c_u8p, c_f32p, C.c_int, C.c_int, c_f64p, c_f64p]
        lib.lvt_track_with_external_corners.argtypes = [vp, c_u8p, c_u8p, C.c_int, C.c_int, c_f64p, C.c_int,
                                                        c_f64p, C.c_int, c_f64p, c_f64p]
        lib.lvt_get_status.argtypes = [vp]
        lib.lvt_get_status.restype = C.c_int
        lib.lvt_debug_point_capacity.argtypes = [vp]
        lib.lvt_debug_point_capacity.restype = C.c_int
        lib.lvt_get_last_status.argtypes = [vp]
        lib.lvt_get_last_status.restype = C.c_int
        lib.lvt_params_default.argtypes = [C.POINTER(Params)]
        lib.lvt_params_from_file.argtypes = [C.POINTER(Params), C.c_char_p]
        lib.lvt_params_from_file.restype = C.c_int
        lib.lvt_get_frame_info.argtypes = [vp, C.POINTER(FrameInfo)]
        lib.lvt_get_last_pose.argtypes = [vp, c_f64p, c_f64p]
        lib.lvt_debug_get_features.argtypes = [vp, C.c_int, c_f32p, c_u8p, C.c_int]
        lib.lvt_debug_get_points.argtypes = [vp, C.c_int, c_f64p, c_u8p, c_i32p, c_i32p, c_i32p, C.c_int]
        lib.lvt_set_brief_pairs.argtypes = [C.c_void_p]
        lib.lvtk_ctx_create.restype = vp
        lib.lvtk_ctx_create.argtypes = [C.POINTER(Params), C.c_int]
        lib.lvtk_ctx_destroy.argtypes = [vp]
        lib.lvtk_is_gpu.restype = C.c_int
        lib.lvtk_agast.argtypes = [vp, c_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, c_i32p]
        lib.lvtk_detect.argtypes = [vp, c_u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, c_i32p]
        lib.lvtk_brief.argtypes = [vp, c_u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, c_u8p, c_i32p]
        lib.lvtk_extract.argtypes = [vp, c_u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, c_u8p, C.c_int, c_i32p]
        lib.lvtk_match_projected.argtypes = [vp, c_f64p, c_u8p, C.c_int, c_f64p, c_f64p, C.c_void_p, c_u8p, C.c_int,
                                             c_u8p, C.c_int, c_i32p, c_f32p, c_f32p, c_i32p, c_i32p]
        lib.lvtk_row_match.argtypes = [vp, C.c_void_p, c_u8p, C.c_int, c_u8p, C.c_void_p, c_u8p, C.c_int, c_u8p,
                                       c_i32p, c_i32p, c_i32p]
        lib.lvtk_solve_pose.argtypes = [vp, c_f64p, c_f32p, C.c_int, c_f64p, c_f64p, c_f64p, c_f64p, c_u8p]
        lib.lvtk_triangulate.argtypes = [vp, c_f64p, c_f64p, c_f32p, c_f32p, C.c_int, c_f64p, c_u8p]
        lib.lvt_set_rectification.argtypes = [vp, C.POINTER(Rectify), C.POINTER(Rectify)]
        lib.lvt_set_rectification.restype = C.c_int
        lib.lvtk_rectify_maps.argtypes = [vp, C.POINTER(Rectify), C.c_int, C.c_int, c_f32p, c_f32p]
        lib.lvtk_rectify.argtypes = [vp, c_u8p, C.c_int, C.c_int, C.c_int, C.POINTER(Rectify), c_u8p]
        lib.lvt_pool_reserve.argtypes = [vp, C.c_int]
        lib.lvt_pool_upload.argtypes = [vp, C.c_int, c_u8p, c_u8p]
        lib.lvt_track_pool.argtypes = [vp, C.c_int, C.c_int, c_f64p, C.POINTER(FrameInfo)]
        lib.lvt_pool_upload_rgbd.argtypes = [vp, C.c_int, c_u8p, c_f32p]
        lib.lvt_track_batch.argtypes = [vp, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int, c_f64p,
                                        C.POINTER(FrameInfo)]
        lib.lvt_track_batch_rgbd.argtypes = [vp, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                             c_f64p, C.POINTER(FrameInfo)]
        lib.lvt_alloc_pinned.restype = C.c_void_p
        lib.lvt_alloc_pinned.argtypes = [C.c_size_t]
        lib.lvt_free_pinned.argtypes = [C.c_void_p]
        lib.lvt_set_profiling.argtypes = [C.c_int]
        lib.lvt_get_kernel_times.argtypes = [c_f64p, C.POINTER(C.c_long), C.c_int]
        lib.lvt_kernel_name.argtypes = [C.c_int]
        lib.lvt_kernel_name.restype = C.c_char_p
        lib.lvtk_last_error.restype = C.c_char_p
        lib.lvt_last_batch_ms.argtypes = [vp]
        lib.lvt_last_batch_ms.restype = C.c_double
        lib.lvt_launch_count.restype = C.c_long
        self.is_gpu = bool(lib.lvtk_is_gpu())

    def has(self, name):
        return hasattr(self.lib, name)

    def set_profiling(self, on):
        self.lib.lvt_set_profiling(int(on))

    def reset_kernel_times(self):
        self.lib.lvt_reset_kernel_times()

    def kernel_times(self):
        """{kernel name: (total ms, launches)} measured with CUDA events on the launching stream"""
        ms = np.zeros(32)
        cnt = (C.c_long * 32)()
        n = self.lib.lvt_get_kernel_times(_ptr(ms, c_f64p), cnt, 32)
        return {self.lib.lvt_kernel_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    def pinned_empty(self, shape, dtype=np.uint8):
        """numpy array over page-locked host memory from lvt_alloc_pinned (kept alive by the array's base object)"""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = self.lib.lvt_alloc_pinned(n)
        if not ptr:
            raise LvtError("lvt_alloc_pinned(%d) failed" % n)
        lib = self.lib

        class _Owner:
            def __del__(self_inner):
                lib.lvt_free_pinned(ptr)
        buf = (C.c_uint8 * n).from_address(ptr)
        owner = _Owner()
        buf._owner = owner
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def last_error(self):
        return self.lib.lvtk_last_error().decode()

    def launch_count(self):
        return int(self.lib.lvt_launch_count())

    # -- parameters -----------------------------------------------------------------------
    def default_params(self, **overrides):
        p = Params()
        self.lib.lvt_params_default(C.byref(p))
        for k, v in overrides.items():
            setattr(p, k, v)
        return p

    def params_from_file(self, path):
        p = Params()
        if not self.lib.lvt_params_from_file(C.byref(p), os.fsencode(path)):
            raise LvtError("cannot read config %s" % path)
        return p

    def set_brief_pairs(self, pairs=None):
        if pairs is None:
            return self.lib.lvt_set_brief_pairs(None)
        a = np.ascontiguousarray(pairs, dtype=np.int8).reshape(256, 4)
        return self.lib.lvt_set_brief_pairs(a.ctypes.data)

    # -- systems --------------------------------------------------------------------------
    def create(self, params, sensor=SENSOR_STEREO):
        h = self.lib.lvt_create_from_params(C.byref(params), sensor)
        if not h:
            raise LvtError("lvt_create_from_params failed (%s)" % self.path)
        return System(self, h, params)

    def create_from_file(self, path, sensor=SENSOR_STEREO):
        h = self.lib.lvt_create(os.fsencode(path), sensor)
        if not h:
            return None
        return System(self, h, self.params_from_file(path))

    def context(self, params, device=-1):
        h = self.lib.lvtk_ctx_create(C.byref(params), device)
        if not h:
            raise LvtError("lvtk_ctx_create failed (%s)" % self.path)
        return Context(self, h, params)


def _check(rc, what):
    if rc != 0:
        raise LvtError("%s failed with status %d" % (what, rc))


class System:
    """Mirror of lvt_system (lvt/src/lvt_system.h:57-70) over the C ABI."""

    def __init__(self, library, handle, params):
        self.library, self.lib, self.h, self.params = library, library.lib, handle, params
        self._R = np.zeros((3, 3), np.float64)
        self._t = np.zeros(3, np.float64)
        # raise when a (void) tracking call failed instead of handing back the previous pose;
        # False reproduces the reference's behaviour (outputs untouched, nothing reported)
        self.check_status = True

    def destroy(self):
        if self.h:
            self.lib.lvt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def reset(self):
        self.lib.lvt_reset(self.h)

    def get_state(self):
        return self.lib.lvt_get_status(self.h)

    def last_status(self):
        """status of the last tracking call (the reference's calls are void): 0 ok, < 0 LVTK_ERR_*"""
        return int(self.lib.lvt_get_last_status(self.h))

    def _pose_out(self):
        if self.check_status:
            rc = self.last_status()
            if rc != 0:
                raise LvtError("tracking call failed with status %d (%s)" % (rc, self.library.last_error()))
        return self._R.copy(), self._t.copy()

    def track(self, left, right):
        """Stereo: two HxW uint8 images.  Returns (R 3x3, t 3) camera->world."""
        left = np.ascontiguousarray(left, np.uint8)
        right = np.ascontiguousarray(right, np.uint8)
        assert left.shape == right.shape and left.ndim == 2
        self.lib.lvt_track(self.h, _u8(left), _u8(right), left.shape[0], left.shape[1],
                           _ptr(self._R, c_f64p), _ptr(self._t, c_f64p))
        return self._pose_out()

    def track_rgbd(self, gray, depth):
        gray = np.ascontiguousarray(gray, np.uint8)
        depth = np.ascontiguousarray(depth, np.float32)
        assert gray.shape == depth.shape and gray.ndim == 2
        self.lib.lvt_track_rgbd(self.h, _u8(gray), _ptr(depth, c_f32p), gray.shape[0], gray.shape[1],
                                _ptr(self._R, c_f64p), _ptr(self._t, c_f64p))
        return self._pose_out()

    def track_with_external_corners(self, left, right, corners_left, corners_right):
        left = np.ascontiguousarray(left, np.uint8)
        right = np.ascontiguousarray(right, np.uint8)
        cl = np.ascontiguousarray(corners_left, np.float64).reshape(-1, 2)
        cr = np.ascontiguousarray(corners_right, np.float64).reshape(-1, 2)
        self.lib.lvt_track_with_external_corners(self.h, _u8(left), _u8(right), left.shape[0], left.shape[1],
                                                 _ptr(cl, c_f64p), len(cl), _ptr(cr, c_f64p), len(cr),
                                                 _ptr(self._R, c_f64p), _ptr(self._t, c_f64p))
        return self._pose_out()

    def set_rectification(self, left=None, right=None):
        """left / right: Rectify (raw images are rectified on the way in) or None, None (off)."""
        _check(self.lib.lvt_set_rectification(self.h, C.byref(left) if left is not None else None,
                                              C.byref(right) if right is not None else None), "lvt_set_rectification")

    def pool_reserve(self, n_frames):
        _check(self.lib.lvt_pool_reserve(self.h, n_frames), "lvt_pool_reserve")

    def pool_upload(self, frame, left, right):
        """stereo: (left, right) u8 images; RGB-D: (gray u8, depth float32 metres)"""
        left = np.ascontiguousarray(left, np.uint8)
        if np.asarray(right).dtype == np.float32:
            depth = np.ascontiguousarray(right, np.float32)
            _check(self.lib.lvt_pool_upload_rgbd(self.h, frame, _u8(left), _ptr(depth, c_f32p)), "lvt_pool_upload_rgbd")
            return
        right = np.ascontiguousarray(right, np.uint8)
        _check(self.lib.lvt_pool_upload(self.h, frame, _u8(left), _u8(right)), "lvt_pool_upload")

    def track_batch(self, a, b, want_infos=True):
        """n frames in host memory through lvt_track_batch / lvt_track_batch_rgbd.  a: n left (gray) images u8,
        b: n right images u8 or n depth images float32 (lists of arrays, or arrays [n][H][W]).
        Returns (poses n x 12, infos list or None)."""
        n = len(a)
        rgbd = np.asarray(b[0]).dtype == np.float32
        a = [np.ascontiguousarray(x, np.uint8) for x in a]
        b = [np.ascontiguousarray(x, np.float32 if rgbd else np.uint8) for x in b]
        pa = (C.c_void_p * n)(*[x.ctypes.data for x in a])
        pb = (C.c_void_p * n)(*[x.ctypes.data for x in b])
        poses = np.zeros((n, 12), np.float64)
        infos = (FrameInfo * n)() if want_infos else None
        fn = self.lib.lvt_track_batch_rgbd if rgbd else self.lib.lvt_track_batch
        _check(fn(self.h, n, pa, pb, a[0].shape[0], a[0].shape[1], _ptr(poses, c_f64p), infos),
               "lvt_track_batch_rgbd" if rgbd else "lvt_track_batch")
        return poses, ([infos[i].as_dict() for i in range(n)] if want_infos else None)

    def track_pool(self, first, n, want_infos=True):
        """Track resident frames [first, first+n): returns (poses n x 12, infos list or None)."""
        poses = np.zeros((n, 12), np.float64)
        infos = (FrameInfo * n)() if want_infos else None
        _check(self.lib.lvt_track_pool(self.h, first, n, _ptr(poses, c_f64p), infos), "lvt_track_pool")
        return poses, ([infos[i].as_dict() for i in range(n)] if want_infos else None)

    def frame_counters(self, i=-1):
        """profiling aid: fixed-point rounds (map, retry, staged, row passes) and pose-solver evaluations of pool
        frame i of the last batch (i < 0: of the last blocking call)"""
        cyc = (C.c_longlong * 8)()
        rnd = (C.c_int * 8)()
        self.lib.lvt_debug_phase_cycles(C.c_void_p(self.h), i, cyc, rnd)
        return {"rounds": list(rnd)[:4], "lm_evaluations": rnd[4]}

    def frame_marks(self, i=-1):
        """profiling aid (CUDA library): nanosecond marks of pool frame i of the last batch (i < 0: of the last blocking
        call): [0..7] inside track_b (start, staged rounds, promotion, staged compaction, row matching, triangulation,
        append, state), [8], [9] start / end of the frame's early map pass in the batched engine; 0 = phase not run"""
        m = (C.c_longlong * 12)()
        self.lib.lvt_debug_frame_marks(C.c_void_p(self.h), i, m)
        return list(m)

    def last_batch_ms(self):
        return float(self.lib.lvt_last_batch_ms(self.h))

    def point_capacity(self):
        return int(self.lib.lvt_debug_point_capacity(self.h))

    def frame_info(self):
        fi = FrameInfo()
        _check(self.lib.lvt_get_frame_info(self.h, C.byref(fi)), "lvt_get_frame_info")
        return fi.as_dict()

    def last_pose(self):
        q = np.zeros(4)
        t = np.zeros(3)
        _check(self.lib.lvt_get_last_pose(self.h, _ptr(q, c_f64p), _ptr(t, c_f64p)), "lvt_get_last_pose")
        return q, t

    def features(self, which=0):
        n = self.lib.lvt_debug_get_features(self.h, which, None, None, 0)
        if n < 0:
            raise LvtError("lvt_debug_get_features failed")
        xy = np.zeros((n, 2), np.float32)
        desc = np.zeros((n, 32), np.uint8)
        if n:
            self.lib.lvt_debug_get_features(self.h, which, _ptr(xy, c_f32p), _u8(desc), n)
        return xy, desc

    def points(self, which=0):
        n = self.lib.lvt_debug_get_points(self.h, which, None, None, None, None, None, 0)
        if n < 0:
            raise LvtError("lvt_debug_get_points failed")
        xyz = np.zeros((n, 3), np.float64)
        desc = np.zeros((n, 32), np.uint8)
        cnt = np.zeros(n, np.int32)
        age = np.zeros(n, np.int32)
        mi = np.zeros(n, np.int32)
        if n:
            self.lib.lvt_debug_get_points(self.h, which, _ptr(xyz, c_f64p), _u8(desc), _ptr(cnt, c_i32p),
                                          _ptr(age, c_i32p), _ptr(mi, c_i32p), n)
        return {"xyz": xyz, "desc": desc, "counter": cnt, "age": age, "match_idx": mi}


class Context:
    """The seam ABI of include/lvt_kernels.h, one method per entry point."""

    def __init__(self, library, handle, params):
        self.library, self.lib, self.h, self.params = library, library.lib, handle, params

    def destroy(self):
        if self.h:
            self.lib.lvtk_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    @staticmethod
    def _img(img):
        img = np.asarray(img)
        assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
        return img, img.shape[0], img.shape[1], img.strides[0]

    def rectify_maps(self, rect, rows, cols):
        mx = np.zeros((rows, cols), np.float32)
        my = np.zeros((rows, cols), np.float32)
        _check(self.lib.lvtk_rectify_maps(self.h, C.byref(rect), rows, cols, _ptr(mx, c_f32p), _ptr(my, c_f32p)),
               "lvtk_rectify_maps")
        return mx, my

    def rectify(self, img, rect):
        img, rows, cols, stride = self._img(img)
        out = np.zeros((rows, cols), np.uint8)
        _check(self.lib.lvtk_rectify(self.h, _u8(img), rows, cols, stride, C.byref(rect), _u8(out)), "lvtk_rectify")
        return out

    def agast(self, img, threshold, nonmax=True, cap=None):
        img, rows, cols, stride = self._img(img)
        cap = cap or rows * cols
        out = np.zeros(cap, KP_DTYPE)
        n = C.c_int(0)
        _check(self.lib.lvtk_agast(self.h, _u8(img), rows, cols, stride, threshold, int(nonmax), out.ctypes.data,
                                   cap, C.byref(n)), "lvtk_agast")
        return out[:n.value].copy()

    def detect(self, img, cap=None):
        img, rows, cols, stride = self._img(img)
        cap = cap or rows * cols // 2
        out = np.zeros(cap, KP_DTYPE)
        n = C.c_int(0)
        _check(self.lib.lvtk_detect(self.h, _u8(img), rows, cols, stride, out.ctypes.data, cap, C.byref(n)),
               "lvtk_detect")
        return out[:n.value].copy()

    def brief(self, img, kps):
        img, rows, cols, stride = self._img(img)
        kps = np.ascontiguousarray(kps, KP_DTYPE)
        out = np.zeros(max(len(kps), 1), KP_DTYPE)
        desc = np.zeros((max(len(kps), 1), 32), np.uint8)
        n = C.c_int(0)
        _check(self.lib.lvtk_brief(self.h, _u8(img), rows, cols, stride, kps.ctypes.data, len(kps), out.ctypes.data,
                                   _u8(desc), C.byref(n)), "lvtk_brief")
        return out[:n.value].copy(), desc[:n.value].copy()

    def extract(self, img, cap=16384):
        img, rows, cols, stride = self._img(img)
        out = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        _check(self.lib.lvtk_extract(self.h, _u8(img), rows, cols, stride, out.ctypes.data, _u8(desc), cap,
                                     C.byref(n)), "lvtk_extract")
        return out[:n.value].copy(), desc[:n.value].copy()

    def match_projected(self, pts_xyz, pts_desc, q_wxyz, t, kps, desc, matched=None, retry_below=50):
        pts_xyz = np.ascontiguousarray(pts_xyz, np.float64).reshape(-1, 3)
        pts_desc = np.ascontiguousarray(pts_desc, np.uint8).reshape(-1, 32)
        kps = np.ascontiguousarray(kps, KP_DTYPE)
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        m, n = len(pts_xyz), len(kps)
        flags = np.zeros(max(n, 1), np.uint8) if matched is None else np.ascontiguousarray(matched, np.uint8).copy()
        q = np.ascontiguousarray(q_wxyz, np.float64)
        t = np.ascontiguousarray(t, np.float64)
        idx = np.zeros(max(m, 1), np.int32)
        d1 = np.zeros(max(m, 1), np.float32)
        d2 = np.zeros(max(m, 1), np.float32)
        cnt = C.c_int(0)
        retried = C.c_int(0)
        _check(self.lib.lvtk_match_projected(self.h, _ptr(pts_xyz, c_f64p), _u8(pts_desc), m, _ptr(q, c_f64p),
                                             _ptr(t, c_f64p), kps.ctypes.data, _u8(desc), n, _u8(flags), retry_below,
                                             _ptr(idx, c_i32p), _ptr(d1, c_f32p), _ptr(d2, c_f32p), C.byref(cnt),
                                             C.byref(retried)), "lvtk_match_projected")
        return {"idx": idx[:m], "d1": d1[:m], "d2": d2[:m], "count": cnt.value, "retried": retried.value,
                "matched": flags[:n]}

    def row_match(self, kps_l, desc_l, kps_r, desc_r, matched_l=None, matched_r=None):
        kps_l = np.ascontiguousarray(kps_l, KP_DTYPE)
        kps_r = np.ascontiguousarray(kps_r, KP_DTYPE)
        desc_l = np.ascontiguousarray(desc_l, np.uint8).reshape(-1, 32)
        desc_r = np.ascontiguousarray(desc_r, np.uint8).reshape(-1, 32)
        nl, nr = len(kps_l), len(kps_r)
        ml = np.zeros(max(nl, 1), np.uint8) if matched_l is None else np.ascontiguousarray(matched_l, np.uint8).copy()
        mr = np.zeros(max(nr, 1), np.uint8) if matched_r is None else np.ascontiguousarray(matched_r, np.uint8).copy()
        q = np.zeros(max(nl, 1), np.int32)
        t = np.zeros(max(nl, 1), np.int32)
        n = C.c_int(0)
        _check(self.lib.lvtk_row_match(self.h, kps_l.ctypes.data, _u8(desc_l), nl, _u8(ml), kps_r.ctypes.data,
                                       _u8(desc_r), nr, _u8(mr), _ptr(q, c_i32p), _ptr(t, c_i32p), C.byref(n)),
               "lvtk_row_match")
        return {"query": q[:n.value].copy(), "train": t[:n.value].copy(), "matched_left": ml[:nl],
                "matched_right": mr[:nr]}

    def solve_pose(self, pts_xyz, uv, q_wxyz, t):
        pts_xyz = np.ascontiguousarray(pts_xyz, np.float64).reshape(-1, 3)
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        m = len(pts_xyz)
        q = np.ascontiguousarray(q_wxyz, np.float64)
        t = np.ascontiguousarray(t, np.float64)
        qo = np.zeros(4)
        to = np.zeros(3)
        marks = np.zeros(max(m, 1), np.uint8)
        _check(self.lib.lvtk_solve_pose(self.h, _ptr(pts_xyz, c_f64p), _ptr(uv, c_f32p), m, _ptr(q, c_f64p),
                                        _ptr(t, c_f64p), _ptr(qo, c_f64p), _ptr(to, c_f64p), _u8(marks)),
               "lvtk_solve_pose")
        return qo, to, marks[:m]

    def triangulate(self, q_wxyz, t, uv_left, uv_right):
        uv_left = np.ascontiguousarray(uv_left, np.float32).reshape(-1, 2)
        uv_right = np.ascontiguousarray(uv_right, np.float32).reshape(-1, 2)
        n = len(uv_left)
        q = np.ascontiguousarray(q_wxyz, np.float64)
        t = np.ascontiguousarray(t, np.float64)
        xyz = np.zeros((max(n, 1), 3), np.float64)
        valid = np.zeros(max(n, 1), np.uint8)
        _check(self.lib.lvtk_triangulate(self.h, _ptr(q, c_f64p), _ptr(t, c_f64p), _ptr(uv_left, c_f32p),
                                         _ptr(uv_right, c_f32p), n, _ptr(xyz, c_f64p), _u8(valid)),
               "lvtk_triangulate")
        return xyz[:n], valid[:n]
