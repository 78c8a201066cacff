"""Blocking lvt_track: wall time per frame and where the host spends it, for a few staging settings.
Optionally (LVT_STATS_LIB=path of a -DLVT_NMS_STATS build) the NMS kernel's internal statistics."""
import os, sys, time, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lvt_b200
from lvt_b200 import configs, synth, capi

name = sys.argv[1] if len(sys.argv) > 1 else "kitti_synth"
p = configs.make_params(name)
n = 60
st = synth.StereoStream(n_frames=n, seed=0, **configs.CONFIGS[name]["stream"])
frames = [tuple(np.ascontiguousarray(a) for a in st.frame(t)) for t in range(n)]

stats_lib = os.environ.get("LVT_STATS_LIB")
if stats_lib:
    lib = capi.Library(stats_lib)
    vo = lib.create(p, 1)
    for t in range(12):
        vo.track(*frames[t])
    tr = np.zeros((4096, 12), dtype=np.int64)
    lib.lib.lvt_debug_nms_trace(tr.ctypes.data_as(C.c_void_p))
    nx, ny = (p.img_width + 47) // 48, (p.img_height + 47) // 48  # kNmsTile
    tr = tr[: nx * ny * 2]
    mhz = 1965.0
    t0 = tr[:, 0].min()
    print("NMS trace of the last full launch: %d CTAs, kernel span %.1f us" % (len(tr), (tr[:, 1].max() - t0) / 1e3))
    print("  CTA start (us): p50 %.1f p90 %.1f max %.1f | CTA end: p50 %.1f p90 %.1f max %.1f" % (
        np.percentile(tr[:, 0] - t0, 50) / 1e3, np.percentile(tr[:, 0] - t0, 90) / 1e3, (tr[:, 0] - t0).max() / 1e3,
        np.percentile(tr[:, 1] - t0, 50) / 1e3, np.percentile(tr[:, 1] - t0, 90) / 1e3, (tr[:, 1] - t0).max() / 1e3))
    for k, nm in [(2, "stage"), (3, "propagate"), (4, "classify"), (5, "resolve"), (6, "flush"), (11, "longest single resolve")]:
        v = tr[:, k] / mhz
        print("  %-24s mean %.1f p90 %.1f max %.1f us" % (nm, v.mean(), np.percentile(v, 90), v.max()))
    for k, nm in [(7, "slow candidates"), (8, "corners in window"), (9, "survivors"), (10, "sweeps")]:
        v = tr[:, k]
        print("  %-24s mean %.1f p90 %.0f max %d" % (nm, v.mean(), np.percentile(v, 90), v.max()))
    worst = np.argsort(tr[:, 1])[-5:]
    for w in worst:
        print("  late CTA %d: start %.1f end %.1f" % (w, (tr[w, 0] - t0) / 1e3, (tr[w, 1] - t0) / 1e3), [round(x / mhz, 1) for x in tr[w, 2:7]], list(tr[w, 7:11]), round(tr[w, 11] / mhz, 1))
    tt = np.zeros((2048, 12), dtype=np.int64)
    lib.lib.lvt_debug_tile_trace(tt.ctypes.data_as(C.c_void_p))
    nt = ((p.img_width + p.detection_cell_size - 1) // p.detection_cell_size) * ((p.img_height + p.detection_cell_size - 1) // p.detection_cell_size)
    print("tile trace (last launch, %d tiles): span %.1f us" % (nt, (tt[:nt, 1].max() - tt[:nt, 0].min()) / 1e3))
    for i in range(nt):
        print("  tile %2d n=%4d fallback=%d raster %.1f | sort %.1f | hist %.1f | radii %.1f | select %.1f | compact %.1f us" % (
            (i, tt[i, 10], tt[i, 11]) + tuple(tt[i, k] / mhz for k in (2, 3, 4, 5, 6, 7))))
    sys.exit(0)

lib = lvt_b200.load()
lib.lib.lvt_debug_host_times.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
for threads, bands in [(0, 0), (1, 1), (2, 2), (3, 3), (4, 4), (4, 8), (6, 6)]:
    if threads:
        os.environ["LVT_B200_UPLOAD_THREADS"] = str(threads)
        os.environ["LVT_B200_UPLOAD_BANDS"] = str(bands)
    vo = lib.create(p, 1)
    for t in range(10):
        vo.track(*frames[t])
    out = (C.c_double * 4)()
    lib.lib.lvt_debug_host_times(C.c_void_p(vo.h), out, 1)
    t0 = time.perf_counter()
    for t in range(10, n):
        vo.track(*frames[t])
    dt = time.perf_counter() - t0
    lib.lib.lvt_debug_host_times(C.c_void_p(vo.h), out, 1)
    k = out[3]
    print("threads %d bands %d: %.1f us/frame | stage %.1f | enqueue %.1f | wait %.1f" % (threads, bands, 1e6 * dt / (n - 10), out[0] / k, out[1] / k, out[2] / k))
    del vo
