import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lvt_b200
from lvt_b200 import configs, synth
lib = lvt_b200.load()
name = sys.argv[1] if len(sys.argv) > 1 else "kitti_synth"
p = configs.make_params(name)
n = 40
st = synth.StereoStream(n_frames=n, seed=0, **configs.CONFIGS[name]["stream"])
vo = lib.create(p, 1)
vo.pool_reserve(n)
for t in range(n):
    vo.pool_upload(t, *st.frame(t))
vo.track_pool(0, 10)
NN = int(os.environ.get("PROBE_N", "30"))
poses, infos = vo.track_pool(10, NN)
print("batch ms", vo.last_batch_ms(), "per frame", vo.last_batch_ms() / NN)
cyc = (C.c_longlong * 8)()
rnd = (C.c_int * 8)()
names = ["A.match", "A.bookkeep", "(gap)", "pose", "(gap)", "B.clean", "B.staged+tri"]  # phase marks: GPU-wide ns timer
prev_end = None
gaps = []
acc = np.zeros(7)
clean = 0.0
for i in range(NN):
    lib.lib.lvt_debug_phase_cycles(C.c_void_p(vo.h), i, cyc, rnd)
    c = np.array(list(cyc), dtype=np.float64)
    d = np.diff(c[:8])
    if prev_end is not None and c[0] > 0:
        gaps.append((c[0] - prev_end) / 1e3)
    prev_end = c[7]
    acc += d
    clean += list(rnd)[5]
    if i < 12:
        print(i, "rounds", list(rnd), "tri", infos[i]["triangulated"], "staged", infos[i]["staged_before"], " ".join("%s=%.0fus" % (nm, v / 1e3) for nm, v in zip(names, d)))
bm = (C.c_longlong * 12)()
# timeline of the overlapped chain, relative to the previous frame's pose (us): early map pass start / end, the previous
# frame's track_b end, this frame's track_a (rest) start
cy = (C.c_longlong * 8)(); rn = (C.c_int * 8)()
prev = None
print("overlap timeline (us after the previous frame's pose was written): early pass start, end | previous track_b start, end | rest of track_a start")
for i in range(NN):
    lib.lib.lvt_debug_phase_cycles(C.c_void_p(vo.h), i, cy, rn)
    lib.lib.lvt_debug_frame_marks(C.c_void_p(vo.h), i, bm)
    cur = (list(cy), list(bm))
    if prev is not None and i < 14 and bm[8] > 0:
        p4 = prev[0][4]
        print("  %2d  early %.1f .. %.1f | track_b %.1f .. %.1f | rest %.1f" % (i, (bm[8] - p4) / 1e3, (bm[9] - p4) / 1e3,
              (prev[0][5] - p4) / 1e3, (prev[0][7] - p4) / 1e3, (cy[0] - p4) / 1e3))
    prev = cur
bn = ["staged rounds", "promotion", "staged compaction", "row matching", "triangulation", "append", "state+prediction"]
print("track_b phases (us since the previous mark; - = not run):")
for i in range(12):
    lib.lib.lvt_debug_frame_marks(C.c_void_p(vo.h), i, bm)
    m = list(bm)
    prev, parts = m[0], []
    for k in range(1, 8):
        if m[k] > 0:
            parts.append("%s=%.1f" % (bn[k - 1], (m[k] - prev) / 1e3)); prev = m[k]
    print("  ", i, " ".join(parts))
print("map culling next to the pose solver: %.1f us (cycles / 1965)" % (clean / NN / 1965.0))
print("frame end -> next frame's track_a start: mean %.1f us" % (np.mean(gaps) if gaps else 0.0))
print("mean us:", " ".join("%s=%.1f" % (nm, v / NN / 1e3) for nm, v in zip(names, acc)), "total=%.1f" % (acc.sum() / NN / 1e3))
lib.reset_kernel_times(); lib.set_profiling(True)
vo.track_pool(20, 20, want_infos=False)
lib.set_profiling(False)
for k, (ms, cnt) in lib.kernel_times().items():
    if cnt: print("%-22s %8.1f us x %d" % (k, 1e3 * ms / cnt, cnt))
# isolated kernel times: the blocking per-frame path (one stream, nothing overlapped)
vo2 = lib.create(p, 1)
for t in range(6):
    vo2.track(*st.frame(t))
lib.reset_kernel_times(); lib.set_profiling(True)
import time
t0 = time.perf_counter()
for t in range(6, 26):
    vo2.track(*st.frame(t))
dt = time.perf_counter() - t0
lib.set_profiling(False)
print("blocking lvt_track with profiling events: %.1f us/frame" % (1e6 * dt / 20))
tot = 0
for k, (ms, cnt) in lib.kernel_times().items():
    if cnt:
        print("  %-22s %8.1f us x %d" % (k, 1e3 * ms / cnt, cnt)); tot += ms
print("  sum per frame %.1f us" % (1e3 * tot / 20))
t0 = time.perf_counter()
for t in range(26, 40):
    vo2.track(*st.frame(t))
print("blocking lvt_track: %.1f us/frame" % (1e6 * (time.perf_counter() - t0) / 14))
