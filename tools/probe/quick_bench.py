"""value / blocking / batch frames/s of config 2 in a few seconds (development aid; bench.py is the record)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lvt_b200
from lvt_b200 import configs, synth
name = sys.argv[1] if len(sys.argv) > 1 else "kitti_synth"
cfg = configs.CONFIGS[name]
sensor = cfg["sensor"]
n = 300 if sensor == 1 else 120
p = configs.make_params(name)
st = (synth.StereoStream if sensor == 1 else synth.RgbdStream)(n_frames=n, seed=0, **cfg["stream"])
fr = [st.frame(t) for t in range(n)]
lib = lvt_b200.load()
H, W = fr[0][0].shape
pa = lib.pinned_empty((n, H, W))
pb = lib.pinned_empty((n, H, W), np.uint8 if sensor == 1 else np.float32)
for i in range(n):
    pa[i], pb[i] = fr[i]
vo = lib.create(p, sensor)
vo.pool_reserve(n)
for i in range(n):
    vo.pool_upload(i, pa[i], pb[i])
w = n // 3
import ctypes as C
ht = (C.c_double * 4)()
vo.track_pool(0, w, want_infos=False)
lib.lib.lvt_debug_host_times(C.c_void_p(vo.h), ht, 1)
l0 = lib.launch_count()
vo.track_pool(w, n - w, want_infos=False)
lib.lib.lvt_debug_host_times(C.c_void_p(vo.h), ht, 0)
print("engine: host enqueue %.1f us/frame, then waiting %.1f us/frame; %.2f kernel launches/frame" % (ht[1] / (n - w), ht[2] / (n - w), (lib.launch_count() - l0) / (n - w)))
print("value    %7.0f frames/s (%.1f us/frame)" % ((n - w) / (vo.last_batch_ms() * 1e-3), 1e3 * vo.last_batch_ms() / (n - w)))
vo.destroy()
vo = lib.create(p, sensor)
trk = (lambda a, b: vo.track(a, b)) if sensor == 1 else (lambda a, b: vo.track_rgbd(a, b))
for i in range(w):
    trk(pa[i], pb[i])
t0 = time.perf_counter()
for i in range(w, n):
    trk(pa[i], pb[i])
dt = time.perf_counter() - t0
print("blocking %7.0f frames/s (%.1f us/frame, page-locked buffers)" % ((n - w) / dt, 1e6 * dt / (n - w)))
vo.destroy()
vo = lib.create(p, sensor)
vo.track_batch([pa[i] for i in range(w)], [pb[i] for i in range(w)], want_infos=False)
t0 = time.perf_counter()
vo.track_batch([pa[i] for i in range(w, n)], [pb[i] for i in range(w, n)], want_infos=False)
dt = time.perf_counter() - t0
print("batch    %7.0f frames/s (page-locked buffers)" % ((n - w) / dt))
