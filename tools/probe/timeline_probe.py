"""Device-side timeline of the blocking lvt_track call (LVT_B200_TIMELINE=1), page-locked caller buffers."""
import ctypes as C
import os, sys, time
os.environ["LVT_B200_TIMELINE"] = "1"
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lvt_b200
from lvt_b200 import configs, synth
name = sys.argv[1] if len(sys.argv) > 1 else "kitti_synth"
p = configs.make_params(name)
n = 160
st = synth.StereoStream(n_frames=n, seed=0, **configs.CONFIGS[name]["stream"])
lib = lvt_b200.load()
H, W = st.H, st.W
buf = lib.pinned_empty((n, 2, H, W))
for t in range(n):
    buf[t, 0], buf[t, 1] = st.frame(t)
vo = lib.create(p, 1)
R = np.zeros((3, 3)); tt = np.zeros(3)
u8p, f64p = C.POINTER(C.c_uint8), C.POINTER(C.c_double)
args = [(vo.h, buf[t, 0].ctypes.data_as(u8p), buf[t, 1].ctypes.data_as(u8p), H, W, R.ctypes.data_as(f64p), tt.ctypes.data_as(f64p)) for t in range(n)]
trk = lib.lib.lvt_track
for t in range(10):
    trk(*args[t])
ht = (C.c_double * 4)()
lib.lib.lvt_debug_host_times(C.c_void_p(vo.h), ht, 1)
t0 = time.perf_counter()
for t in range(10, n):
    trk(*args[t])
dt = time.perf_counter() - t0
lib.lib.lvt_debug_host_times(C.c_void_p(vo.h), ht, 0)
print("blocking lvt_track: %.1f us/frame; host: upload enqueue %.1f | kernel enqueue %.1f | waiting %.1f us" %
      (1e6 * dt / (n - 10), ht[0] / ht[3], ht[1] / ht[3], ht[2] / ht[3]))
