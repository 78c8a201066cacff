"""Per-kernel device times (CUDA events) of the blocking path for any configuration."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lvt_b200
from lvt_b200 import configs, synth
name = sys.argv[1] if len(sys.argv) > 1 else "tum_synth"
n = 40
cfg = configs.CONFIGS[name]
p = configs.make_params(name)
st = (synth.StereoStream if cfg["sensor"] == 1 else synth.RgbdStream)(n_frames=n, seed=0, **cfg["stream"])
frames = [tuple(np.ascontiguousarray(a) for a in st.frame(t)) for t in range(n)]
lib = lvt_b200.load()
vo = lib.create(p, cfg["sensor"])
run = (lambda a, b: vo.track(a, b)) if cfg["sensor"] == 1 else (lambda a, b: vo.track_rgbd(a, b))
for t in range(10):
    run(*frames[t])
lib.reset_kernel_times(); lib.set_profiling(True)
for t in range(10, 25):
    run(*frames[t])
lib.set_profiling(False)
tot = 0
for k, (ms, cnt) in lib.kernel_times().items():
    if cnt:
        print("  %-22s %8.1f us x %d" % (k, 1e3 * ms / cnt, cnt)); tot += ms
print("  sum per frame %.1f us" % (1e3 * tot / 15))
t0 = time.perf_counter()
for t in range(25, n):
    run(*frames[t])
print("blocking: %.1f us/frame" % (1e6 * (time.perf_counter() - t0) / (n - 25)), vo.frame_info())
import ctypes as C
cyc = (C.c_longlong * 8)(); rnd = (C.c_int * 8)()
lib.lib.lvt_debug_phase_cycles(C.c_void_p(vo.h), -1, cyc, rnd)
c = list(cyc)
print("rounds (map, retry, staged, row):", list(rnd), "| track_a: match %.1f us, bookkeeping %.1f us | track_b: clean %.1f us, rest %.1f us" % (
    (c[1] - c[0]) / 1e3, (c[2] - c[1]) / 1e3, (c[6] - c[5]) / 1e3, (c[7] - c[6]) / 1e3))
