"""K handles on K host threads, frames resident in HBM: aggregate frames/s, per-thread host enqueue / wait time."""
import os, sys, time, threading
import ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lvt_b200
from lvt_b200 import configs, synth
K = int(sys.argv[1]) if len(sys.argv) > 1 else 4
name = sys.argv[2] if len(sys.argv) > 2 else "kitti_synth"
n = 200
cfg = configs.CONFIGS[name]
p = configs.make_params(name)
lib = lvt_b200.load()
vos = []
for k in range(K):
    st = synth.StereoStream(n_frames=n, seed=k, **cfg["stream"])
    vo = lib.create(p, 1)
    vo.pool_reserve(n)
    for t in range(n):
        vo.pool_upload(t, *st.frame(t))
    vos.append(vo)
for vo in vos:
    vo.track_pool(0, 100, want_infos=False)
res = [None] * K
bar = threading.Barrier(K + 1)
def work(k):
    vo = vos[k]
    ht = (C.c_double * 4)()
    lib.lib.lvt_debug_host_times(C.c_void_p(vo.h), ht, 1)
    bar.wait()
    t0 = time.perf_counter()
    vo.track_pool(100, 100, want_infos=False)
    dt = time.perf_counter() - t0
    lib.lib.lvt_debug_host_times(C.c_void_p(vo.h), ht, 0)
    res[k] = (dt, ht[1] / 100, ht[2] / 100, vo.last_batch_ms())
th = [threading.Thread(target=work, args=(k,)) for k in range(K)]
for t in th: t.start()
bar.wait()
t0 = time.perf_counter()
for t in th: t.join()
wall = time.perf_counter() - t0
print("K=%d: aggregate %.0f frames/s (wall %.1f ms)" % (K, K * 100 / wall, wall * 1e3))
for k, r in enumerate(res):
    print("  seq %d: call %.1f ms, host enqueue %.1f us/frame, waiting %.1f us/frame, device %.1f ms" % (k, r[0] * 1e3, r[1], r[2], r[3]))
