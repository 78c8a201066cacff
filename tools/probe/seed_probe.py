"""resident-path frames/s of config 2 for the synthetic sequences 0..7 (the 8 ranks of an 8-GPU run track one each):
how much of the 1 -> 8 GPU efficiency is the spread between sequences (the slowest one sets the job's time)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lvt_b200
from lvt_b200 import configs, synth
name = "kitti_synth"
cfg = configs.CONFIGS[name]
p = configs.make_params(name)
lib = lvt_b200.load()
n = 700
out = []
for seed in range(8):
    st = synth.StereoStream(n_frames=2700, seed=seed, **cfg["stream"])
    vo = lib.create(p, 1)
    vo.pool_reserve(n)
    for t in range(n):
        vo.pool_upload(t, *st.frame(t))
    vo.track_pool(0, 300, want_infos=False)
    ms = 0.0
    rounds = []
    for s in range(300, n, 100):
        vo.track_pool(s, 100, want_infos=False)
        ms += vo.last_batch_ms()
        rounds += [vo.frame_counters(i)["rounds"][0] for i in range(100)]
    fps = (n - 300) / (ms * 1e-3)
    out.append(fps)
    print("seed %d: %.0f frames/s (%.1f us/frame) mean map rounds %.2f" % (seed, fps, 1e3 * ms / (n - 300), np.mean(rounds) if rounds else -1), flush=True)
    vo.destroy()
print("min / mean / seed 0: %.0f / %.0f / %.0f -> slowest / seed 0 = %.3f" % (min(out), np.mean(out), out[0], min(out) / out[0]))
