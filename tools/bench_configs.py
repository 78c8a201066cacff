#!/usr/bin/env python
"""Frames/s of the BASELINE.json configurations that are parity cases rather than bench.py lines
(config 3: 640x480 RGB-D, config 5: 752x480 EuRoC shape at ~5000 keypoints/frame), next to the CPU
oracle on the same frames.  Prints one JSON object per configuration.
    python tools/bench_configs.py [n_frames]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lvt_b200
from lvt_b200 import capi, configs, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
lib = lvt_b200.load()
orc = capi.Library(os.path.join(ROOT, "oracle", "_build", "liblvt_oracle.so"))
for name in ("tum_synth", "euroc_synth"):
    cfg = configs.CONFIGS[name]
    p = configs.make_params(name)
    st = (synth.StereoStream if cfg["sensor"] == 1 else synth.RgbdStream)(n_frames=n, seed=0, **cfg["stream"])
    frames = [tuple(np.ascontiguousarray(a) for a in st.frame(t)) for t in range(n)]
    out = {"config": name, "sensor": cfg["sensor"], "frames": n, "size": [p.img_width, p.img_height]}
    for tag, L, nn in (("b200", lib, n), ("cpu_oracle", orc, min(n, 25))):
        vo = L.create(p, cfg["sensor"])
        run = (lambda a, b: vo.track(a, b)) if cfg["sensor"] == 1 else (lambda a, b: vo.track_rgbd(a, b))
        for t in range(5):
            run(*frames[t])
        t0 = time.perf_counter()
        for t in range(5, nn):
            run(*frames[t])
        dt = time.perf_counter() - t0
        fi = vo.frame_info()
        out[tag + "_blocking_fps"] = round((nn - 5) / dt, 1)
        out[tag + "_state"] = fi["state"]
        out["keypoints"] = fi["n_features_left"]
        out["map_points"] = fi["map_points_after"]
        out["tracked"] = fi["tracked"]
        vo.destroy()
    if cfg["sensor"] == 1:
        vo = lib.create(p, 1)
        vo.pool_reserve(n)
        for t in range(n):
            vo.pool_upload(t, *frames[t])
        vo.track_pool(0, 10, want_infos=False)
        vo.track_pool(10, n - 10, want_infos=False)
        out["b200_resident_fps"] = round((n - 10) / (vo.last_batch_ms() * 1e-3), 1)
        vo.destroy()
    print(json.dumps(out), flush=True)
