#!/usr/bin/env python
"""Benchmark of the per-frame track() path (BASELINE.json metric: stereo frames/s at 1242x375).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A "step" is one pass of the hot path over FRAMES_PER_STEP consecutive frames of a synthetic stereo
stream (1242x375, ~2000 keypoints/frame; the local map is the
emergent ~2000 points).  N > 1: N independent sequences, one per GPU (weak scaling), the only
collective is one NCCL broadcast of the calibration block.

`value`  : whole-job frames/s with the frames already resident in HBM (lvt_track_pool), timed on the
           device with CUDA events, max over ranks.
`e2e`    : the same metric through the reference's own C call, lvt_track(handle, left*, right*, ...),
           with HOST buffers -- H2D of both images and D2H of the pose inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from lvt_b200 import capi, configs, synth  # noqa: E402

CONFIG = "kitti_synth"
FRAMES_PER_STEP = 50
REF_FRAMES_PER_STEP = 10
CPU_SAMPLE_FRAMES = 400
METRIC = "stereo frames/sec at 1242x375"
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liblvt_oracle.so")


def load_traffic():
    """DRAM bytes (read + write) per launch from the committed `ncu --set full` capture (profiles/traffic.json,
    written by tools/summarize_profiles.py); {} when absent"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)["dram_bytes_per_launch"]
    except Exception:
        return {}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        self.paused = False
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def pause(self):
        """end of one timed region: stop sampling, keep the lines"""
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.proc.wait()
            self.paused = True

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if not getattr(self, "paused", False):
            time.sleep(0.15)
            self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_frames(stream, first, n):
    """n stereo pairs as one contiguous host array [n][2][H][W] (what a caller of lvt_track holds)."""
    out = np.empty((n, 2, stream.H, stream.W), np.uint8)
    for i in range(n):
        out[i, 0], out[i, 1] = stream.frame(first + i)
    return out


def cpu_baseline(params, stream, n_frames):
    """the oracle (CPU port of the reference path) timed on this box's host cores: wall clock around
    track() only, as examples/kitti/kitti_example.cpp:129-131 does"""
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    orc = capi.Library(ORACLE_SO)
    vo = orc.create(params, 1)
    frames = host_frames(stream, 0, n_frames)
    for i in range(min(10, n_frames)):  # warm-up, discarded
        vo.track(frames[i, 0], frames[i, 1])
    t0 = time.perf_counter()
    for i in range(10, n_frames):
        vo.track(frames[i, 0], frames[i, 1])
    dt = time.perf_counter() - t0
    ok = vo.get_state() == capi.STATE_TRACKING
    vo.destroy()
    return {"value": (n_frames - 10) / dt, "unit": "frames/s", "cores": 2, "kind": "port",
            "sample": "%d consecutive frames of the same stream after 10 warm-up frames, 1 sequence, 2 threads "
                      "(left/right extraction, as lvt_image_features_handler.cpp:204-206), tracking=%s" % (n_frames - 10, ok)}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the genuine sources cannot be built
    here) on all the host threads it can use: N sequences in parallel, 2 threads each."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.gpus
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    orc = capi.Library(ORACLE_SO)
    params = configs.make_params(CONFIG)
    fps_step = REF_FRAMES_PER_STEP
    total = (args.warmup + args.steps) * fps_step
    streams = [synth.StereoStream(n_frames=total, seed=s, **configs.CONFIGS[CONFIG]["stream"]) for s in range(n)]
    frames = [host_frames(st, 0, total) for st in streams]
    vos = [orc.create(params, 1) for _ in range(n)]
    cores = os.cpu_count() or 1
    workers = max(1, min(n, cores // 2))

    def run_seq(s, lo, hi):
        for i in range(lo, hi):
            vos[s].track(frames[s][i, 0], frames[s][i, 1])

    def run_all(lo, hi):
        pending = list(range(n))
        lock = threading.Lock()

        def worker():
            while True:
                with lock:
                    if not pending:
                        return
                    s = pending.pop()
                run_seq(s, lo, hi)  # ctypes releases the GIL inside lvt_track
        th = [threading.Thread(target=worker) for _ in range(workers)]
        [t.start() for t in th]
        [t.join() for t in th]

    run_all(0, args.warmup * fps_step)
    t0 = time.perf_counter()
    run_all(args.warmup * fps_step, total)
    dt = time.perf_counter() - t0
    value = n * args.steps * fps_step / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": n, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "1242x375 synthetic stereo stream (SURVEY 8d config 2), %d independent sequence(s); "
                                   "reference arm: %d frames per step per sequence (bounded sample)" % (n, fps_step),
                       "frames_per_step": fps_step, "sequences": n},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": min(cores, 2 * workers), "kind": "port",
                             "sample": "%d sequences x %d frames, %d sequences at a time, 2 threads each, %d host cores"
                                       % (n, args.steps * fps_step, workers, cores)},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist
    import lvt_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = lvt_b200.load()

    # the only collective: broadcast of the calibration / parameter block from rank 0 over NCCL
    p0 = configs.make_params(CONFIG)
    blob = torch.from_numpy(p0.to_array() if rank == 0 else np.zeros_like(p0.to_array())).cuda()
    if world > 1:
        dist.broadcast(blob, src=0)
    params = capi.Params.from_array(blob.cpu().numpy())

    fps_step = args.frames_per_step
    n_value = (args.warmup + 2 * args.steps) * fps_step  # warm-up, timed, profiled repeat
    n_e2e = (args.warmup + args.steps) * fps_step
    stream = synth.StereoStream(n_frames=max(n_value, n_e2e, CPU_SAMPLE_FRAMES), seed=rank, **configs.CONFIGS[CONFIG]["stream"])
    H, W = stream.H, stream.W

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- value: frames resident in HBM, pipelined, device-timed ----------------------
    vo = lib.create(params, 1)
    vo.pool_reserve(n_value)
    chunk = host_frames(stream, 0, n_value)
    for i in range(n_value):
        vo.pool_upload(i, chunk[i, 0], chunk[i, 1])
    for s in range(args.warmup):
        vo.track_pool(s * fps_step, fps_step, want_infos=False)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = lib.launch_count()
    dev_ms, infos = 0.0, []
    t0 = time.perf_counter()
    for s in range(args.warmup, args.warmup + args.steps):
        _, inf = vo.track_pool(s * fps_step, fps_step)
        dev_ms += vo.last_batch_ms()
        infos += inf
    torch.cuda.synchronize()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    launches = lib.launch_count() - launches0
    barrier()
    sampler.pause()
    dev_ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks(wall_ms)
    n_timed = args.steps * fps_step
    value = world * n_timed / (dev_ms * 1e-3)
    lost = sum(1 for i in infos if i["state"] != capi.STATE_TRACKING)

    # ---------------- roofline: per-kernel CUDA-event times over an identical repeat --------------
    lib.reset_kernel_times()
    lib.set_profiling(True)
    for s in range(args.warmup + args.steps, args.warmup + 2 * args.steps):
        vo.track_pool(s * fps_step, fps_step, want_infos=False)
    lib.set_profiling(False)
    ktimes = lib.kernel_times()
    vo.destroy()
    peak, peak_src = load_peaks()
    mean = lambda k: float(np.mean([i[k] for i in infos]))  # noqa: E731
    n_l, n_r, m_map, m_trk = mean("n_features_left"), mean("n_features_right"), mean("map_points_before"), mean("tracked")
    # algorithmic bytes per launch; a launch covers the stereo pair
    alg = {
        "score_kernel": 2 * W * H,
        "nms_tile_kernel": 2 * W * H,
        "tile_kernel": 12 * 2 * 3900 + 12 * (n_l + n_r) / 0.8,
        # descriptors + the feature index its extra CTA builds (positions in, two CSRs out)
        "brief_kernel": 8 * (n_l + n_r) + 32 * (n_l + n_r) + 57 * 57 * (n_l + n_r) + 2 * 8 * (n_l + n_r) + 2 * 4 * (n_l + n_r),
        "index_kernel": 2 * 8 * (n_l + n_r) + 2 * 4 * (n_l + n_r),
        "mapcand_kernel": 32 * (m_map + n_l) + 8 * n_l + 24 * m_map + 12 * m_map,
        "rowcand_kernel": 2 * (32 + 8) * n_l,
        "track_a_kernel": 4 * 8 * m_map + 12 * m_map + 32 * m_trk,
        "pose_kernel": 12 * 32 * m_trk,
        "track_b_kernel": 68 * m_map + 8 * n_l,
    }
    per_kernel = {}
    for k, (ms, cnt) in ktimes.items():
        if cnt:
            per_kernel[k] = {"avg_us": 1e3 * ms / cnt, "launches": cnt, "share": ms}
    tot = sum(v["share"] for v in per_kernel.values()) or 1.0
    for v in per_kernel.values():
        v["share"] = v["share"] / tot
    dom = max((k for k in per_kernel if k in alg), key=lambda k: per_kernel[k]["share"] * tot, default=None)
    roofline = None
    traffic = load_traffic()
    if dom:
        achieved = alg[dom] / (per_kernel[dom]["avg_us"] * 1e-6) / 1e9
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic.get(dom), "traffic_source": "profiles/traffic.json (ncu --set full, dram read + write per launch)",
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": alg[dom],
                    "avg_launch_us": per_kernel[dom]["avg_us"],
                    "note": "latency/ALU-bound at one stereo pair per launch (1.45 MB of traffic per frame); see DESIGN.md",
                    "per_kernel": {k: {"avg_us": round(v["avg_us"], 2), "share": round(v["share"], 4),
                                       "GBps": round(alg[k] / (v["avg_us"] * 1e-6) / 1e9, 2) if k in alg else None,
                                       "traffic": traffic.get(k)}
                                   for k, v in per_kernel.items()}}

    # ---------------- e2e: lvt_track with host buffers, H2D + D2H inside the timed region -----------
    vo2 = lib.create(params, 1)
    frames = host_frames(stream, 0, n_e2e)
    R = np.zeros((3, 3))
    t = np.zeros(3)
    import ctypes as C
    f64p = C.POINTER(C.c_double)
    u8p = C.POINTER(C.c_uint8)
    track = lib.lib.lvt_track
    Rp, tp = R.ctypes.data_as(f64p), t.ctypes.data_as(f64p)
    ptrs = [(frames[i, 0].ctypes.data_as(u8p), frames[i, 1].ctypes.data_as(u8p)) for i in range(n_e2e)]
    for i in range(args.warmup * fps_step):
        track(vo2.h, ptrs[i][0], ptrs[i][1], H, W, Rp, tp)
    barrier()
    sampler.start()  # second timed region: the samples of both go into one record
    t0 = time.perf_counter()
    for i in range(args.warmup * fps_step, n_e2e):
        track(vo2.h, ptrs[i][0], ptrs[i][1], H, W, Rp, tp)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_ok = vo2.get_state() == capi.STATE_TRACKING
    barrier()
    clocks = sampler.stop()  # sampled during both timed regions (value and e2e)
    e2e_s = max_over_ranks(e2e_s)
    e2e_value = world * n_timed / e2e_s
    gt = stream.ground_truth_t(n_e2e - 1, params.fx, params.baseline)
    drift = float(np.linalg.norm(t - gt))
    vo2.destroy()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(params, stream, CPU_SAMPLE_FRAMES)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic",
                "config": {"workload": "config 2: 1242x375 synthetic stereo stream (SURVEY 8d), 1 sequence per GPU, "
                                       "max_keypoints_per_cell=250 (~2000 keypoints/frame), emergent local map",
                           "frames_per_step": fps_step, "sequences": world,
                           "mean_keypoints_left": round(n_l, 1), "mean_map_points": round(m_map, 1),
                           "mean_tracked": round(m_trk, 1), "frames_lost": lost,
                           "cache": "each frame is read once (inputs: %.0f MB per rank resident in HBM, > 126 MB L2)"
                                    % (n_value * 2 * W * H / 1e6),
                           "timing": "CUDA events around each lvt_track_pool call (first extraction launch .. result copy), "
                                     "summed over steps, max over ranks; wall clock %.1f ms/step" % (wall_ms / args.steps)},
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": fps_step * 2 * W * H,
                        "d2h_bytes_per_step": fps_step * 176, "call": "lvt_track (reference C ABI), host buffers, blocking per frame",
                        "tracking_ok": e2e_ok, "drift_vs_ground_truth_m": drift},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=FRAMES_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
