#!/usr/bin/env python
"""Benchmark of the per-frame track() path (BASELINE.json metric: stereo frames/s at 1242x375; ATE vs reference).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)
    python bench.py --config euroc --seqs-per-gpu 2 ...      # BASELINE config 5 (16 sequences on 8 GPUs)
    python bench.py --config tum                             # BASELINE config 3 (640x480 RGB-D)

A "step" is one pass of the hot path over `frames_per_step` consecutive frames of a synthetic stream
(default = config 2: 1242x375, ~2000 keypoints/frame, the local map is the emergent
~2000 points).  N > 1: independent sequences, `--seqs-per-gpu` per GPU (weak scaling), the only collective
is one NCCL broadcast of the calibration block.

`value`  : whole-job frames/s with the frames already resident in HBM (lvt_track_pool), timed on the device
           with CUDA events, max over ranks.
`e2e`    : the same metric through the reference's own C call, lvt_track(handle, left*, right*, ...), with
           HOST buffers -- H2D of both images and D2H of the pose inside the timed region, blocking per frame.
`ate_vs_reference_m`: RMSE over frames of |t_b200 - t_oracle| (same world frame, no alignment) between
           the e2e arm's trajectory and the CPU oracle's on the same frames (rank 0, sequence 0).
"""
import argparse
import ctypes as C
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

WORKLOADS = {
    "kitti": dict(name="kitti_synth", metric="stereo frames/sec at 1242x375", frames_per_step=100,
                  workload="config 2: 1242x375 synthetic stereo stream (SURVEY 8d), max_keypoints_per_cell=250 "
                           "(~2000 keypoints/frame), emergent local map"),
    "euroc": dict(name="euroc_synth", metric="stereo frames/sec at 752x480", frames_per_step=20,
                  workload="config 5: 752x480 EuRoC-shape synthetic stereo stream (SURVEY 8d), agast_threshold=8, "
                           "max_keypoints_per_cell=1020 (~5000 keypoints/frame), emergent local map"),
    "tum": dict(name="tum_synth", metric="RGB-D frames/sec at 640x480", frames_per_step=20,
                workload="config 3: 640x480 synthetic RGB-D stream (SURVEY 8d, TUM shape), max_keypoints_per_cell=1860 "
                         "(~1500 keypoints/frame), triangulation policy 2 (emergent map ~17 000 points)"),
}
CPU_SAMPLE_FRAMES = 400
PROFILE_STEPS = 4
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liblvt_oracle.so")
f64p = C.POINTER(C.c_double)
u8p = C.POINTER(C.c_uint8)
f32p = C.POINTER(C.c_float)


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
    """SM clock and throttle reasons of this rank's GPU during the timed regions (B200_PROFILING.md recipe).  Sampled
    in-process through NVML (two queries every 50 ms); a `nvidia-smi -lms 50` child per rank, as in round 1, costs the
    ranks of an 8-GPU job 4 % of `value` (measured: the host's enqueue time per frame goes from 22 to 39 us while eight
    of them poll the driver).  Falls back to nvidia-smi when the NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []
        self.sm, self.smax, self.reasons = [], [], set()
        self.nvml = self.handle = None
        self.run = threading.Event()
        self.quit = False
        self.source = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                phys = int(vis.split(",")[gpu_index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.source = "NVML, 50 ms"
            threading.Thread(target=self._poll, daemon=True).start()
        except Exception:
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        bits = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        try:
            self.smax.append(float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)))
        except Exception:
            pass
        while not self.quit:
            self.run.wait()
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                for name, bit in bits:
                    if r & int(bit):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def start(self):
        self.paused = False
        if os.environ.get("LVT_BENCH_NO_CLOCKS"):  # A/B aid: does the sampler itself disturb the ranks?
            return
        if self.nvml:
            self.run.set()
            return
        try:
            self.source = "nvidia-smi -lms 50"
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
        """end of one timed region: stop sampling, keep the samples"""
        if self.nvml:
            self.run.clear()
            return
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.proc.wait()
            self.paused = True

    def stop(self):
        if self.nvml:
            self.run.clear()
            self.quit = True
            self.run.set()
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.smax) if self.smax else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}
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
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def make_stream(wl, n_frames, seed):
    cfg = configs.CONFIGS[wl["name"]]
    cls = synth.StereoStream if cfg["sensor"] == 1 else synth.RgbdStream
    return cls(n_frames=n_frames, seed=seed, **cfg["stream"])


def host_frames(stream, first, n, sensor):
    """n frames as host arrays (what a caller of lvt_track holds): stereo [n][2][H][W] u8; RGB-D gray [n][H][W] u8 and
    depth [n][H][W] f32"""
    if sensor == 1:
        out = np.empty((n, 2, stream.H, stream.W), np.uint8)
        for i in range(n):
            out[i, 0], out[i, 1] = stream.frame(first + i)
        return out, None
    gray = np.empty((n, stream.H, stream.W), np.uint8)
    depth = np.empty((n, stream.H, stream.W), np.float32)
    for i in range(n):
        gray[i], depth[i] = stream.frame(first + i)
    return gray, depth


def frame_caller(lib, vo, sensor, frames, depth, poses):
    """bind lvt_track / lvt_track_rgbd on host buffers; pose of frame i lands in poses[i] (R row-major, t)"""
    fn = lib.lib.lvt_track if sensor == 1 else lib.lib.lvt_track_rgbd
    H, W = frames.shape[-2:]
    h = vo.h
    args = []
    for i in range(len(frames)):
        Rp = C.cast(poses[i].ctypes.data, f64p)
        tp = C.cast(poses[i, 9:].ctypes.data, f64p)
        if sensor == 1:
            args.append((h, frames[i, 0].ctypes.data_as(u8p), frames[i, 1].ctypes.data_as(u8p), H, W, Rp, tp))
        else:
            args.append((h, frames[i].ctypes.data_as(u8p), depth[i].ctypes.data_as(f32p), H, W, Rp, tp))
    return fn, args


def oracle_run(params, sensor, stream, n_frames, skip):
    """the oracle (CPU port of the reference path) on the first `n_frames` frames of `stream` (the very frames the
    b200 arms track): wall clock around track() only, as examples/kitti/kitti_example.cpp:129-131 does; the first
    `skip` frames are warm-up"""
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    orc = capi.Library(ORACLE_SO)
    vo = orc.create(params, sensor)
    frames, depth = host_frames(stream, 0, n_frames, sensor)
    poses = np.zeros((n_frames, 12))
    fn, args = frame_caller(orc, vo, sensor, frames, depth, poses)
    infos, dt = [], 0.0
    for i in range(n_frames):
        t0 = time.perf_counter()
        fn(*args[i])
        t1 = time.perf_counter()
        if i >= skip:
            dt += t1 - t0
        infos.append(vo.frame_info())
    ok = vo.get_state() == capi.STATE_TRACKING
    vo.destroy()
    return (n_frames - skip) / dt, poses, infos, ok


def run_reference(args, wl):
    """--impl reference: the reference's CPU path (oracle port; the genuine sources cannot be built here,
    DESIGN.md) on all the host threads it can use: every sequence of the job in parallel, 2 threads each,
    on the same frames and frames_per_step as the b200 arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_seq = args.gpus * args.seqs_per_gpu
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    orc = capi.Library(ORACLE_SO)
    sensor = configs.CONFIGS[wl["name"]]["sensor"]
    params = configs.make_params(wl["name"])
    fps_step = args.frames_per_step
    total = (args.warmup + args.steps) * fps_step
    vos, calls = [], []
    for s in range(n_seq):
        # the canvas depends on the stream length: build it as the b200 arm does, so that both arms see the same frames
        frames, depth = host_frames(make_stream(wl, stream_length(args, args.seqs_per_gpu), s), 0, total, sensor)
        vo = orc.create(params, sensor)
        vos.append(vo)
        poses = np.zeros((total, 12))  # kept alive with the call arguments: the C side writes R, t into it
        calls.append(frame_caller(orc, vo, sensor, frames, depth, poses) + (frames, depth, poses))
    cores = os.cpu_count() or 1
    workers = max(1, min(n_seq, cores // 2))

    def run_all(lo, hi):
        pending = list(range(n_seq))
        lock = threading.Lock()

        def worker():
            while True:
                with lock:
                    if not pending:
                        return
                    s = pending.pop()
                fn, a = calls[s][0], calls[s][1]
                for i in range(lo, hi):
                    fn(*a[i])  # ctypes releases the GIL inside the call
        th = [threading.Thread(target=worker) for _ in range(workers)]
        [t.start() for t in th]
        [t.join() for t in th]

    run_all(0, args.warmup * fps_step)
    t0 = time.perf_counter()
    run_all(args.warmup * fps_step, total)
    dt = time.perf_counter() - t0
    ok = all(v.get_state() == capi.STATE_TRACKING for v in vos)
    value = n_seq * args.steps * fps_step / dt
    line = {"impl": "reference", "metric": wl["metric"], "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": bench_config(wl, fps_step, args.gpus, args.seqs_per_gpu),
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": min(cores, 2 * workers), "kind": "port",
                             "sample": "%d sequence(s) x %d frames (the b200 arm's timed frames), %d at a time, 2 threads "
                                       "each (left/right extraction, lvt_image_features_handler.cpp:204-206), %d host "
                                       "cores, tracking=%s" % (n_seq, args.steps * fps_step, workers, cores, ok)},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def stream_length(args, K):
    """frames of the synthetic stream (its canvas depends on it): warm-up + timed + profiled repeat of the b200 arm"""
    prof_steps = min(args.steps, PROFILE_STEPS) if K == 1 else 0
    return (args.warmup + args.steps + prof_steps) * args.frames_per_step


def bench_config(wl, fps_step, gpus, seqs_per_gpu):
    """identical in both arms"""
    return {"workload": wl["workload"], "frames_per_step": fps_step, "sequences": gpus * seqs_per_gpu,
            "sequences_per_gpu": seqs_per_gpu}


def algorithmic_bytes(W, H, sensor, st):
    """SURVEY 8d / DESIGN.md section 4: algorithmic bytes per launch.  A launch covers the frame's images
    (2 for stereo).  st: measured workload statistics."""
    imgs = 2 if sensor == 1 else 1
    n_l, n_r, m_map, m_trk, staged = st["n_l"], st["n_r"], st["map"], st["tracked"], st["staged"]
    feats = n_l + n_r
    return {
        "score_kernel": imgs * W * H,                                   # detect: every pixel read once
        "nms_tile_kernel": imgs * 12 * st["n_pre"],                      # NMS: 12 B per pre-NMS corner
        "tile_kernel": imgs * 12 * st["n_post"] + 12 * feats / st["border_keep"],  # ANMS: survivors in, kept out
        "brief_kernel": imgs * W * H + 8 * feats / st["border_keep"] + 32 * feats,  # image + keypoints in, descriptors out
        "index_kernel": 8 * feats + 2 * 4 * feats,
        "mapcand_kernel": 32 * (m_map + n_l) + 8 * n_l + 24 * m_map + 12 * m_map,
        "stagedcand_kernel": 32 * (staged + n_l) + 8 * n_l + 24 * staged + 12 * staged,
        "rowcand_kernel": 2 * (32 + 8) * n_l,
        "track_a_kernel": 44 * m_map + 32 * m_trk,
        # the batched engine cuts the map pass in two: the rounds over the lists (early part) / book-keeping + solver inputs
        "track_a_kernel[early part]": 44 * m_map,
        "track_a_kernel[rest]": 32 * m_trk,
        "mapcand_kernel[early]": 32 * (m_map + n_l) + 8 * n_l + 24 * m_map + 12 * m_map,
        "mapcand_kernel[appended points]": 32 * n_l + 8 * n_l,
        "pose_kernel": st["lm_evals"] * 32 * m_trk,                     # per evaluation: xyz f64 x3 + uv f32 x2
        "track_b_kernel": 68 * m_map + 8 * n_l,
        "depth_gate_kernel": 12 * n_l,
    }


def run_b200(args, wl):
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
    name = wl["name"]
    sensor = configs.CONFIGS[name]["sensor"]
    K = args.seqs_per_gpu

    # the only collective: broadcast of the calibration / parameter block from rank 0 over NCCL
    p0 = configs.make_params(name)
    blob = torch.from_numpy(p0.to_array() if rank == 0 else np.zeros_like(p0.to_array())).cuda()
    if world > 1:
        dist.broadcast(blob, src=0)
    params = capi.Params.from_array(blob.cpu().numpy())

    fps_step = args.frames_per_step
    prof_steps = min(args.steps, PROFILE_STEPS) if K == 1 else 0
    n_value = (args.warmup + args.steps + prof_steps) * fps_step  # warm-up, timed, profiled repeat
    n_e2e = (args.warmup + args.steps) * fps_step
    n_timed = args.steps * fps_step
    seeds = [rank * K + k for k in range(K)]
    assert n_value == stream_length(args, K) and n_e2e <= n_value
    streams = [make_stream(wl, n_value, s) for s in seeds]
    H, W = streams[0].H, streams[0].W

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

    def in_threads(fn):
        """fn(k) for the K sequences of this rank, one host thread each (ctypes releases the GIL)"""
        if K == 1:
            return [fn(0)]
        out, err = [None] * K, []

        def work(k):
            try:
                torch.cuda.set_device(local)
                out[k] = fn(k)
            except Exception as e:  # noqa: BLE001
                err.append(repr(e))
        th = [threading.Thread(target=work, args=(k,)) for k in range(K)]
        [t.start() for t in th]
        [t.join() for t in th]
        if err:
            raise RuntimeError("; ".join(err))
        return out

    # ---------------- value: frames resident in HBM, pipelined, device-timed ----------------------
    resident = True
    vos = [lib.create(params, sensor) for _ in range(K)]
    value_poses = [np.zeros((0, 12)) for _ in range(K)]
    value_infos = [[] for _ in range(K)]
    if resident:
        def upload(k):
            vos[k].pool_reserve(n_value)
            for s0 in range(0, n_value, fps_step):
                chunk, dep = host_frames(streams[k], s0, fps_step, sensor)
                for i in range(fps_step):
                    if sensor == 1:
                        vos[k].pool_upload(s0 + i, chunk[i, 0], chunk[i, 1])
                    else:
                        vos[k].pool_upload(s0 + i, chunk[i], dep[i])
        in_threads(upload)

        def steps(k, lo, hi):
            ms = 0.0
            for s in range(lo, hi):
                poses, inf = vos[k].track_pool(s * fps_step, fps_step)
                ms += vos[k].last_batch_ms()
                value_poses[k] = np.concatenate([value_poses[k], poses])
                value_infos[k] += inf
            return ms
        in_threads(lambda k: steps(k, 0, args.warmup))
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    host_t = (C.c_double * 4)()
    for vo in vos:
        lib.lib.lvt_debug_host_times(C.c_void_p(vo.h), host_t, 1)  # reset: host time of the timed calls only
    launches0 = lib.launch_count()
    dev_ms = wall_ms = 0.0
    if resident:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        t0 = time.perf_counter()
        per_seq_ms = in_threads(lambda k: steps(k, args.warmup, args.warmup + args.steps))
        torch.cuda.synchronize()
        wall_ms = 1e3 * (time.perf_counter() - t0)
        ev1.record()
        ev1.synchronize()
        # one sequence: the sum of the per-call device times (CUDA events around each lvt_track_pool call on its own
        # stream); several concurrent sequences: the device-side span around all of them
        dev_ms = per_seq_ms[0] if K == 1 else ev0.elapsed_time(ev1)
    launches = lib.launch_count() - launches0
    lib.lib.lvt_debug_host_times(C.c_void_p(vos[0].h), host_t, 0)
    host_enqueue_us = max_over_ranks(host_t[1] / max(1, args.steps * fps_step))
    host_wait_us = max_over_ranks(host_t[2] / max(1, args.steps * fps_step))
    barrier()
    sampler.pause()
    dev_ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks(wall_ms)

    # ---------------- roofline: per-kernel CUDA-event times over a short identical repeat ----------
    ktimes, ktimes_single, lm_evals = {}, {}, []
    if resident and prof_steps:
        lib.reset_kernel_times()
        lib.set_profiling(True)
        for s in range(args.warmup + args.steps, args.warmup + args.steps + prof_steps):
            vos[0].track_pool(s * fps_step, fps_step, want_infos=False)
            lm_evals += [vos[0].frame_counters(i)["lm_evaluations"] for i in range(fps_step)]
        lib.set_profiling(False)
        ktimes = lib.kernel_times()
        # the same kernels at one frame per extraction launch ("report both single-frame and batched")
        os.environ["LVT_B200_GROUP"] = "1"
        single = lib.create(params, sensor)
        single.pool_reserve(2 * fps_step)
        chunk, dep = host_frames(streams[0], 0, 2 * fps_step, sensor)
        for i in range(2 * fps_step):
            single.pool_upload(i, chunk[i, 0] if sensor == 1 else chunk[i], chunk[i, 1] if sensor == 1 else dep[i])
        single.track_pool(0, fps_step, want_infos=False)
        del os.environ["LVT_B200_GROUP"]
        lib.reset_kernel_times()
        lib.set_profiling(True)
        single.track_pool(fps_step, fps_step, want_infos=False)
        lib.set_profiling(False)
        ktimes_single = lib.kernel_times()
        single.destroy()
    for vo in vos:
        vo.destroy()

    # ---------------- e2e: lvt_track with host buffers, H2D + D2H inside the timed region -----------
    vos2 = [lib.create(params, sensor) for _ in range(K)]
    e2e_poses = [np.zeros((n_e2e, 12)) for _ in range(K)]
    callers = []
    for k in range(K):
        frames, depth = host_frames(streams[k], 0, n_e2e, sensor)
        # the caller's frames live in page-locked host memory (lvt_alloc_pinned): the H2D copy inside the call is
        # one DMA per image straight from these buffers; the pageable variant (staged) is reported next to it
        pf = lib.pinned_empty(frames.shape, frames.dtype)
        pf[...] = frames
        pd = None
        if depth is not None:
            pd = lib.pinned_empty(depth.shape, depth.dtype)
            pd[...] = depth
        callers.append(frame_caller(lib, vos2[k], sensor, pf, pd, e2e_poses[k]) + (frames, depth, pf, pd))

    def e2e_range(k, lo, hi):
        fn, a = callers[k][0], callers[k][1]
        for i in range(lo, hi):
            fn(*a[i])
    in_threads(lambda k: e2e_range(k, 0, args.warmup * fps_step))
    barrier()
    sampler.start()  # second timed region: the samples of both go into one record
    launches1 = lib.launch_count()
    t0 = time.perf_counter()
    in_threads(lambda k: e2e_range(k, args.warmup * fps_step, n_e2e))
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_launches = lib.launch_count() - launches1
    e2e_ok = all(v.get_state() == capi.STATE_TRACKING and v.last_status() == 0 for v in vos2)
    barrier()
    clocks = sampler.stop()  # sampled during both timed regions (value and e2e)
    e2e_s = max_over_ranks(e2e_s)
    e2e_value = world * K * n_timed / e2e_s
    _ = e2e_launches
    value = world * K * n_timed / (dev_ms * 1e-3)
    drift = None
    if sensor == 1:
        gt = streams[0].ground_truth_t(n_e2e - 1, params.fx, params.baseline)
        drift = float(np.linalg.norm(e2e_poses[0][-1, 9:] - gt))
    e2e_infos_last = [v.frame_info() for v in vos2]
    for vo in vos2:
        vo.destroy()
    # the same blocking calls on PAGEABLE buffers (staged through pinned memory inside the call), sequence 0 only
    vp = lib.create(params, sensor)
    pg_poses = np.zeros((n_e2e, 12))
    fn_pg, a_pg = frame_caller(lib, vp, sensor, callers[0][2], callers[0][3], pg_poses)
    for i in range(args.warmup * fps_step):
        fn_pg(*a_pg[i])
    t0 = time.perf_counter()
    for i in range(args.warmup * fps_step, n_e2e):
        fn_pg(*a_pg[i])
    torch.cuda.synchronize()
    e2e_pageable = n_timed / (time.perf_counter() - t0)
    e2e_pageable_same = bool(np.array_equal(pg_poses, e2e_poses[0]))
    vp.destroy()

    # ---------------- e2e, look-ahead: lvt_track_batch on the same host buffers, one call per step -----------
    def batch_arm(pinned):
        vs = [lib.create(params, sensor) for _ in range(K)]
        bufs = []
        for k in range(K):
            fr, dp = callers[k][2], callers[k][3]
            if pinned:
                pf = lib.pinned_empty(fr.shape, fr.dtype)
                pf[...] = fr
                pd = None
                if dp is not None:
                    pd = lib.pinned_empty(dp.shape, dp.dtype)
                    pd[...] = dp
                fr, dp = pf, pd
            bufs.append((fr, dp))
        out = [np.zeros((0, 12)) for _ in range(K)]

        def run(k, lo, hi):
            fr, dp = bufs[k]
            for s in range(lo, hi):
                sl = slice(s * fps_step, (s + 1) * fps_step)
                if sensor == 1:
                    p, _ = vs[k].track_batch(list(fr[sl, 0]), list(fr[sl, 1]), want_infos=False)
                else:
                    p, _ = vs[k].track_batch(list(fr[sl]), list(dp[sl]), want_infos=False)
                out[k] = np.concatenate([out[k], p])
        in_threads(lambda k: run(k, 0, args.warmup))
        barrier()
        t0 = time.perf_counter()
        in_threads(lambda k: run(k, args.warmup, args.warmup + args.steps))
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        ok = all(v.get_state() == capi.STATE_TRACKING for v in vs)
        same = bool(np.array_equal(out[0], e2e_poses[0]))
        for v in vs:
            v.destroy()
        return world * K * n_timed / dt, ok, same
    batch_value, batch_ok, batch_same = batch_arm(False)
    batch_pinned_value, _, batch_pinned_same = batch_arm(True)

    # ---------------- workload statistics, roofline bookkeeping -----------------------------------
    infos = value_infos[0][args.warmup * fps_step:] if resident else e2e_infos_last
    mean = lambda k: float(np.mean([i[k] for i in infos]))  # noqa: E731
    st = {"n_l": mean("n_features_left"), "n_r": mean("n_features_right"), "map": mean("map_points_before"),
          "tracked": mean("tracked"), "staged": mean("staged_before"),
          "lm_evals": float(np.mean(lm_evals)) if lm_evals else None}
    lost = sum(1 for i in infos if i["state"] != capi.STATE_TRACKING)
    roofline = None
    if rank == 0 and ktimes:
        # corner counts before / after AGAST's NMS on one frame (whole image as one tile, the seam call)
        ctx = lib.context(params)
        img = streams[0].frame(n_e2e // 2)[0]
        st["n_pre"] = float(len(ctx.agast(img, params.agast_threshold, nonmax=False)))
        st["n_post"] = float(len(ctx.agast(img, params.agast_threshold, nonmax=True)))
        ctx.destroy()
        st["border_keep"] = (W - 56.0) * (H - 56.0) / (W * H)
        if st["lm_evals"] is None:
            st["lm_evals"] = 22.0
        alg = algorithmic_bytes(W, H, sensor, st)
        peak, peak_src = load_peaks()
        traffic = load_traffic()
        def table(kt):
            pk = {k: {"avg_us": 1e3 * ms / cnt, "launches": cnt, "ms": ms} for k, (ms, cnt) in kt.items() if cnt}
            fr = max(1, pk.get("pose_kernel", pk.get("track_b_kernel", {"launches": 1}))["launches"])
            for k, v in pk.items():
                v["launches_per_frame"] = v["launches"] / fr
                v["us_per_frame"] = 1e3 * v["ms"] / fr
                # `alg` is per frame.  The kernels of the tracking chain process one frame per launch (the regular
                # mapcand_kernel only runs for the first frame of a call); an extraction launch covers
                # 1 / launches_per_frame frames
                per_frame = k.startswith(("track_", "mapcand", "rowcand", "pose_", "stagedcand", "depth_gate"))
                v["alg_per_launch"] = (alg[k] if per_frame else alg[k] / v["launches_per_frame"]) if k in alg else None
            return pk, fr
        if "track_a_kernel[early part]" in ktimes and ktimes["track_a_kernel[early part]"][1]:
            # with the early part launched separately, "track_a_kernel" is the rest of the pass
            alg = dict(alg)
            alg["track_a_kernel"] = alg["track_a_kernel[rest]"]
        per_kernel, frames_prof = table(ktimes)
        per_kernel_single, _ = table(ktimes_single) if ktimes_single else ({}, 1)
        tot = sum(v["ms"] for v in per_kernel.values()) or 1.0
        dom = max((k for k in per_kernel if k in alg), key=lambda k: per_kernel[k]["ms"], default=None)
        if dom:
            achieved = per_kernel[dom]["alg_per_launch"] / (per_kernel[dom]["avg_us"] * 1e-6) / 1e9
            frame_bytes = sum(alg[k] for k in per_kernel if k in alg)
            roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic.get(dom),
                        "traffic_source": "profiles/traffic.json (ncu --set full, dram read + write per launch)",
                        "peak_source": peak_src, "algorithmic_bytes_per_launch": per_kernel[dom]["alg_per_launch"],
                        "frames_per_extraction_launch": round(1.0 / per_kernel["score_kernel"]["launches_per_frame"], 2)
                        if "score_kernel" in per_kernel else None,
                        "avg_launch_us": per_kernel[dom]["avg_us"],
                        "lm_evaluations_per_frame": st["lm_evals"],
                        "frame": {"algorithmic_bytes": frame_bytes, "GBps": frame_bytes * value / world / K / 1e9,
                                  "frac": frame_bytes * value / world / K / 1e9 / peak,
                                  "note": "all kernels of one frame / time per frame of `value` (one sequence)"},
                        "note": "latency/ALU-bound at one stereo pair per launch; see DESIGN.md section 4",
                        "per_kernel": {k: {"avg_us": round(v["avg_us"], 2), "share": round(v["ms"] / tot, 4),
                                           "launches_per_frame": round(v["launches_per_frame"], 3),
                                           "alg_bytes_per_launch": round(v["alg_per_launch"]) if k in alg else None,
                                           "GBps": round(v["alg_per_launch"] / (v["avg_us"] * 1e-6) / 1e9, 2) if k in alg else None,
                                           "traffic": traffic.get(k)}
                                       for k, v in per_kernel.items()},
                        "per_kernel_one_frame_per_launch": {
                            k: {"avg_us": round(v["avg_us"], 2), "launches_per_frame": round(v["launches_per_frame"], 3),
                                "GBps": round(v["alg_per_launch"] / (v["avg_us"] * 1e-6) / 1e9, 2) if k in alg else None}
                            for k, v in per_kernel_single.items()}}

    # ---------------- CPU baseline + ATE vs the reference path (rank 0, sequence 0) ---------------
    cpu = ate = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cpu = min(CPU_SAMPLE_FRAMES, n_e2e)
        cpu_fps, o_poses, o_infos, ok = oracle_run(params, sensor, streams[0], n_cpu, 10)
        cpu = {"value": cpu_fps, "unit": "frames/s", "cores": 2 if sensor == 1 else 1, "kind": "port",
               "sample": "the first %d frames of sequence 0 after 10 warm-up frames, 1 sequence, %s, tracking=%s"
                         % (n_cpu, "2 threads (left/right extraction, as lvt_image_features_handler.cpp:204-206)"
                            if sensor == 1 else "1 thread", ok)}
        d = e2e_poses[0][:n_cpu, 9:] - o_poses[:, 9:]
        ate = {"ate_vs_reference_m": float(np.sqrt(np.mean(np.sum(d * d, axis=1)))),
               "max_translation_diff_m": float(np.abs(d).max()),
               "max_rotation_diff": float(np.abs(e2e_poses[0][:n_cpu, :9] - o_poses[:, :9]).max()),
               "frames_compared": n_cpu, "bound_m": 1e-3,
               "what": "e2e arm (lvt_track, host buffers) vs the CPU oracle on the same frames; RMSE of |t - t_ref|, no alignment"}
        if resident:
            m = min(n_cpu, len(value_infos[0]))
            ate["frame_info_mismatches"] = sum(1 for a, b in zip(value_infos[0][:m], o_infos[:m]) if a != b)
            dv = value_poses[0][:m, 9:] - o_poses[:m, 9:]
            ate["value_arm_ate_vs_reference_m"] = float(np.sqrt(np.mean(np.sum(dv * dv, axis=1))))
            ate["frame_infos_compared"] = m

    if rank == 0:
        imgs_bytes = (2 * W * H) if sensor == 1 else (W * H * 5)
        line = {"metric": wl["metric"], "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": bench_config(wl, fps_step, world, K),
                "workload_stats": {"mean_keypoints_left": round(st["n_l"], 1), "mean_map_points": round(st["map"], 1),
                                   "mean_tracked": round(st["tracked"], 1), "frames_lost": lost,
                                   "cache": "each frame is read once (inputs: %.0f MB per sequence resident in HBM, > 126 MB L2)"
                                            % (n_value * imgs_bytes / 1e6),
                                   "timing": ("CUDA events around each lvt_track_pool call (first extraction launch .. result copy), "
                                              "summed over steps" if K == 1 else
                                              "CUDA events around the %d concurrent sequences of the rank" % K) +
                                             ", max over ranks; wall clock %.1f ms/step" % (wall_ms / args.steps),
                                   "frames_per_extraction_launch": "4 (lvt_track_pool / lvt_track_batch group size)",
                                   "host_us_per_frame": {"enqueue": round(host_enqueue_us, 1), "waiting": round(host_wait_us, 1),
                                                         "note": "host thread of one sequence inside lvt_track_pool, max over ranks"}},
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": K * fps_step * imgs_bytes,
                        "d2h_bytes_per_step": K * fps_step * 176,
                        "call": ("lvt_track" if sensor == 1 else "lvt_track_rgbd") + " (reference C ABI), page-locked host buffers, blocking per frame",
                        "pageable_buffers_value_one_sequence": e2e_pageable, "pageable_poses_identical": e2e_pageable_same,
                        "tracking_ok": e2e_ok, "drift_vs_ground_truth_m": drift,
                        "batch": {"value": batch_value, "unit": "frames/s",
                                  "call": "lvt_track_batch%s, %d frames per call, pageable host buffers: staging, H2D and "
                                          "extraction of later frames run behind the tracking of earlier ones"
                                          % ("" if sensor == 1 else "_rgbd", fps_step),
                                  "tracking_ok": batch_ok, "poses_identical_to_blocking_calls": batch_same,
                                  "page_locked_buffers_value": batch_pinned_value,
                                  "page_locked_poses_identical": batch_pinned_same}},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        if ate:
            line.update(ate)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="kitti", choices=sorted(WORKLOADS))
    ap.add_argument("--seqs-per-gpu", type=int, default=1)
    ap.add_argument("--frames-per-step", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    wl = WORKLOADS[args.config]
    if args.frames_per_step <= 0:
        args.frames_per_step = wl["frames_per_step"]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
