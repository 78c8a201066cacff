"""bench.py's contract with the driver, as far as it can be checked without a GPU: the reference arm (the CPU oracle on the
b200 arm's frames) prints ONE JSON line with the keys the driver reads, under torchrun only rank 0 prints, and the
configuration block is the one the b200 arm prints."""
import json
import os
import subprocess
import sys

from helpers import ROOT

import bench  # noqa: E402  (repo root is on sys.path through helpers)


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_prints_one_line_with_the_contract_keys():
    lines = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--frames-per-step", "5"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "stereo frames/sec at 1242x375" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["dtype"] == "u8"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same configuration block as the b200 arm prints for these arguments
    wl = bench.WORKLOADS["kitti"]
    assert d["config"] == bench.bench_config(wl, 5, 1, 1)
    assert "tracking=True" in d["cpu_baseline"]["sample"]


def test_reference_arm_other_ranks_stay_silent():
    assert _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--frames-per-step", "3"],
                env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_reference_arm_other_configurations():
    for cfg, metric in (("euroc", "stereo frames/sec at 752x480"), ("tum", "RGB-D frames/sec at 640x480")):
        lines = _run(["--impl", "reference", "--config", cfg, "--steps", "1", "--warmup", "0", "--frames-per-step", "3"])
        d = json.loads(lines[0])
        assert d["metric"] == metric and d["value"] > 0 and d["config"]["workload"].startswith("config")


def test_algorithmic_bytes_follow_survey_8d():
    st = dict(n_l=2000.0, n_r=2000.0, map=2100.0, tracked=1850.0, staged=100.0, lm_evals=12.0, n_pre=50000.0, n_post=6000.0,
              border_keep=0.8)
    alg = bench.algorithmic_bytes(1242, 375, 1, st)
    assert alg["score_kernel"] == 2 * 1242 * 375                      # W*H per image
    assert alg["pose_kernel"] == 12 * 32 * 1850                       # 32*M per LM evaluation
    assert alg["brief_kernel"] == 2 * 1242 * 375 + 8 * 4000 / 0.8 + 32 * 4000   # W*H + 8N + 32N' per image
    assert alg["track_a_kernel"] == alg["track_a_kernel[early part]"] + alg["track_a_kernel[rest]"]
