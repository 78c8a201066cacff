#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.
    python tools/summarize_profiles.py <tag>      e.g. r1
reads gpurun_out/launches_<tag>.csv and gpurun_out/<tag>_*.ncu-rep (needs `ncu` on PATH)."""
import csv
import glob
import io
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__cluster_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_bytes.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def launches(tag):
    path = os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % tag)
    if not os.path.exists(path):
        return None
    lines = [ln for ln in open(path, errors="ignore") if ln.startswith('"')]
    rows = list(csv.reader(io.StringIO("".join(lines))))
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        name = r[ik].split("(")[0].replace("lvtb::", "")
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else v  # -> us
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = ["# ncu launch list summary (%s)" % tag, "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <n> python bench.py --steps 2 --warmup 1` (r1: -s 1200 -c 420; r2: -s 400 -c 800, inside the value arm)",
           "(cold-cache, serialised launches: compare SHARES, not absolutes)", "",
           "| kernel | launches | avg us | total us | share |", "|---|---:|---:|---:|---:|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| %s | %d | %.2f | %.1f | %.1f%% |" % (k, n, t / n, t, 100 * t / tot))
    out.append("")
    out.append("total captured device time: %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
    return "\n".join(out) + "\n"


def full(rep, seen, traffic_out):
    """every kernel (launch shape) of a report once -- the LONGEST captured launch of each: the retry-pass launches of
    nms_tile_kernel / tile_kernel exit at once when no image needs them -- with the judged metrics + DRAM traffic"""
    if rep.endswith(".csv"):  # exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv`): the report itself was too big to bring back
        text = open(rep, errors="ignore").read()
    else:
        text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    if len(rows) < 3:
        return None
    hdr, units = rows[0], rows[1]
    best = OrderedDict()
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        name = d.get("Kernel Name", ("?", ""))[0].split("(")[0].replace("lvtb::", "")
        # one function, several jobs (launch shapes): the map pass is cut in two in the batched engine, mapcand_kernel
        # lists map points (296 CTAs) or staged / appended points (148 CTAs)
        grid = d.get("launch__grid_size", ("", ""))[0].replace(",", "").split(".")[0]
        label = {("track_a_kernel", "8"): "track_a_kernel[early part]", ("track_a_kernel", "1"): "track_a_kernel",
                 ("mapcand_kernel", "296"): "mapcand_kernel[early]", ("mapcand_kernel", "148"): "stagedcand_kernel"}.get((name, grid), name)
        try:
            dur = float(d["gpu__time_duration.sum"][0].replace(",", ""))
        except Exception:
            dur = 0.0
        if label not in best or dur > best[label][0]:
            best[label] = (dur, d)
    parts = []
    for name, (_, d) in best.items():
        if name in seen:
            continue
        seen.add(name)
        out = ["## %s  (%s)" % (name, os.path.basename(rep)), "", "| metric | value | unit |", "|---|---:|---|"]
        for k in KEYS:
            if k in d:
                out.append("| %s | %s | %s |" % (k, d[k][0], d[k][1]))
        try:
            rd = float(d["dram__bytes_read.sum"][0].replace(",", ""))
            wr = float(d["dram__bytes_write.sum"][0].replace(",", ""))
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            traffic = rd * mult.get(d["dram__bytes_read.sum"][1], 1) + wr * mult.get(d["dram__bytes_write.sum"][1], 1)
            out.append("| **traffic = dram read + write** | %.0f | byte |" % traffic)
            traffic_out[name] = traffic
        except Exception:
            pass
        parts.append("\n".join(out) + "\n")
    return "\n".join(parts)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(OUT, exist_ok=True)
    s = launches(tag)
    if s:
        open(os.path.join(OUT, "%s_launches.md" % tag), "w").write(s)
        print(s)
    parts = ["# ncu --set full captures (%s): `ncu --set full --clock-control none --import-source on -s <skip> -c <n> "
             "python tools/probe/phase_probe.py` (steady-state frames of lvt_track_pool) and tools/probe/rectify_probe.py\n" % tag]
    seen, traffic = set(), {}
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "%s_*.ncu-rep" % tag)) +
                      glob.glob(os.path.join(ROOT, "gpurun_out", "%s_*.raw.csv" % tag))):
        f = full(rep, seen, traffic)
        if f:
            parts.append(f)
    if len(parts) > 1:
        open(os.path.join(OUT, "%s_ncu_full.md" % tag), "w").write("\n".join(parts))
        pass
    if traffic:
        import json
        # DRAM bytes (read + write) per launch of every captured kernel: bench.py's roofline.traffic
        json.dump({"source": "%s_ncu_full.md" % tag, "dram_bytes_per_launch": traffic},
                  open(os.path.join(OUT, "traffic.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
