#!/bin/bash
# final one-GPU measurements of round 2: tests, the three BASELINE configurations, sequences per GPU, ncu launch list + full capture
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_gpu_tests.log 2>&1; tail -3 gpurun_out/r2m_gpu_tests.log
timeout 600 python bench.py > gpurun_out/r2m_kitti.json 2> gpurun_out/r2m_kitti.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2m_kitti_ref.json 2> gpurun_out/r2m_kitti_ref.err
timeout 600 python bench.py --config euroc --steps 10 --warmup 3 > gpurun_out/r2m_euroc.json 2> gpurun_out/r2m_euroc.err
timeout 600 python bench.py --config tum --steps 10 --warmup 3 > gpurun_out/r2m_tum.json 2> gpurun_out/r2m_tum.err
for k in 2 4 8; do
  timeout 600 python bench.py --seqs-per-gpu $k --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/r2m_kitti_k$k.json 2> gpurun_out/r2m_kitti_k$k.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 800 --csv --log-file gpurun_out/launches_r2m.csv \
  python bench.py --steps 2 --warmup 1 --frames-per-step 40 --no-cpu-baseline > gpurun_out/r2m_ncu_bench.log 2>&1
# the report of a --set full capture with sources is > 64 MiB (gpurun's limit for what comes back): export the raw page
# on the box and bring back the CSV
timeout 900 ncu --set full --clock-control none --import-source on -s 330 -c 64 -o /tmp/r2m_frame -f \
  python tools/probe/phase_probe.py > gpurun_out/r2m_frame.log 2>&1
ncu -i /tmp/r2m_frame.ncu-rep --page raw --csv > gpurun_out/r2m_frame.raw.csv 2> gpurun_out/r2m_frame.export.err
tail -n 2 gpurun_out/r2m_*.err
ls -la gpurun_out/r2m_*
