#!/bin/bash
# one-GPU sweep: the three BASELINE configurations and 1/2/4/8 concurrent sequences per GPU (config 2 and 5)
set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_kitti.json 2> gpurun_out/r2c_kitti.err
python bench.py --config euroc --steps 10 --warmup 3 > gpurun_out/r2c_euroc.json 2> gpurun_out/r2c_euroc.err
python bench.py --config tum --steps 10 --warmup 3 > gpurun_out/r2c_tum.json 2> gpurun_out/r2c_tum.err
for k in 2 4 8; do
  python bench.py --seqs-per-gpu $k --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/r2c_kitti_k$k.json 2> gpurun_out/r2c_kitti_k$k.err
done
for k in 2 4; do
  python bench.py --config euroc --seqs-per-gpu $k --steps 6 --warmup 2 --no-cpu-baseline > gpurun_out/r2c_euroc_k$k.json 2> gpurun_out/r2c_euroc_k$k.err
done
tail -n 3 gpurun_out/r2c_*.err
