"""The N > 1 path on CPU: world_size-2 gloo, calibration broadcast, sequence sharding, gather.
The per-rank tracking here is driven through the same binding with the CPU oracle standing in
for a GPU (there is none in this container); the sharding / collective logic is what is tested."""
import os
import subprocess
import sys

from helpers import ROOT

WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["LVT_ROOT"])
from lvt_b200 import capi, configs, dist as ldist, synth

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["LVT_PORT"],
                        rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
lib = capi.Library(os.path.join(os.environ["LVT_ROOT"], "oracle", "_build", "liblvt_oracle.so"))
# only rank 0 knows the calibration; the others start from defaults
p = configs.make_params("kitti_synth") if rank == 0 else lib.default_params()
p = ldist.broadcast_params(p, src=0)
assert abs(p.fx - 718.856) < 1e-3 and p.img_width == 1242 and p.max_keypoints_per_cell == 250
mine = ldist.shard_sequences(4, rank, world)
assert mine == [rank, rank + 2]
local = {}
for s in mine:
    st = synth.StereoStream(n_frames=3, seed=s, **configs.CONFIGS["kitti_synth"]["stream"])
    vo = lib.create(p, 1)
    poses = []
    for t in range(3):
        R, tt = vo.track(*st.frame(t))
        poses.append(np.concatenate([R.ravel(), tt]))
    local[s] = np.array(poses)
allp = ldist.gather_trajectories(local)
assert sorted(allp) == [0, 1, 2, 3]
for s in range(4):
    assert allp[s].shape == (3, 12) and 0.3 < allp[s][2, 9] < 1.2  # ~0.43 m per frame along +x
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_ranks_gloo(tmp_path, oracle):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LVT_ROOT=ROOT, LVT_PORT="29541", MASTER_ADDR="127.0.0.1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o


def test_shard_and_param_roundtrip(oracle):
    from lvt_b200 import capi, configs, dist as ldist
    assert ldist.shard_sequences(16, 3, 8) == [3, 11]
    assert sum(len(ldist.shard_sequences(16, r, 8)) for r in range(8)) == 16
    p = configs.make_params("euroc_synth")
    q = capi.Params.from_array(p.to_array())
    assert q.as_dict() == p.as_dict()
