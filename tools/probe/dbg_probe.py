"""Cycle traces of the map pass (track_a) and of one pose evaluation: run with
LVT_B200_SYNC=1 LVT_B200_TRACKDBG=1 LVT_B200_POSEDBG=1."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lvt_b200
from lvt_b200 import configs, synth
lib = lvt_b200.load()
name = sys.argv[1] if len(sys.argv) > 1 else "kitti_synth"
p = configs.make_params(name)
st = synth.StereoStream(n_frames=12, seed=0, **configs.CONFIGS[name]["stream"])
vo = lib.create(p, 1)
for t in range(12):
    vo.track(*st.frame(t))
print(vo.frame_info())
