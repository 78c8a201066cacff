import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lvt_b200
from lvt_b200 import configs, synth
lib = lvt_b200.load()
p = configs.make_params("kitti_synth")
st = synth.StereoStream(n_frames=12, seed=0, **configs.CONFIGS["kitti_synth"]["stream"])
vo = lib.create(p, 1)
for t in range(12):
    vo.track(*st.frame(t))
