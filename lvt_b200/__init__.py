"""lvt_b200 -- B200-native (sm_100a) per-frame visual-odometry front end behind LVT's C ABI.

`load()` binds lvt_b200/lib/liblvt_b200.so, the CUDA library built by __graft_entry__.build().
There is no CPU path: a missing library or a missing GPU is an error, never a fallback.
"""
import os

from .capi import (KP_DTYPE, SENSOR_RGBD, SENSOR_STEREO, STATE_LOST, STATE_NOT_INITIALIZED, STATE_TRACKING, Context,
                   FrameInfo, Library, LvtError, Params, System)

# LVT_B200_LIB: another build of the same library (A/B runs of the probes against an older build)
LIB_PATH = os.environ.get("LVT_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "liblvt_b200.so")
_lib = None


def load():
    """The product library.  Raises LvtError when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        _lib = Library(LIB_PATH)
        if not _lib.is_gpu:
            raise LvtError("%s is not the CUDA build" % LIB_PATH)
    return _lib
