"""B200-native HigherHRNet-W48 teacher inference + HeatmapParser decode.

Drop-in for the reference's hot path (andres-fr/realtime-pose-estimation):
``PoseHigherResolutionNet`` (rtpe/third_party/pose_higher_hrnet.py:259-686) and
``HeatmapParser`` (rtpe/third_party/group.py:125-287), executed by hand-written
sm_100a CUDA kernels behind the C ABI declared in include/brtpe.h.
"""
from .synth import synth_decode_batch  # noqa: F401
