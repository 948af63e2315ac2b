"""B200-native HigherHRNet-W48 teacher inference + HeatmapParser decode.

Drop-in for the reference's hot path (andres-fr/realtime-pose-estimation):
``PoseHigherResolutionNet`` (rtpe/third_party/pose_higher_hrnet.py:259-686) and
``HeatmapParser`` (rtpe/third_party/group.py:125-287), executed by hand-written
sm_100a CUDA kernels behind the C ABI declared in include/brtpe.h.
"""
from .synth import synth_decode_batch  # noqa: F401
from .heatmap_parser import HeatmapParser, Params  # noqa: F401
from .hhrnet import PoseHigherResolutionNet, BasicBlock, Bottleneck, HighResolutionModule, \
    NoOpModule  # noqa: F401
from .precision import network_to_half, tofp16, tofp32, BN_convert_float, \
    get_hrnet_w48_teacher, W48_KWARGS  # noqa: F401
from .students import AttentionStudent, AttentionStudentSteps, CamStudent, ContextAwareModule, \
    MultistageStudent, RefinerStudent, SELayer, SkipConv, StemHRNet  # noqa: F401
from .engine import eval_student  # noqa: F401
from .teacher_dump import TeacherDumpWriter, TeacherDumper, load_teacher_data, \
    HEATMAPS_ORDER  # noqa: F401
from . import preprocess  # noqa: F401
from . import coco_results  # noqa: F401
from ._lib import BrtpeError, LIB_PATH  # noqa: F401
