"""orbslam2_dualcam_b200 -- B200-native ORB front-end / matcher / bundle-adjustment path of ORB-SLAM2-DualCam.

Host side (Python) above a C-ABI shared library (csrc/ -> lib/liborbslam2_dualcam_b200.so, declared in include/).
All computation is in the sm_100a kernels of that library; there is no CPU fallback.
"""
from .capi import KP_DTYPE, OrbError, lib  # noqa: F401
from .extractor import ORBextractor  # noqa: F401
from .matcher import ORBmatcher  # noqa: F401
from .vocabulary import ORBVocabulary, parse_text_vocabulary  # noqa: F401
from .optimizer import DistributedOptimizer, Optimizer, compact_problem, shard_problem  # noqa: F401
