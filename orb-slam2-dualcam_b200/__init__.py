"""orbslam2_dualcam_b200 -- B200-native ORB front-end / matcher / bundle-adjustment path of ORB-SLAM2-DualCam.

Host side (Python) above a C-ABI shared library (csrc/ -> liborbslam2_dualcam_b200.so, declared in include/).
"""
from . import synth  # noqa: F401
