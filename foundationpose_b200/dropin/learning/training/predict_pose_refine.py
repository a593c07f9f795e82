"""learning/training/predict_pose_refine.py of the reference: `PoseRefinePredictor` (predict_pose_refine.py:92-239)
on libfpose.so."""
from Utils import *  # noqa: F401,F403
from foundationpose_b200.estimater import PoseRefinePredictor  # noqa: F401
