"""learning/training/predict_score.py of the reference: `ScorePredictor` (predict_score.py:117-226) on libfpose.so."""
from Utils import *  # noqa: F401,F403  (the reference module star-imports Utils; drivers rely on the names)
from foundationpose_b200.estimater import ScorePredictor  # noqa: F401
