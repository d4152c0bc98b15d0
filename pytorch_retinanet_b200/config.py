"""Hot-path constants, same names and values as the reference's ``retinanet/config.py:27-42,67-87``
(only the entries the dense per-anchor path reads)."""
from typing import List

# anchor generator (config.py:27-42)
ANCHOR_SIZES: List[List[float]] = [[x, x * 2 ** (1 / 3), x * 2 ** (2 / 3)] for x in [32, 64, 128, 256, 512]]
ANCHOR_STRIDES: List[int] = [8, 16, 32, 64, 128]
ANCHOR_ASPECT_RATIOS: List[float] = [0.5, 1.0, 2.0]
ANCHOR_OFFSET: float = 0.0

# box regression / inference (config.py:67-75)
BBOX_REG_WEIGHTS = [1.0, 1.0, 1.0, 1.0]
SCORE_THRES: float = 0.05
NMS_THRES: float = 0.5
MAX_DETECTIONS_PER_IMAGE: int = 100

# matcher (config.py:81-82)
IOU_THRESHOLDS_FOREGROUND: float = 0.5
IOU_THRESHOLDS_BACKGROUND: float = 0.4

# losses (config.py:85-87)
FOCAL_LOSS_GAMMA: float = 2.0
FOCAL_LOSS_ALPHA: float = 0.25
SMOOTH_L1_LOSS_BETA: float = 0.1
