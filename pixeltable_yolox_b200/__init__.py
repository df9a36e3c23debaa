"""B200-native (sm_100a) YOLOX detection hot path behind the Python API of pixeltable-yolox.

    from pixeltable_yolox_b200 import Yolox, YoloxModule, YoloxProcessor, YoloxConfig

The compute path is hand-written CUDA (tcgen05/TMEM/TMA implicit-GEMM convs, fused decode,
score filter, NMS, SimOTA) behind the C-ABI of include/yx_b200.h; PyTorch provides device memory,
streams and torch.distributed only. There is no CPU fallback.
"""
from .config import YoloxConfig
from .boxes import bboxes_iou, postprocess
from .darknet import CspDarknet
from .losses import IouLoss
from .processor import Detections, YoloxProcessor
from .yolo_head import YoloxHead
from .yolo_pafpn import YoloPafpn
from .yolox import Yolox, YoloxModule

__all__ = ["Yolox", "YoloxModule", "YoloxProcessor", "YoloxHead", "YoloPafpn", "CspDarknet", "YoloxConfig",
           "IouLoss", "Detections", "postprocess", "bboxes_iou"]
__version__ = "0.1.0"
