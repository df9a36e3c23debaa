"""Model part of the reference's YoloxConfig (yolox/config.py:17-177, 412-469): the fields that
select and build the detection model, and the six named configs. Training/dataloader options of
the reference are outside the hot path and are not mirrored."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Literal, Optional

import torch.nn as nn


@dataclass
class YoloxConfig:
    name: str
    num_classes: int = 80
    depth: float = 1.00
    width: float = 1.00
    depthwise: bool = False
    act: Literal["silu", "relu", "lrelu"] = "silu"
    input_size: tuple = (640, 640)
    test_size: tuple = (640, 640)
    test_conf: float = 0.01
    nmsthre: float = 0.65
    seed: Optional[Any] = None

    @classmethod
    def get_named_config(cls, name: str) -> Optional["YoloxConfig"]:
        return _NAMED_CONFIG.get(name.replace("-", "_"))

    def validate(self):
        h, w = self.input_size
        assert h % 32 == 0 and w % 32 == 0, "input size must be multiples of 32"

    def get_model(self):
        """Same side effects as config.py:159-177: the module is cached on the config object, BN
        eps/momentum are (re)applied, cls/obj biases are re-initialised and train() is returned."""
        from .yolo_head import YoloxHead
        from .yolo_pafpn import YoloPafpn
        from .yolox import YoloxModule

        if getattr(self, "model", None) is None:
            in_channels = [256, 512, 1024]
            backbone = YoloPafpn(self.depth, self.width, in_channels=in_channels, depthwise=self.depthwise, act=self.act)
            head = YoloxHead(self.num_classes, self.width, in_channels=in_channels, depthwise=self.depthwise, act=self.act)
            self.model = YoloxModule(backbone, head)
        for m in self.model.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eps = 1e-3
                m.momentum = 0.03
        self.model.head.initialize_biases(1e-2)
        self.model.train()
        return self.model


def _named(name, depth, width, depthwise=False, size=(640, 640)):
    c = YoloxConfig(name, depth=depth, width=width, depthwise=depthwise, input_size=size, test_size=size)
    return c


_NAMED_CONFIG = {
    c.name: c
    for c in (
        _named("yolox_s", 0.33, 0.50),
        _named("yolox_m", 0.67, 0.75),
        _named("yolox_l", 1.0, 1.0),
        _named("yolox_x", 1.33, 1.25),
        _named("yolox_tiny", 0.33, 0.375, size=(416, 416)),
        _named("yolox_nano", 0.33, 0.25, depthwise=True, size=(416, 416)),
    )
}
