"""YOLOPAFPN neck with the reference's layout (yolox/models/yolo_pafpn.py:12-116)."""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn

from .darknet import CspDarknet
from .network_blocks import BaseConv, CspLayer, DWConv, _B200Block


class YoloPafpn(_B200Block):
    def __init__(self, depth: float = 1.0, width: float = 1.0,
                 in_features: Sequence[str] = ("dark3", "dark4", "dark5"),
                 in_channels: Sequence[int] = [256, 512, 1024], depthwise: bool = False, act: str = "silu"):
        super().__init__()
        self.backbone = CspDarknet(depth, width, depthwise=depthwise, act=act)
        self.in_features = in_features
        self.in_channels = in_channels
        Conv = DWConv if depthwise else BaseConv
        c3, c4, c5 = (int(c * width) for c in in_channels)
        n = round(3 * depth)
        self.upsample = nn.Upsample(scale_factor=2, mode="nearest")
        self.lateral_conv0 = BaseConv(c5, c4, 1, 1, act=act)
        self.C3_p4 = CspLayer(2 * c4, c4, n, False, depthwise=depthwise, act=act)
        self.reduce_conv1 = BaseConv(c4, c3, 1, 1, act=act)
        self.C3_p3 = CspLayer(2 * c3, c3, n, False, depthwise=depthwise, act=act)
        self.bu_conv2 = Conv(c3, c3, 3, 2, act=act)
        self.C3_n3 = CspLayer(2 * c3, c4, n, False, depthwise=depthwise, act=act)
        self.bu_conv1 = Conv(c4, c4, 3, 2, act=act)
        self.C3_n4 = CspLayer(2 * c4, c5, n, False, depthwise=depthwise, act=act)

    def _train_forward(self, input):
        feats = self.backbone._train_forward(input)
        x2, x1, x0 = (feats[f] for f in self.in_features)
        fpn_out0 = self.lateral_conv0._train_forward(x0)
        f_out0 = self.C3_p4._train_forward(torch.cat([self.upsample(fpn_out0), x1], 1))
        fpn_out1 = self.reduce_conv1._train_forward(f_out0)
        from .streams import mark_ready

        # the head's level chains fork from the point their input was produced: the stride-8 level (the heaviest) then runs
        # next to the bottom-up half of the neck, like the lanes of the inference plan
        pan_out2 = mark_ready(self.C3_p3._train_forward(torch.cat([self.upsample(fpn_out1), x2], 1)))
        p_out1 = torch.cat([self.bu_conv2._train_forward(pan_out2), fpn_out1], 1)
        pan_out1 = mark_ready(self.C3_n3._train_forward(p_out1))
        p_out0 = torch.cat([self.bu_conv1._train_forward(pan_out1), fpn_out0], 1)
        pan_out0 = self.C3_n4._train_forward(p_out0)
        return (pan_out2, pan_out1, pan_out0)

    def lower_image(self, b, img):
        """All four torch.cat / two nn.Upsample of the reference (yolo_pafpn.py:97-113) disappear:
        producers write straight into channel segments of the consumer's input buffer, and the
        lateral convs store their result twice (once replicated 2x2 into the upsampled segment)."""
        if tuple(self.in_features) != ("dark3", "dark4", "dark5"):
            raise NotImplementedError("the B200 lowering implements the default dark3/dark4/dark5 PAFPN")
        B, H, W = img.shape[0], img.shape[2], img.shape[3]
        c3 = self.reduce_conv1.conv.out_channels
        c4 = self.lateral_conv0.conv.out_channels
        cat_p4 = b.new_feat(B, H // 16, W // 16, [c4, c4])    # [up(fpn_out0) | dark4]
        cat_p3 = b.new_feat(B, H // 8, W // 8, [c3, c3])      # [up(fpn_out1) | dark3]
        cat_n3 = b.new_feat(B, H // 16, W // 16, [c3, c3])    # [bu_conv2(pan_out2) | fpn_out1]
        cat_n4 = b.new_feat(B, H // 32, W // 32, [c4, c4])    # [bu_conv1(pan_out1) | fpn_out0]
        feats = self.backbone.lower_image(b, img, outs={"dark3": cat_p3.seg(1), "dark4": cat_p4.seg(1)})
        x0 = feats["dark5"]
        self.lateral_conv0.lower(b, x0, out=cat_n4.seg(1), ups=cat_p4.seg(0))
        f_out0 = self.C3_p4.lower(b, cat_p4)
        self.reduce_conv1.lower(b, f_out0, out=cat_n3.seg(1), ups=cat_p3.seg(0))
        pan_out2 = self.C3_p3.lower(b, cat_p3)
        self.bu_conv2.lower(b, pan_out2, out=cat_n3.seg(0))
        pan_out1 = self.C3_n3.lower(b, cat_n3)
        self.bu_conv1.lower(b, pan_out1, out=cat_n4.seg(0))
        pan_out0 = self.C3_n4.lower(b, cat_n4)
        return (pan_out2, pan_out1, pan_out0)
