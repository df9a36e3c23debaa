"""CSPDarknet backbone with the reference's layout (yolox/models/darknet.py:95-177)."""
from __future__ import annotations

import torch.nn as nn

from .network_blocks import BaseConv, CspLayer, DWConv, Focus, SPPBottleneck, _B200Block


class CspDarknet(_B200Block):
    def __init__(self, dep_mul, wid_mul, out_features=("dark3", "dark4", "dark5"), depthwise=False, act="silu"):
        super().__init__()
        assert out_features, "please provide output features of Darknet"
        self.out_features = out_features
        Conv = DWConv if depthwise else BaseConv
        c = int(wid_mul * 64)
        d = max(round(dep_mul * 3), 1)
        self.stem = Focus(3, c, ksize=3, act=act)

        def stage(cin, cout, n, **kw):
            return [Conv(cin, cout, 3, 2, act=act), CspLayer(cout, cout, n=n, depthwise=depthwise, act=act, **kw)]

        self.dark2 = nn.Sequential(*stage(c, c * 2, d))
        self.dark3 = nn.Sequential(*stage(c * 2, c * 4, d * 3))
        self.dark4 = nn.Sequential(*stage(c * 4, c * 8, d * 3))
        self.dark5 = nn.Sequential(
            Conv(c * 8, c * 16, 3, 2, act=act),
            SPPBottleneck(c * 16, c * 16, activation=act),
            CspLayer(c * 16, c * 16, n=d, shortcut=False, depthwise=depthwise, act=act),
        )

    def _train_forward(self, x):
        outputs = {}
        x = self.stem._train_forward(x)
        outputs["stem"] = x
        for name in ("dark2", "dark3", "dark4", "dark5"):
            for blk in getattr(self, name):
                x = blk._train_forward(x)
            outputs[name] = x
        return {k: v for k, v in outputs.items() if k in self.out_features}

    def lower_image(self, b, img, outs=None):
        """Lower the backbone; ``outs`` optionally maps a stage name to the Feat its output must be
        written into (the PAFPN passes slices of its concat buffers here)."""
        outs = outs or {}
        feats = {}
        x = self.stem.lower_image(b, img, out=outs.get("stem"))
        feats["stem"] = x
        for name in ("dark2", "dark3", "dark4", "dark5"):
            blocks = list(getattr(self, name))
            for i, blk in enumerate(blocks):
                last = i == len(blocks) - 1
                x = blk.lower(b, x, out=outs.get(name) if last else None)
            feats[name] = x
        return feats
