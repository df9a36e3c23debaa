"""torch-CPU fp32 restatement (test infrastructure) of the reference's eval forward:
CspDarknet -> YoloPafpn -> YoloxHead -> decode_outputs, driven directly by a reference-layout
state_dict. Files restated: yolox/models/network_blocks.py:27-208, darknet.py:95-177,
yolo_pafpn.py:83-116, yolo_head.py:140-251. BN is applied unfolded (eps 1e-3, config.py:165) so
the arithmetic is the reference's own, not the folded form the CUDA path uses."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-3
ACTS = {"silu": F.silu, "relu": F.relu, "lrelu": lambda x: F.leaky_relu(x, 0.1)}


def _has(sd, prefix):
    return any(k.startswith(prefix) for k in sd)


def base_conv(sd, p, x, stride, act):
    w = sd[p + ".conv.weight"]
    k = w.shape[-1]
    groups = x.shape[1] // w.shape[1]
    y = F.conv2d(x, w, None, stride, (k - 1) // 2, 1, groups)
    y = F.batch_norm(y, sd[p + ".bn.running_mean"], sd[p + ".bn.running_var"], sd[p + ".bn.weight"],
                     sd[p + ".bn.bias"], False, 0.0, BN_EPS)
    return act(y)


def conv_any(sd, p, x, stride, act):
    """BaseConv or DWConv (depthwise 3x3 then pointwise), told apart by the keys present."""
    if _has(sd, p + ".dconv."):
        return base_conv(sd, p + ".pconv", base_conv(sd, p + ".dconv", x, stride, act), 1, act)
    return base_conv(sd, p, x, stride, act)


def csp_layer(sd, p, x, shortcut, act):
    x1 = base_conv(sd, p + ".conv1", x, 1, act)
    x2 = base_conv(sd, p + ".conv2", x, 1, act)
    i = 0
    while _has(sd, f"{p}.m.{i}."):
        y = conv_any(sd, f"{p}.m.{i}.conv2", base_conv(sd, f"{p}.m.{i}.conv1", x1, 1, act), 1, act)
        x1 = y + x1 if shortcut else y
        i += 1
    return base_conv(sd, p + ".conv3", torch.cat((x1, x2), 1), 1, act)


def spp(sd, p, x, act):
    x = base_conv(sd, p + ".conv1", x, 1, act)
    x = torch.cat([x] + [F.max_pool2d(x, k, 1, k // 2) for k in (5, 9, 13)], 1)
    return base_conv(sd, p + ".conv2", x, 1, act)


def focus(sd, p, x, act):
    tl, tr = x[..., ::2, ::2], x[..., ::2, 1::2]
    bl, br = x[..., 1::2, ::2], x[..., 1::2, 1::2]
    return base_conv(sd, p + ".conv", torch.cat((tl, bl, tr, br), 1), 1, act)


def backbone(sd, x, act, p="backbone.backbone"):
    x = focus(sd, p + ".stem", x, act)
    feats = {}
    for name in ("dark2", "dark3", "dark4"):
        x = conv_any(sd, f"{p}.{name}.0", x, 2, act)
        x = csp_layer(sd, f"{p}.{name}.1", x, True, act)
        feats[name] = x
    x = conv_any(sd, p + ".dark5.0", x, 2, act)
    x = spp(sd, p + ".dark5.1", x, act)
    x = csp_layer(sd, p + ".dark5.2", x, False, act)
    feats["dark5"] = x
    return feats


def pafpn(sd, x, act, p="backbone"):
    f = backbone(sd, x, act, p + ".backbone")
    x2, x1, x0 = f["dark3"], f["dark4"], f["dark5"]
    up = lambda t: F.interpolate(t, scale_factor=2, mode="nearest")
    fpn_out0 = base_conv(sd, p + ".lateral_conv0", x0, 1, act)
    f_out0 = csp_layer(sd, p + ".C3_p4", torch.cat([up(fpn_out0), x1], 1), False, act)
    fpn_out1 = base_conv(sd, p + ".reduce_conv1", f_out0, 1, act)
    pan_out2 = csp_layer(sd, p + ".C3_p3", torch.cat([up(fpn_out1), x2], 1), False, act)
    p_out1 = torch.cat([conv_any(sd, p + ".bu_conv2", pan_out2, 2, act), fpn_out1], 1)
    pan_out1 = csp_layer(sd, p + ".C3_n3", p_out1, False, act)
    p_out0 = torch.cat([conv_any(sd, p + ".bu_conv1", pan_out1, 2, act), fpn_out0], 1)
    pan_out0 = csp_layer(sd, p + ".C3_n4", p_out0, False, act)
    return pan_out2, pan_out1, pan_out0


def head(sd, feats, act, strides=(8, 16, 32), decode=True, sigmoid=True, p="head"):
    outs, hw = [], []
    for k, x in enumerate(feats):
        x = base_conv(sd, f"{p}.stems.{k}", x, 1, act)
        c = conv_any(sd, f"{p}.cls_convs.{k}.1", conv_any(sd, f"{p}.cls_convs.{k}.0", x, 1, act), 1, act)
        r = conv_any(sd, f"{p}.reg_convs.{k}.1", conv_any(sd, f"{p}.reg_convs.{k}.0", x, 1, act), 1, act)
        cls = F.conv2d(c, sd[f"{p}.cls_preds.{k}.weight"], sd[f"{p}.cls_preds.{k}.bias"])
        reg = F.conv2d(r, sd[f"{p}.reg_preds.{k}.weight"], sd[f"{p}.reg_preds.{k}.bias"])
        obj = F.conv2d(r, sd[f"{p}.obj_preds.{k}.weight"], sd[f"{p}.obj_preds.{k}.bias"])
        if sigmoid:
            obj, cls = obj.sigmoid(), cls.sigmoid()
        o = torch.cat([reg, obj, cls], 1)
        hw.append(tuple(o.shape[-2:]))
        outs.append(o.flatten(start_dim=2))
    out = torch.cat(outs, dim=2).permute(0, 2, 1)
    if decode:
        out = decode_outputs(out, hw, strides)
    return out, hw


def decode_outputs(outputs, hw, strides):
    """yolo_head.py:233-251."""
    grids, st = [], []
    for (h, w), s in zip(hw, strides):
        yv, xv = torch.meshgrid([torch.arange(h), torch.arange(w)], indexing="ij")
        g = torch.stack((xv, yv), 2).view(1, -1, 2)
        grids.append(g)
        st.append(torch.full((1, g.shape[1], 1), s))
    grids = torch.cat(grids, 1).to(outputs.dtype).to(outputs.device)
    st = torch.cat(st, 1).to(outputs.dtype).to(outputs.device)
    return torch.cat([(outputs[..., 0:2] + grids) * st, torch.exp(outputs[..., 2:4]) * st, outputs[..., 4:]], -1)


@torch.no_grad()
def forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, act: str = "silu", decode=True, sigmoid=True):
    """YoloxModule.forward (eval): x [B,3,H,W] fp32 raw 0..255 -> [B, A, 5+nc] fp32."""
    a = ACTS[act]
    sd = {k: v.float() for k, v in sd.items() if v.is_floating_point()}
    out, _ = head(sd, pafpn(sd, x.float(), a), a, decode=decode, sigmoid=sigmoid)
    return out


# ------------------------------------------------------------------------------------------------
# seeded, non-degenerate weights that do not depend on any model class (SURVEY.md 8c recipe):
# random conv weights, BN gamma~U(0.5,1.5), beta~N(0,0.2), cls/obj bias~N(-2,1.5), then BN running
# statistics calibrated layer by layer on U[0,255] images so that activations neither die nor blow up.
# ------------------------------------------------------------------------------------------------
def seeded_state_dict(template: Dict[str, torch.Tensor], seed: int, image_hw: Tuple[int, int] = (64, 64),
                      calib_batch: int = 8, act: str = "silu", big_boxes: bool = False,
                      calib_x: torch.Tensor = None, device=None) -> Dict[str, torch.Tensor]:
    """`device`: run the BN calibration forwards there (weight preparation only; the named configs at 640^2 take minutes
    on CPU). The returned state_dict is always on the CPU."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in template.items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros_like(v)
        elif k.endswith(".bn.weight"):
            sd[k] = torch.empty(v.shape).uniform_(0.5, 1.5, generator=g)
        elif k.endswith(".bn.bias"):
            sd[k] = torch.empty(v.shape).normal_(0, 0.2, generator=g)
        elif k.endswith("running_mean"):
            sd[k] = torch.zeros(v.shape)
        elif k.endswith("running_var"):
            sd[k] = torch.ones(v.shape)
        elif k.endswith(".weight"):
            fan_in = v[0].numel()
            sd[k] = torch.empty(v.shape).uniform_(-1, 1, generator=g) * (3.0 / fan_in) ** 0.5
        elif k.endswith(".bias"):
            if ".cls_preds." in k or ".obj_preds." in k:
                sd[k] = torch.empty(v.shape).normal_(-0.5, 1.5, generator=g)
            else:
                sd[k] = torch.empty(v.shape).normal_(0, 0.1, generator=g)
        else:
            sd[k] = v.clone()
    for k in list(sd):
        if ".reg_preds." in k and k.endswith(".weight"):
            sd[k] = sd[k] * (0.1 if big_boxes else 0.3)   # keep exp(wh) finite
    # calibration: run the eval graph with batch statistics, recording them as running stats
    from pixeltable_yolox_b200.synthetic import images

    x = torch.from_numpy(images(calib_batch, image_hw[0], image_hw[1], seed=seed + 100))
    if calib_x is not None:      # include the evaluation images: tiny maps give too few BN samples otherwise
        x = torch.cat([calib_x.float(), x], 0)
    if device is not None:
        sd = {k: v.to(device) for k, v in sd.items()}
        x = x.to(device)
    _calibrate(sd, x, ACTS[act])
    # keep the raw box regressions in a sane range (|v| <= 2.5) so that exp(wh) stays well conditioned
    with torch.no_grad():
        a = ACTS[act]
        raw, hw = head(sd, pafpn(sd, x, a), a, decode=False, sigmoid=False)
        off = 0
        for k, (h, w) in enumerate(hw):
            m = raw[:, off:off + h * w, :4].abs().max().item()
            off += h * w
            if m > 2.5:
                sd[f"head.reg_preds.{k}.weight"] = sd[f"head.reg_preds.{k}.weight"] * (2.5 / m)
                sd[f"head.reg_preds.{k}.bias"] = sd[f"head.reg_preds.{k}.bias"] * (2.5 / m)
    return {k: v.cpu() for k, v in sd.items()} if device is not None else sd


@torch.no_grad()
def _calibrate(sd, x, act):
    import contextlib

    orig = F.batch_norm

    def calibrating_bn(inp, rm, rv, weight=None, bias=None, training=False, momentum=0.0, eps=1e-5):
        mean = inp.mean(dim=(0, 2, 3))
        var = inp.var(dim=(0, 2, 3), unbiased=False)
        rm.copy_(mean)
        rv.copy_(torch.maximum(var, 0.05 * var.mean() + 1e-4))  # few samples at stride 32: floor the variance
        return orig(inp, rm, rv, weight, bias, False, 0.0, eps)

    F.batch_norm = calibrating_bn
    try:
        head(sd, pafpn(sd, x, act), act)
    finally:
        F.batch_norm = orig
