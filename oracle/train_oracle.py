"""TEST INFRASTRUCTURE (checker only, never on the product path): numpy fp64 restatement of one training block of the
reference -- BaseConv in train mode, yolox/models/network_blocks.py:27-52: Conv2d(bias=False, pad=(k-1)//2) ->
BatchNorm2d (batch statistics, running statistics updated with `momentum` and the unbiased variance) -> SiLU -- and of
what autograd computes for it inside Trainer.train_one_iter (yolox/core/trainer.py:104-118: loss.backward()).

Pinned by tests/golden/trainblock.npz: forward output, running statistics and autograd gradients of the UNMODIFIED
reference's BaseConv on the seeded cases of tests/cases.py (tests/golden/make_golden.py: gen_trainblock;
tests/test_oracle_golden.py checks this file against them).
"""
from __future__ import annotations

import numpy as np


def _windows(xp, k, s, oh, ow):
    """[B, C, oh, ow, k, k] view of the padded input."""
    B, C = xp.shape[:2]
    st = xp.strides
    return np.lib.stride_tricks.as_strided(xp, (B, C, oh, ow, k, k), (st[0], st[1], st[2] * s, st[3] * s, st[2], st[3]), writeable=False)


def conv_forward(x, w, stride):
    """nn.Conv2d(bias=False, padding=(k-1)//2): x [B,Ci,H,W], w [Co,Ci,k,k] -> [B,Co,oh,ow]."""
    x = np.asarray(x, np.float64); w = np.asarray(w, np.float64)
    k = w.shape[2]; pad = (k - 1) // 2
    H, W = x.shape[2:]
    oh, ow = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    return np.einsum("bchwrs,ocrs->bohw", _windows(xp, k, stride, oh, ow), w, optimize=True)


def conv_backward(x, w, dy, stride):
    """(dx, dw) of conv_forward for the output gradient dy."""
    x = np.asarray(x, np.float64); w = np.asarray(w, np.float64); dy = np.asarray(dy, np.float64)
    k = w.shape[2]; pad = (k - 1) // 2
    B, Ci, H, W = x.shape
    oh, ow = dy.shape[2:]
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    dw = np.einsum("bchwrs,bohw->ocrs", _windows(xp, k, stride, oh, ow), dy, optimize=True)
    dxp = np.zeros_like(xp)
    for r in range(k):
        for q in range(k):
            dxp[:, :, r:r + stride * oh:stride, q:q + stride * ow:stride] += np.einsum("bohw,oc->bchw", dy, w[:, :, r, q], optimize=True)
    return dxp[:, :, pad:pad + H, pad:pad + W], dw


def _sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def bn_silu_forward(y, gamma, beta, eps, momentum, running_mean, running_var):
    """Train-mode BatchNorm2d + SiLU on y [B,C,H,W]. Returns (a, mean, invstd, new_running_mean, new_running_var)."""
    y = np.asarray(y, np.float64)
    n = y.shape[0] * y.shape[2] * y.shape[3]
    mean = y.mean((0, 2, 3)); var = y.var((0, 2, 3))
    invstd = 1.0 / np.sqrt(var + eps)
    z = (y - mean[None, :, None, None]) * invstd[None, :, None, None] * np.asarray(gamma, np.float64)[None, :, None, None] \
        + np.asarray(beta, np.float64)[None, :, None, None]
    unbiased = var * n / max(n - 1, 1)
    return (z * _sigmoid(z), mean, invstd, (1 - momentum) * np.asarray(running_mean, np.float64) + momentum * mean,
            (1 - momentum) * np.asarray(running_var, np.float64) + momentum * unbiased)


def bn_silu_backward(y, gamma, beta, mean, invstd, da):
    """(dy, dgamma, dbeta) for the gradient da of the block output."""
    y = np.asarray(y, np.float64); da = np.asarray(da, np.float64)
    g = np.asarray(gamma, np.float64)[None, :, None, None]
    xh = (y - mean[None, :, None, None]) * invstd[None, :, None, None]
    z = xh * g + np.asarray(beta, np.float64)[None, :, None, None]
    s = _sigmoid(z)
    dz = da * s * (1.0 + z * (1.0 - s))
    dbeta = dz.sum((0, 2, 3)); dgamma = (dz * xh).sum((0, 2, 3))
    n = y.shape[0] * y.shape[2] * y.shape[3]
    dy = g * invstd[None, :, None, None] * (dz - dbeta[None, :, None, None] / n - xh * dgamma[None, :, None, None] / n)
    return dy, dgamma, dbeta


def base_conv_train(x, w, gamma, beta, go, stride, eps=1e-3, momentum=0.03):
    """Forward + backward of one BaseConv block in train mode; dict with the arrays trainblock.npz holds."""
    co = w.shape[0]
    y = conv_forward(x, w, stride)
    a, mean, invstd, rm, rv = bn_silu_forward(y, gamma, beta, eps, momentum, np.zeros(co), np.ones(co))
    dy, dgamma, dbeta = bn_silu_backward(y, gamma, beta, mean, invstd, go)
    dx, dw = conv_backward(x, w, dy, stride)
    return dict(y=a, dx=dx, dw=dw, dgamma=dgamma, dbeta=dbeta, running_mean=rm, running_var=rv, conv_out=y, dconv_out=dy)
