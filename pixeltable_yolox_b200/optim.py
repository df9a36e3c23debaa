"""The reference's optimizer step and EMA update (yolox/config.py:307-333: SGD momentum 0.9, nesterov, weight decay on conv /
linear weights only; yolox/utils/ema.py:15-58: ModelEMA with decay * (1 - exp(-updates / 2000))) as ONE kernel launch over
every parameter and buffer of the model (csrc/yx_train.cu: sgd_ema_kernel), instead of ~3 foreach launches per group plus
two elementwise kernels per state tensor for the EMA."""
from __future__ import annotations

import copy
import math
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops


def reference_param_groups(model: nn.Module):
    """pg0 (BN weights, no decay), pg1 (conv / linear weights, decay), pg2 (biases, no decay): config.py:316-324."""
    pg0, pg1, pg2 = [], [], []
    for k, v in model.named_modules():
        if hasattr(v, "bias") and isinstance(v.bias, nn.Parameter):
            pg2.append(v.bias)
        if isinstance(v, nn.BatchNorm2d) or "bn" in k:
            pg0.append(v.weight)
        elif hasattr(v, "weight") and isinstance(v.weight, nn.Parameter):
            pg1.append(v.weight)
    return pg0, pg1, pg2


class FusedSgdEma:
    """step(): p <- SGD(p, grad) for every parameter, then ema <- d * ema + (1 - d) * value for every floating-point
    state tensor (parameters and BN running statistics), all in one launch. `ema` is a deep copy of the model in eval
    mode (ModelEMA.ema); pass ema=False for the optimizer alone."""

    CHUNK = 16384

    def __init__(self, model: nn.Module, lr: float, momentum: float = 0.9, weight_decay: float = 5e-4, nesterov: bool = True,
                 ema: bool = True, ema_decay: float = 0.9998, updates: int = 0, direct_grads: bool = False, peer_group=None):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("FusedSgdEma needs the model on a CUDA device (the B200 path has no CPU fallback)")
        self.model, self.lr, self.momentum, self.nesterov = model, float(lr), float(momentum), bool(nesterov)
        self.ema_decay, self.updates = float(ema_decay), int(updates)
        self.ema = None
        if ema:
            self.ema = copy.deepcopy(model).eval()
            for p in self.ema.parameters():
                p.requires_grad_(False)
        pg0, pg1, pg2 = reference_param_groups(model)
        wd = {id(p): 0.0 for p in pg0 + pg2}
        wd.update({id(p): float(weight_decay) for p in pg1})
        self.params = [p for p in model.parameters() if p.requires_grad]
        missing = [p for p in self.params if id(p) not in wd]
        if missing:
            raise ValueError(f"{len(missing)} trainable parameters belong to no reference parameter group")
        self.bufs = [torch.zeros_like(p) for p in self.params]      # same strides as the parameter (NCHW or channels_last)
        # direct_grads: every .grad is a view into ONE flat fp32 buffer (zero_grad = one memset, a multi-GPU step all-reduces
        # the buffer itself), and the conv / BatchNorm backward kernels add their weight gradients into it in place
        # (train_conv.direct_target: opt-in per parameter). Not for modules wrapped in DistributedDataParallel.
        self.flat_grad = None
        if direct_grads:
            from . import train_conv

            n_flat = sum((p.numel() + 3) // 4 * 4 for p in self.params)      # every slot starts on a 16-byte boundary (vector loads)
            self._peer = None
            if peer_group is not None:
                # the flat buffer as SYMMETRIC memory: every rank of the node maps every other rank's buffer over NVLink, and
                # step_allreduce_captured() runs the all-reduce and the update as one kernel over those mappings
                import torch.distributed as dist
                import torch.distributed._symmetric_memory as symm

                self.flat_grad = symm.empty(n_flat, dtype=torch.float32, device=dev)
                self.flat_grad.zero_()
                flags = symm.empty(64, dtype=torch.int32, device=dev)
                flags.zero_()
                hg = symm.rendezvous(self.flat_grad, group=peer_group)
                hf = symm.rendezvous(flags, group=peer_group)
                self._peer = dict(grad=[int(v) for v in hg.buffer_ptrs], flag=[int(v) for v in hf.buffer_ptrs], rank=int(hg.rank),
                                  world=int(hg.world_size), flags=flags, state=torch.zeros(4, dtype=torch.int32, device=dev),
                                  handles=(hg, hf))
                assert self._peer["grad"][self._peer["rank"]] == self.flat_grad.data_ptr()
                torch.cuda.synchronize(dev)
                dist.barrier(group=peer_group)               # every rank's flags are zero before anyone signals
            else:
                self.flat_grad = torch.zeros(n_flat, dtype=torch.float32, device=dev)
            off = 0
            for p in self.params:
                assert p.is_contiguous() or p.is_contiguous(memory_format=torch.channels_last), "parameter storage is not dense"
                p.grad = torch.as_strided(self.flat_grad, p.shape, p.stride(), off)
                p._yx_direct_grad = True              # the backward kernels may add into p.grad (train_conv.direct_target)
                off += (p.numel() + 3) // 4 * 4
        self._wd = [wd[id(p)] for p in self.params]
        self._steps = 0
        self._table = None
        self._dev = dev
        # CUDA-graph mode: the captured launch reads {lr, d, 1 - d} from this device buffer; set_hyper() refreshes it from a
        # pinned host mirror before every replay (no host sync)
        self.hyper = torch.zeros(3, dtype=torch.float32, device=dev)
        self._hyper_host = torch.zeros(3, dtype=torch.float32).pin_memory()

    def _build_table(self):
        """(re)built when a gradient tensor was reallocated (zero_grad(set_to_none=True) gives new storage each step only
        if the allocator hands out a different block; pointers are compared on every step, which is 460 integer reads)."""
        rows, ema_sd = [], (dict(self.ema.state_dict()) if self.ema is not None else {})
        names = {id(v): k for k, v in self.model.state_dict(keep_vars=True).items()}
        seen = set()
        for p, buf, w in zip(self.params, self.bufs, self._wd):
            e = ema_sd.get(names.get(id(p)))
            # the kernel walks every tensor as flat memory: dense storage, and one element order for the parameter, its
            # gradient, its momentum buffer and its EMA copy (all NCHW, or all channels_last)
            assert p.dtype == torch.float32 and p.grad is not None
            assert p.is_contiguous() or p.is_contiguous(memory_format=torch.channels_last), "parameter storage is not dense"
            order = [st for st, n in zip(p.stride(), p.shape) if n > 1]          # strides of size-1 dims carry no order
            for other in (p.grad, buf, e):
                assert other is None or [st for st, n in zip(other.stride(), other.shape) if n > 1] == order, \
                    "parameter / gradient / momentum / EMA memory formats differ"
            rows.append((p.data_ptr(), p.grad.data_ptr(), buf.data_ptr(), 0 if e is None else e.data_ptr(), p.numel(),
                         int(np.float32(w).view(np.int32)) & 0xffffffff))
            seen.add(names.get(id(p)))
        if self.ema is not None:            # floating-point buffers (BN running mean / var): EMA only
            for k, v in self.model.state_dict().items():
                if k in seen or not v.dtype.is_floating_point:
                    continue
                rows.append((v.data_ptr(), 0, 0, ema_sd[k].data_ptr(), v.numel(), 0))
        table = np.array(rows, dtype=np.int64).reshape(-1, 6)
        chunks = [(t, e) for t, r in enumerate(rows) for e in range(0, r[4], self.CHUNK)]
        self._table = torch.from_numpy(table).to(self._dev)
        self._chunks = torch.from_numpy(np.array(chunks, dtype=np.int32).reshape(-1, 2)).to(self._dev)
        self._grad_ptrs = [p.grad.data_ptr() for p in self.params]

    def close(self):
        """Give the parameters back to plain autograd accumulation (direct_grads): their .grad stays a view of the flat buffer
        until it is replaced, but the backward kernels no longer write into it."""
        if self.flat_grad is not None:
            self.join()
            for p in self.params:
                if hasattr(p, "_yx_direct_grad"):
                    del p._yx_direct_grad

    def join(self):
        if self.flat_grad is not None:
            from . import train_conv

            train_conv.join_wgrad(self._dev)       # side-stream weight gradients land before the update reads them

    def zero_grad(self):
        """Keeps the gradient storage (the pointer table stays valid): grads are zeroed in place."""
        if self.flat_grad is not None:
            self.flat_grad.zero_()
            return
        grads = [p.grad for p in self.params if p.grad is not None]
        if grads:
            torch._foreach_zero_(grads)          # a handful of multi-tensor launches instead of one per parameter

    def set_hyper(self, lr: Optional[float] = None):
        """Advance the EMA ramp by one update and publish {lr, d, 1 - d} to the device buffer a captured step() reads.
        Call once before each graph replay (the replay itself runs the optimizer + EMA launch)."""
        if lr is not None:
            self.lr = float(lr)
        d = 0.0
        if self.ema is not None:
            self.updates += 1
            d = self.ema_decay * (1 - math.exp(-self.updates / 2000))
        self._hyper_host[0] = self.lr
        self._hyper_host[1] = float(np.float32(d))
        self._hyper_host[2] = float(np.float32(1.0 - d))
        self.hyper.copy_(self._hyper_host, non_blocking=True)
        self._steps += 1

    @torch.no_grad()
    def step_captured(self):
        """The launch to record inside torch.cuda.graph(): hyper-parameters come from `self.hyper` at run time."""
        if self._table is None:
            raise RuntimeError("FusedSgdEma.step_captured: run at least one eager step() first (pointer table, momentum init)")
        self.join()
        ops.sgd_ema_step(self._table, self._chunks, self.CHUNK, self.lr, self.momentum, self.nesterov, False, 0.0, self.hyper)

    @torch.no_grad()
    def step_allreduce_captured(self):
        """Gradient all-reduce (mean over the ranks of `peer_group`) + optimizer step + EMA update as ONE launch over NVLink
        peer memory; hyper-parameters from `self.hyper` like step_captured(). Every rank must call it the same number of times."""
        if self._peer is None:
            raise RuntimeError("FusedSgdEma.step_allreduce_captured needs direct_grads=True and a peer_group")
        if self._table is None:
            raise RuntimeError("FusedSgdEma.step_allreduce_captured: run at least one eager step() first (pointer table, momentum init)")
        self.join()
        P = self._peer
        ops.allreduce_sgd_ema_step(self._table, self._chunks, self.CHUNK, self.momentum, self.nesterov, False, self.hyper,
                                   P["grad"], P["flag"], self.flat_grad.numel(), P["rank"], P["world"], P["state"])

    @torch.no_grad()
    def step(self, lr: Optional[float] = None):
        if lr is not None:
            self.lr = float(lr)
        self.join()
        if self._table is None or any(p.grad is None or p.grad.data_ptr() != q for p, q in zip(self.params, self._grad_ptrs)):
            if any(p.grad is None for p in self.params):
                raise RuntimeError("FusedSgdEma.step: a trainable parameter has no gradient")
            self._build_table()
        d = 0.0
        if self.ema is not None:
            self.updates += 1
            d = self.ema_decay * (1 - math.exp(-self.updates / 2000))
        ops.sgd_ema_step(self._table, self._chunks, self.CHUNK, self.lr, self.momentum, self.nesterov, self._steps == 0, d)
        self._steps += 1
