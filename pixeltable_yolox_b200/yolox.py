"""Yolox / YoloxModule with the reference's interface (yolox/models/yolox.py:21-131)."""
from __future__ import annotations

import os
import urllib.request
from pathlib import Path
from typing import Iterable, Optional, Union

import torch
import torch.nn as nn

from .config import YoloxConfig
from .processor import Detections, YoloxProcessor
from .yolo_head import YoloxHead
from .yolo_pafpn import YoloPafpn

HOME = Path(os.environ.get("YOLOX_HOME", str(Path.home() / ".cache" / "yolox")))


class Yolox:
    module: "YoloxModule"
    processor: YoloxProcessor

    def __init__(self, module: "YoloxModule", processor: YoloxProcessor):
        self.module = module
        self.processor = processor

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, config: Optional[YoloxConfig] = None,
                        device: str = "cpu") -> "Yolox":
        module = YoloxModule.from_pretrained(pretrained_model_name_or_path, config, device)
        processor = YoloxProcessor(config or pretrained_model_name_or_path)
        return cls(module, processor)

    def __call__(self, inputs: Iterable, threshold: float = 0.5):
        if isinstance(inputs, torch.Tensor):
            return self.module(inputs)  # deprecated call pattern of the reference (yolox.py:42-44)
        from PIL import Image

        images = [im if isinstance(im, Image.Image) else Image.open(im) for im in inputs]
        tensor = self.processor(images)
        dev = next(self.module.parameters()).device
        if self.module.training or not self.module.head.decode_in_inference:
            output = self.module(tensor.to(dev))
            return self.processor.postprocess(images, output, threshold=threshold)
        # eval: forward + decode + score filter + NMS as ONE captured graph (detect()); identical rows to
        # module(tensor) followed by processor.postprocess (tests/test_gpu_named_configs.py), without the
        # [B, A, 5+nc] prediction tensor ever leaving the engine
        dets, _, counts = self.module.detect(tensor.to(dev), conf_thre=threshold, nms_thre=self.processor.config.nmsthre,
                                             nms_variant=self.processor.nms_variant)
        return self.processor.format_detections(images, dets, counts)


class YoloxModule(nn.Module):
    """eval forward = one captured plan of sm_100a kernels (cached per input shape);
    training forward = PyTorch autograd ops + the batched SimOTA kernel."""

    def __init__(self, backbone: Optional[YoloPafpn] = None, head: Optional[YoloxHead] = None):
        super().__init__()
        self.backbone = backbone if backbone is not None else YoloPafpn()
        self.head = head if head is not None else YoloxHead(80)
        self._engines = {}
        self._fingerprint = None
        self.micro_batch = 64       # images per pass (see engine.InferenceEngine)
        self.use_cuda_graph = True

    # ------------------------------------------------------------------ engine cache
    # An engine holds raw pointers (native plan, packed weights, activation buffers): it belongs to exactly one module
    # object and is never copied or pickled with it. Contract: weights are BN-folded and packed when an engine is built;
    # every eval call compares a fingerprint of all parameters and buffers (storage pointer + in-place version counter,
    # BN eps, head strides / class count) and rebuilds when it changed, so load_state_dict on a sub-module,
    # initialize_biases(), param.data.copy_() or .half() are all picked up. Not detected: writes that bypass the
    # version counter (raw CUDA kernels, .data_ptr() writes from outside torch); call invalidate_engine() after those.
    def invalidate_engine(self):
        for e in self._engines.values():
            e.close()
        self._engines = {}
        self._fingerprint = None

    def __deepcopy__(self, memo):
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        import copy

        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k == "_engines" else (None if k == "_fingerprint" else copy.deepcopy(v, memo))
        return new

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_engines"] = {}
        state["_fingerprint"] = None
        return state

    def _weights_fingerprint(self):
        fp = [(t.data_ptr(), t._version) for t in self.parameters()]
        fp += [(t.data_ptr(), t._version) for t in self.buffers()]
        fp += [m.eps for m in self.modules() if isinstance(m, nn.BatchNorm2d)]
        fp += [tuple(self.head.strides), self.head.num_classes]
        return fp

    def train(self, mode: bool = True):
        self.invalidate_engine()
        return super().train(mode)

    def _apply(self, fn, *args, **kwargs):
        self.invalidate_engine()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.invalidate_engine()
        return super().load_state_dict(*args, **kwargs)

    def engine_for(self, x: torch.Tensor, post: Optional[dict] = None, slot: int = 0):
        """The cached InferenceEngine for this input signature. `slot` selects an independent engine
        (own staging/activation buffers) so that two batches can be in flight on two streams."""
        from .engine import InferenceEngine

        dev = next(self.parameters()).device
        fp = self._weights_fingerprint()
        if fp != self._fingerprint:
            if self._engines:
                self.invalidate_engine()
            self._fingerprint = fp
        key = (tuple(x.shape), x.dtype, dev, self.head.decode_in_inference, self.micro_batch, self.use_cuda_graph,
               tuple(sorted((k, v) for k, v in post.items() if v is not None)) if post else None, slot)
        eng = self._engines.get(key)
        if eng is None:
            eng = InferenceEngine(self, x.shape[0], x.shape[2], x.shape[3], x.dtype, dev,
                                  micro_batch=self.micro_batch, use_graph=self.use_cuda_graph, post=post)
            self._engines[key] = eng
        return eng

    # ------------------------------------------------------------------ forward
    def forward(self, x, targets=None):
        if self.training:
            assert targets is not None
            from .train_conv import packed_weights

            with packed_weights(self):         # one launch packs every conv weight when a WeightPacker is attached
                fpn_outs = self.backbone._train_forward(x)
                loss, iou_loss, conf_loss, cls_loss, l1_loss, num_fg = self.head(fpn_outs, targets, x)
            return {"total_loss": loss, "iou_loss": iou_loss, "l1_loss": l1_loss, "conf_loss": conf_loss,
                    "cls_loss": cls_loss, "num_fg": num_fg}
        p = next(self.parameters())
        if not p.is_cuda:
            raise RuntimeError("YoloxModule.forward(eval) on CPU: the B200 path runs hand-written sm_100a CUDA "
                               "only and has no CPU fallback; move the module to a CUDA device")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"expected an image batch [B,3,H,W], got {tuple(x.shape)}")
        if x.dtype not in (torch.float32, torch.uint8):
            x = x.float()  # pixel values 0..255 are exact in fp32/bf16/fp16
        out = self.engine_for(x).forward(x).clone()
        return out if self.head.output_dtype is None else out.to(self.head.output_dtype)

    @torch.no_grad()
    def detect(self, x: torch.Tensor, conf_thre: float = 0.5, nms_thre: float = 0.65, nms_variant: str = "auto",
               max_det: Optional[int] = None):
        """Fused forward + decode + score filter + NMS in one CUDA graph (the BASELINE metric path).
        Returns (dets [B,max_det,7], det_idx [B,max_det], det_count [B]) device tensors."""
        from .boxes import NMS_VARIANTS

        post = dict(conf_thre=float(conf_thre), nms_thre=float(nms_thre), nms_variant=NMS_VARIANTS[nms_variant],
                    max_det=max_det)
        eng = self.engine_for(x, post)
        eng.forward(x)
        return eng.dets, eng.det_idx, eng.det_count

    def visualize(self, x, targets, save_prefix="assign_vis_"):
        raise NotImplementedError("visualize_assign_result is a debugging tool outside the hot path")

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, config: Optional[YoloxConfig] = None,
                        device: str = "cpu") -> "YoloxModule":
        path = str(pretrained_model_name_or_path)
        if os.path.isfile(path):
            if config is None:
                raise ValueError("config must be provided when loading model from a file")
        else:
            config = YoloxConfig.get_named_config(path)
            if config is None:
                raise ValueError(f"Unknown model: {pretrained_model_name_or_path}")
            path = cls.__cached_pretrained_weights(path)
        model = config.get_model().to(device)
        model.eval()
        model.head.training = False
        model.training = False
        weights = torch.load(path, map_location=torch.device(device))
        model.load_state_dict(weights["model"])
        return model

    @classmethod
    def __cached_pretrained_weights(cls, model_id: str) -> str:
        weights_dir = HOME / "weights"
        weights_dir.mkdir(exist_ok=True, parents=True)
        weights_file = weights_dir / f"{model_id}.pth"
        if not weights_file.exists():
            url = f"https://github.com/Megvii-BaseDetection/YOLOX/releases/download/0.1.1rc0/{model_id}.pth"
            urllib.request.urlretrieve(url, f"{weights_file}.tmp")
            os.rename(f"{weights_file}.tmp", weights_file)
        return str(weights_file)
