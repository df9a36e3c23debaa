"""-m gpu: the kernels either side of the detection path -- device letterbox (bit-exact with cv2), evaluator result rows
(bit-exact with the reference's convert_to_coco_format), the training-branch head rows (forward + backward) and the fused
SGD + EMA step."""
import copy
import math
from pathlib import Path

import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu

import pixeltable_yolox_b200 as yx  # noqa: E402
from oracle import preproc_oracle as pre  # noqa: E402
from pixeltable_yolox_b200 import ops  # noqa: E402
from pixeltable_yolox_b200 import synthetic as syn  # noqa: E402

GOLDEN = Path(__file__).resolve().parent / "golden"


def test_letterbox_on_device_is_bit_exact(cuda):
    rng = np.random.default_rng(4)
    sizes = [(1280, 1280), (480, 640), (1280, 1000), (640, 640), (3, 5), (37, 911), (701, 13), (416, 416), (832, 832)]
    sizes += [(int(rng.integers(4, 1400)), int(rng.integers(4, 1400))) for _ in range(15)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    for S in ((640, 640), (416, 416), (320, 512)):
        want = np.stack([pre.letterbox(im, S, np.uint8) for im in imgs])
        got = ops.letterbox_u8(imgs, S, cuda, torch.uint8)
        assert got.shape == (len(imgs), 3, S[0], S[1]) and got.dtype == torch.uint8
        bad = [sizes[i] for i in range(len(imgs)) if not np.array_equal(got[i].cpu().numpy(), want[i])]
        assert not bad, bad
        gotf = ops.letterbox_u8(imgs[:4], S, cuda, torch.float32)
        assert np.array_equal(gotf.cpu().numpy(), want[:4].astype(np.float32))
    gray = [rng.integers(0, 256, (200, 300), dtype=np.uint8)]
    assert np.array_equal(ops.letterbox_u8(gray, (416, 416), cuda)[0, 0].cpu().numpy(), pre.letterbox(gray[0], (416, 416), np.uint8))


def test_processor_device_letterbox_equals_host_path(cuda):
    from PIL import Image

    img = Image.fromarray(cases.config1_image())
    proc = yx.YoloxProcessor("yolox_s")
    host = proc([img, img])
    proc.device = cuda
    dev = proc([img, img])
    assert dev.is_cuda and dev.dtype == torch.float32 and torch.equal(dev.cpu(), host)
    proc.dtype = torch.uint8
    assert torch.equal(proc([img]).cpu().float(), host[:1])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.letterbox_u8([np.zeros((4, 4, 3), np.uint8)], (32, 32), torch.device("cpu"))


def test_coco_rows_match_reference_evaluator(cuda):
    from pixeltable_yolox_b200.coco import convert_to_coco_format

    g = np.load(GOLDEN / "coco.npz")
    c = cases.COCO_CASE
    pred, conf, nms = cases.post_case(c["post_case"])
    dets, _, cnt = ops.postprocess_device(torch.from_numpy(pred).to(cuda), 80, conf, nms, yx.boxes.NMS_VARIANTS["auto_cpu"])
    hs, ws, ids, class_ids = cases.coco_case_meta(pred.shape[0])
    data_list, wise = convert_to_coco_format(dets, cnt, (hs, ws), ids, c["img_size"], class_ids, return_outputs=True)
    assert len(data_list) == int(g["n"])
    assert [d["image_id"] for d in data_list] == g["image_id"].tolist()
    assert [d["category_id"] for d in data_list] == g["category_id"].tolist()
    np.testing.assert_array_equal(np.array([d["bbox"] for d in data_list]), g["bbox"])
    np.testing.assert_array_equal(np.array([d["score"] for d in data_list]), g["score"])
    assert all(d["segmentation"] == [] for d in data_list) and sorted(wise) == g["wise_ids"].tolist()
    # max_det smaller than the number kept: rows beyond it are dropped, the compaction stays dense
    few, _, cnt2 = ops.postprocess_device(torch.from_numpy(pred).to(cuda), 80, conf, nms, yx.boxes.NMS_VARIANTS["auto_cpu"], max_det=2)
    rows = convert_to_coco_format(few, cnt2, (hs, ws), ids, c["img_size"], class_ids)
    assert len(rows) == int(cnt2.clamp(max=2).sum())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("use_l1", [False, True])
def test_training_branch_rows_forward_and_backward(cuda, dtype, use_l1):
    """_TrainRows (one kernel each way) against the torch ops of the reference's training branch
    (cat -> get_output_and_grid -> origin_preds), values and gradients."""
    from pixeltable_yolox_b200.yolo_head import _TrainRows

    head = yx.YoloxHead(80)
    g = torch.Generator().manual_seed(3)
    B, h, w, stride = 3, 13, 21, 16
    mk = lambda c: (torch.randn(B, c, h, w, generator=g) * 0.7).to(dtype).to(cuda)
    reg, obj, cls = mk(4), mk(1), mk(80)
    leaves = [t.clone().requires_grad_(True) for t in (reg, obj, cls)]
    out, origin = _TrainRows.apply(*leaves, stride, use_l1)
    ref_leaves = [t.clone().float().requires_grad_(True) for t in (reg, obj, cls)]
    cat = torch.cat(ref_leaves, 1)
    want, grid = head.get_output_and_grid(cat, 0, stride, cat.type())
    assert out.dtype == torch.float32 and out.shape == want.shape
    torch.testing.assert_close(out, want, rtol=2e-6, atol=1e-6)
    go = torch.randn(out.shape, generator=g).to(cuda)
    loss, ref_loss = (out * go).sum(), (want * go).sum()
    if use_l1:
        want_or = ref_leaves[0].view(B, 1, 4, h, w).permute(0, 1, 3, 4, 2).reshape(B, -1, 4)
        assert torch.equal(origin, want_or)
        gor = torch.randn(origin.shape, generator=g).to(cuda)
        loss, ref_loss = loss + (origin * gor).sum(), ref_loss + (want_or * gor).sum()
    else:
        assert origin.numel() == 0
    loss.backward(); ref_loss.backward()
    tol = dict(rtol=1e-5, atol=1e-6) if dtype == torch.float32 else dict(rtol=1e-2, atol=1e-2)
    for a, b in zip(leaves, ref_leaves):
        assert a.grad.dtype == dtype and a.grad.shape == b.grad.shape
        torch.testing.assert_close(a.grad.float(), b.grad, **tol)


def test_fused_sgd_ema_matches_torch_sgd_and_reference_ema(cuda):
    """FusedSgdEma (one launch) against torch.optim.SGD with the reference's parameter groups and ModelEMA.update's formula."""
    from pixeltable_yolox_b200.optim import FusedSgdEma, reference_param_groups

    torch.manual_seed(0)
    cfg = yx.YoloxConfig("opt", depth=0.33, width=0.25)
    model = cfg.get_model().to(cuda).train()
    twin = copy.deepcopy(model)
    pg0, pg1, pg2 = reference_param_groups(twin)
    opt = torch.optim.SGD(pg0, lr=0.01, momentum=0.9, nesterov=True)
    opt.add_param_group({"params": pg1, "weight_decay": 5e-4})
    opt.add_param_group({"params": pg2})
    ema = copy.deepcopy(twin).eval()
    fused = FusedSgdEma(model, lr=0.01, momentum=0.9, weight_decay=5e-4, nesterov=True, ema=True, ema_decay=0.9998)
    assert sum(p.numel() for p in pg0 + pg1 + pg2) == sum(p.numel() for p in twin.parameters())
    g = torch.Generator().manual_seed(1)
    for step in range(1, 4):
        for p, q in zip(model.parameters(), twin.parameters()):
            gr = torch.randn(p.shape, generator=g).to(cuda) * 0.1
            p.grad = gr.clone() if p.grad is None else p.grad.copy_(gr)
            q.grad = gr.clone()
        with torch.no_grad():                          # BN running statistics move between steps as in training
            for (k, b), (_, b2) in zip(model.named_buffers(), twin.named_buffers()):
                if b.dtype.is_floating_point:
                    b.add_(0.01 * step); b2.add_(0.01 * step)
        fused.step()
        opt.step()
        d = 0.9998 * (1 - math.exp(-step / 2000))
        with torch.no_grad():
            msd = twin.state_dict()
            for k, v in ema.state_dict().items():
                if v.dtype.is_floating_point:
                    v *= d
                    v += (1.0 - d) * msd[k].detach()
        for (k, a), (_, b) in zip(model.state_dict().items(), twin.state_dict().items()):
            if a.dtype.is_floating_point:
                torch.testing.assert_close(a, b, rtol=2e-6, atol=1e-7, msg=lambda m: f"step {step} {k}: {m}")
        for (k, a), (_, b) in zip(fused.ema.state_dict().items(), ema.state_dict().items()):
            if a.dtype.is_floating_point:
                torch.testing.assert_close(a, b, rtol=2e-6, atol=1e-7, msg=lambda m: f"step {step} ema {k}: {m}")
    assert fused.updates == 3


def test_training_step_as_cuda_graph_matches_eager(cuda):
    """The whole training step (forward, SimOTA, losses, backward, fused SGD + EMA) captured once and replayed gives the
    loss, parameters and EMA the eager step gives FROM THE SAME STATE (two models drift apart on random weights: the SimOTA
    costs are near-tied, so 1e-7 differences from cuDNN's atomics flip assignments). Nothing in the step synchronises with
    the host; the learning rate / EMA ramp reach the captured optimizer launch through FusedSgdEma.hyper."""
    from pixeltable_yolox_b200.optim import FusedSgdEma

    torch.manual_seed(0)
    old = (torch.backends.cudnn.deterministic, torch.backends.cudnn.allow_tf32)
    torch.backends.cudnn.deterministic, torch.backends.cudnn.allow_tf32 = True, False    # same wgrad sums in both runs
    cfg = yx.YoloxConfig("graph", depth=0.33, width=0.25)
    m = cfg.get_model().to(cuda).train()
    x = torch.from_numpy(syn.images(2, 128, 128, seed=3)).to(cuda)
    lab = torch.from_numpy(syn.labels(2, max_gt=8, seed=5, size=128.0, counts=[3, 5])).to(cuda)
    opt = FusedSgdEma(m, lr=0.01, ema=True)

    def eager(lr):
        out = m(x, lab)
        opt.zero_grad()
        out["total_loss"].backward()
        opt.step(lr)
        return out["total_loss"].detach().clone()

    def snapshot():
        return ([v.clone() for v in m.state_dict().values()], [b.clone() for b in opt.bufs],
                [v.clone() for v in opt.ema.state_dict().values()], opt.updates)

    def restore(s):
        with torch.no_grad():
            for dst, src in zip(m.state_dict().values(), s[0]): dst.copy_(src)
            for dst, src in zip(opt.bufs, s[1]): dst.copy_(src)
            for dst, src in zip(opt.ema.state_dict().values(), s[2]): dst.copy_(src)
        opt.updates = s[3]

    for _ in range(2):                                   # momentum buffers, pointer table, workspaces
        eager(0.01)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = m(x, lab)
        loss_g = out["total_loss"]
        opt.zero_grad()
        loss_g.backward()
        opt.step_captured()
    for lr in (0.02, 0.005):
        before = snapshot()
        want_loss = eager(lr)
        want = snapshot()
        restore(before)
        opt.set_hyper(lr)
        g.replay()
        got = snapshot()
        torch.testing.assert_close(loss_g.detach(), want_loss, rtol=1e-5, atol=1e-6)
        assert got[3] == want[3]
        for part in range(3):
            for i, (a, b) in enumerate(zip(got[part], want[part])):
                if a.dtype.is_floating_point:
                    torch.testing.assert_close(a, b, rtol=2e-3, atol=2e-4, msg=lambda t: f"part {part} tensor {i}: {t}")
    torch.backends.cudnn.deterministic, torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("act", ["silu", "relu", "lrelu"])
@pytest.mark.parametrize("shape", [(3, 48, 20, 20), (2, 32, 37, 29), (8, 16, 80, 80), (2, 1280, 5, 7)])
@pytest.mark.parametrize("channels_last", [False, True])
def test_fused_train_bn_act_matches_torch(cuda, dtype, act, shape, channels_last):
    """Training-mode BatchNorm2d + activation (one fused pass each way) against nn.BatchNorm2d + the activation module in
    torch: output, running statistics, and the gradients w.r.t. input, gamma and beta."""
    from pixeltable_yolox_b200.network_blocks import _FusedBnAct, get_activation

    g = torch.Generator().manual_seed(4)
    N, Cc, H, W = shape
    x0 = (torch.randn(shape, generator=g) * 1.5 + 0.3).to(dtype).to(cuda)
    if channels_last:
        x0 = x0.contiguous(memory_format=torch.channels_last)
    bn = torch.nn.BatchNorm2d(Cc, eps=1e-3, momentum=0.03).to(cuda).train()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(Cc, generator=g) + 0.5); bn.bias.copy_(torch.randn(Cc, generator=g) * 0.2)
        bn.running_mean.copy_(torch.randn(Cc, generator=g) * 0.1); bn.running_var.copy_(torch.rand(Cc, generator=g) + 0.5)
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    gam, bet = bn.weight.detach().clone().requires_grad_(True), bn.bias.detach().clone().requires_grad_(True)
    x = x0.clone().requires_grad_(True)
    nbt = torch.full((), 5, dtype=torch.int64, device=cuda)
    y = _FusedBnAct.apply(x, gam, bet, rm, rv, bn.eps, bn.momentum, act, nbt)
    go = torch.randn(shape, generator=g).to(dtype).to(cuda)
    y.backward(go)
    # reference in fp32 on the same (already rounded) input
    xr = x0.float().clone().requires_grad_(True)
    yr = get_activation(act, inplace=False)(bn(xr))
    yr.backward(go.float())
    tol = dict(rtol=2e-5, atol=2e-5) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    assert y.dtype == dtype and y.stride() == x0.stride() and x.grad.stride() == x0.stride()
    assert int(nbt) == 6                       # nn.BatchNorm2d.num_batches_tracked, incremented by the forward kernel
    torch.testing.assert_close(y.float(), yr, **tol)
    torch.testing.assert_close(rm, bn.running_mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rv, bn.running_var, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(x.grad.float(), xr.grad, **tol)
    gt = dict(rtol=1e-4, atol=1e-3) if dtype == torch.float32 else dict(rtol=2e-2, atol=5e-2 * (N * H * W) ** 0.5 / 10)
    torch.testing.assert_close(gam.grad, bn.weight.grad, **gt)
    torch.testing.assert_close(bet.grad, bn.bias.grad, **gt)


def test_fused_allreduce_sgd_ema_kernel_on_one_rank_equals_the_plain_optimizer_launch(cuda):
    """yx_allreduce_sgd_ema_step (gradient all-reduce over NVLink peer memory + SGD + EMA in one kernel) with a one-rank group:
    the barriers and the reduce-scatter degenerate, the update must equal yx_sgd_ema_step bit for bit. The multi-rank check
    (against ncclAllReduce, parameters identical across ranks) is tools/gpu_allreduce_sgd_check.py under torchrun."""
    import os

    import torch.distributed as dist

    from pixeltable_yolox_b200 import train_conv
    from pixeltable_yolox_b200.optim import FusedSgdEma

    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        try:
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=cuda)
        except Exception as e:                      # noqa: BLE001
            pytest.skip(f"cannot create a one-rank NCCL group here: {e!r}")
        created = True
    try:
        torch.manual_seed(0)
        m = yx.YoloxConfig("ar", depth=0.33, width=0.25).get_model().to(cuda).train()
        try:
            opt = FusedSgdEma(m, lr=0.01, ema=True, direct_grads=True, peer_group=dist.group.WORLD)
        except Exception as e:                      # noqa: BLE001
            pytest.skip(f"symmetric memory unavailable: {e!r}")
        g = torch.Generator(device=cuda).manual_seed(5)
        for _ in range(2):
            opt.flat_grad.copy_(torch.randn(opt.flat_grad.shape, generator=g, device=cuda) * 0.1)
            opt.step(0.01)

        def snapshot():
            return [v.clone() for v in m.state_dict().values()] + [b.clone() for b in opt.bufs] + [v.clone() for v in opt.ema.state_dict().values()]

        def restore(s):
            with torch.no_grad():
                for dst, src in zip(list(m.state_dict().values()) + opt.bufs + list(opt.ema.state_dict().values()), s):
                    dst.copy_(src)

        for it in range(3):
            opt.flat_grad.copy_(torch.randn(opt.flat_grad.shape, generator=g, device=cuda) * 0.1)
            grads, before, upd = opt.flat_grad.clone(), snapshot(), opt.updates
            opt.set_hyper(0.02); opt.step_captured()
            want = snapshot()
            restore(before); opt.updates = upd; opt.flat_grad.copy_(grads)
            opt.set_hyper(0.02); opt.step_allreduce_captured()
            torch.cuda.synchronize()
            for a, b in zip(snapshot(), want):
                assert torch.equal(a, b)
            assert torch.equal(opt.flat_grad, grads)            # one rank: the reduced slice is the gradient itself
    finally:
        if created:
            dist.destroy_process_group()


def test_captured_training_step_with_branches_replays_bit_identically(cuda):
    """The captured step with everything switched on (tcgen05 convs, batched weight packing, direct gradient accumulation,
    weight gradients on their own stream, network branches on side streams): with lr = 0 the parameters never change, so
    every replay must give the same loss and the same gradients, bit for bit -- a race between the graph's branches, or
    memory handed to two streams at once, shows up as a replay that differs (tools/gpu_train_stress.py: 300 replays of the
    full-size step)."""
    from pixeltable_yolox_b200 import train_conv
    from pixeltable_yolox_b200.optim import FusedSgdEma

    torch.manual_seed(0)
    m = yx.YoloxConfig("stress", depth=0.33, width=0.25).get_model().to(cuda).train().to(memory_format=torch.channels_last)
    x = torch.from_numpy(syn.images(4, 160, 160, seed=3)).to(cuda).contiguous(memory_format=torch.channels_last)
    lab = torch.from_numpy(syn.labels(4, max_gt=16, seed=5, size=160.0, counts=[3, 0, 16, 7])).to(cuda)
    opt = FusedSgdEma(m, lr=0.0, direct_grads=True)
    train_conv.attach_packer(m, torch.bfloat16)
    try:
        def eager():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = m(x, lab)
            opt.zero_grad()
            out["total_loss"].backward()
            opt.step(0.0)
            return out["total_loss"].detach().clone()

        side = torch.cuda.Stream(cuda)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                l_eager = eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = m(x, lab)
            loss = out["total_loss"]
            opt.zero_grad()
            loss.backward()
            opt.join()
            grads = opt.flat_grad.clone()
            opt.step_captured()
        opt.set_hyper(0.0); g.replay(); torch.cuda.synchronize()
        l0, g0 = loss.detach().clone(), grads.clone()
        assert torch.equal(l0, l_eager) and float(g0.abs().max()) > 0
        for _ in range(40):
            opt.set_hyper(0.0); g.replay(); torch.cuda.synchronize()
            assert torch.equal(loss.detach(), l0)
            assert torch.equal(grads, g0)
    finally:
        opt.close()
        m.__dict__.pop("_yx_train_packer", None)
