import os, sys, torch
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
import pixeltable_yolox_b200 as yx
from pixeltable_yolox_b200 import synthetic as syn
dev = torch.device("cuda", 0)
def cos(a, b): return float(torch.dot(a, b) / (a.norm() * b.norm()))
for depth in (2, 3, 5):
  for seed in range(2):
    torch.manual_seed(seed)
    m = yx.YoloxConfig.get_named_config("yolox_nano").get_model().to(dev).train()
    bb = m.backbone.backbone
    x = torch.from_numpy(syn.images(2, 128, 128, seed=seed + 3)).to(dev)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    res = {}
    for name, flag, amp in (("fp32", "0", False), ("ours", "1", True), ("torch16", "0", True)):
        os.environ["YX_TRAIN_CONV"] = flag
        m.load_state_dict(sd); m.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            t = bb.stem._train_forward(x)
            for k, blk in enumerate((bb.dark2, bb.dark3, bb.dark4, bb.dark5)):
                if k + 2 > depth: break
                for sub in blk: t = sub._train_forward(t)
        loss = t.float().square().mean(); loss.backward()
        res[name] = (float(loss), torch.cat([p.grad.flatten() for p in m.parameters() if p.grad is not None]).clone())
    print(f"through dark{depth} seed {seed}: cos vs fp32 ours {cos(res['ours'][1], res['fp32'][1]):.4f} torch16 {cos(res['torch16'][1], res['fp32'][1]):.4f}; loss {res['fp32'][0]:.5f} {res['ours'][0]:.5f} {res['torch16'][0]:.5f}")
