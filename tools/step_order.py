"""Times the drop-in step in both orders (inference half first / training half first)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synth_data as S
import pytorch_retinanet_b200 as P
from types import SimpleNamespace
dev = torch.device("cuda", 0)
cfg = S.CONFIGS[2]; n_img = 16
batch = S.make_batch(cfg, 0, n_img)
d_cls, d_box = batch["cls_preds"].to(dev), batch["bbox_preds"].to(dev)
targets = [{k: v.to(dev) for k, v in t.items()} for t in batch["targets"]]
gen = P.AnchorGenerator().to(dev)
fmaps = [torch.empty((n_img, 1, h, w), device=dev) for h, w in S.grid_sizes(cfg.padded_hw)]
images = SimpleNamespace(image_sizes=batch["im_szs"])
L = P.RetinaNetLosses(cfg.num_classes)
stub = SimpleNamespace(score_thres=0.05, nms_thres=0.5, detections_per_img=100)
def train(anchors):
    x, b = d_cls.detach().requires_grad_(True), d_box.detach().requires_grad_(True)
    out = L(targets, {"cls_preds": x, "bbox_preds": b}, anchors)
    (out["classification_loss"] + out["regression_loss"]).backward()
    return out, x.grad
def infer(anchors):
    return P.process_detections(stub, {"cls_preds": d_cls, "bbox_preds": d_box}, anchors, batch["im_szs"])
def a():
    an = gen(images, fmaps); d = infer(an); t = train(an); return d, t
def b():
    an = gen(images, fmaps); t = train(an); d = infer(an); return d, t
def c():
    an = gen(images, fmaps)
    h = P.process_detections_async(stub, {"cls_preds": d_cls, "bbox_preds": d_box}, an, batch["im_szs"])
    t = train(an); return h.detections(), t
for name, fn in (("infer->train", a), ("train->infer", b), ("infer(async)->train->collect", c)) * 2:
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200): fn()
    torch.cuda.synchronize(); print(f"{name:32s} {(time.perf_counter()-t0)/200*1e3:.3f} ms/step")
