"""Host-side timing of one bench step, phase by phase (CPU enqueue time vs GPU time)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synth_data as S
import pytorch_retinanet_b200 as P
from pytorch_retinanet_b200.box_utils import PackedTargets, match_batch
from pytorch_retinanet_b200.losses import fused_loss_forward
from pytorch_retinanet_b200.detections import postprocess_batch
from types import SimpleNamespace

dev = torch.device("cuda", 0)
cfg = S.CONFIGS[2]
n_img = int(os.environ.get("NIMG", "16"))
batch = S.make_batch(cfg, 0, n_img)
d_cls, d_box = batch["cls_preds"].to(dev), batch["bbox_preds"].to(dev)
targets = [{k: v.to(dev) for k, v in t.items()} for t in batch["targets"]]
gen = P.AnchorGenerator().to(dev)
fmaps = [torch.empty((n_img, 1, h, w), device=dev) for h, w in S.grid_sizes(cfg.padded_hw)]
images = SimpleNamespace(image_sizes=batch["im_szs"])
L = P.RetinaNetLosses(cfg.num_classes)
stub = SimpleNamespace(score_thres=0.05, nms_thres=0.5, detections_per_img=100)
anchors = gen(images, fmaps)


def phase(name, fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    t1 = time.perf_counter()          # CPU enqueue time (async)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name:34s} cpu-enqueue {1e6*(t1-t0)/reps:8.1f} us   total {1e6*(t2-t0)/reps:8.1f} us")


def loss_fb():
    x, b = d_cls.detach().requires_grad_(True), d_box.detach().requires_grad_(True)
    out = L(targets, {"cls_preds": x, "bbox_preds": b}, anchors)
    (out["classification_loss"] + out["regression_loss"]).backward()

def loss_f():
    with torch.no_grad():
        L(targets, {"cls_preds": d_cls, "bbox_preds": d_box}, anchors)

packed = PackedTargets([t["boxes"] for t in targets], [t["labels"] for t in targets], dev)
phase("anchor_generator (cached)", lambda: gen(images, fmaps))
phase("PackedTargets", lambda: PackedTargets([t["boxes"] for t in targets], [t["labels"] for t in targets], dev))
phase("match_batch", lambda: match_batch(anchors[0], 0, packed, anchors[0].shape[0], 0.5, 0.4, False, True))
phase("fused_loss_forward(grad)", lambda: fused_loss_forward(d_cls, d_box, anchors[0], 0, packed, .25, 2., .1, .5, .4, float(n_img), True))
phase("fused_loss_forward(nograd)", lambda: fused_loss_forward(d_cls, d_box, anchors[0], 0, packed, .25, 2., .1, .5, .4, float(n_img), False))
phase("RetinaNetLosses fwd (no_grad)", loss_f)
phase("RetinaNetLosses fwd+bwd", loss_fb)
phase("postprocess_batch", lambda: postprocess_batch(d_cls, d_box, anchors[0], 0, batch["im_szs"], 0.05, 0.5, 100))
phase("process_detections", lambda: P.process_detections(stub, {"cls_preds": d_cls, "bbox_preds": d_box}, anchors, batch["im_szs"]))
def full():
    loss_fb()
    P.process_detections(stub, {"cls_preds": d_cls, "bbox_preds": d_box}, anchors, batch["im_szs"])
phase("full step", full)
