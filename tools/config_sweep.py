"""Times the path on the BASELINE.json configs that are not the bench workload (CUDA events, inputs in HBM):
config 5 (1024x1024, G=500: matcher-heavy training loss), config 4 (batch-256 inference post-processing, with and
without pre_nms_topk=1000), config 1 (the small CPU-reference case)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synth_data as S
import pytorch_retinanet_b200 as P
from pytorch_retinanet_b200.box_utils import PackedTargets
from pytorch_retinanet_b200.losses import fused_loss_forward
from pytorch_retinanet_b200.detections import postprocess_batch

dev = torch.device("cuda", 0)
PEAK = 6549.8


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def offsets(hw):
    o = [0]
    for h, w in S.grid_sizes(hw):
        o.append(o[-1] + 9 * h * w)
    return o


out = {}
# ---- config 5: 16 distinct images (the config's batch of 64 = 4x this; cost is linear in images) ----
cfg = S.CONFIGS[5]
n = 16
b = S.make_batch(cfg, 0, n)
x, bb, anc = b["cls_preds"].to(dev), b["bbox_preds"].to(dev), b["anchors"].to(dev)
tg = [{k: v.to(dev) for k, v in t.items()} for t in b["targets"]]
packed = PackedTargets([t["boxes"] for t in tg], [t["labels"] for t in tg], dev)
A, C = anc.shape[0], cfg.num_classes
ms = timed(lambda: fused_loss_forward(x, bb, anc, 0, packed, .25, 2., .1, .5, .4, float(n), True))
byts = n * (2 * 4 * A * C + 2 * 16 * A) + 16 * A
out["config5_loss_fwd_grad"] = {"images": n, "G": 500, "ms": ms, "images_per_s": n / ms * 1e3, "GBps": byts / ms / 1e6,
                                "frac_of_measured_peak": byts / ms / 1e6 / PEAK, "iou_pairs": A * packed.total}
ms = timed(lambda: postprocess_batch(x, bb, anc, 0, b["im_szs"], 0.05, 0.5, 100))
out["config5_postprocess"] = {"images": n, "ms": ms, "images_per_s": n / ms * 1e3}
del x, bb, b
# ---- config 4: inference, 32 distinct images tiled to batch 256 (16.5 GB of logits) ----
cfg = S.CONFIGS[4]
base = S.make_batch(cfg, 0, 32)
x = base["cls_preds"].to(dev).repeat(8, 1, 1)
bb = base["bbox_preds"].to(dev).repeat(8, 1, 1)
anc = base["anchors"].to(dev)
sz = base["im_szs"] * 8
A, C, n = anc.shape[0], cfg.num_classes, x.shape[0]
byts = n * (4 * A * C + 16 * A + 100 * 28) + 16 * A
for topk in (None, 1000):
    ms = timed(lambda: postprocess_batch(x, bb, anc, 0, sz, 0.05, 0.5, 100, pre_nms_topk=topk, level_offsets=offsets(cfg.padded_hw)), reps=5)
    out[f"config4_postprocess_topk_{topk}"] = {"images": n, "ms": ms, "images_per_s": n / ms * 1e3, "GBps": byts / ms / 1e6,
                                               "frac_of_measured_peak": byts / ms / 1e6 / PEAK}
del x, bb
# ---- config 1 ----
cfg = S.CONFIGS[1]
b = S.make_batch(cfg, 0, 1)
x, bb, anc = b["cls_preds"].to(dev), b["bbox_preds"].to(dev), b["anchors"].to(dev)
tg = [{k: v.to(dev) for k, v in t.items()} for t in b["targets"]]
packed = PackedTargets([t["boxes"] for t in tg], [t["labels"] for t in tg], dev)
ms1 = timed(lambda: fused_loss_forward(x, bb, anc, 0, packed, .25, 2., .1, .5, .4, 1.0, True), reps=100)
ms2 = timed(lambda: postprocess_batch(x, bb, anc, 0, b["im_szs"], 0.05, 0.5, 100), reps=100)
out["config1"] = {"loss_fwd_grad_ms": ms1, "postprocess_ms": ms2, "note": "1 image 512x512, 20 classes: launch/latency bound"}
print(json.dumps(out, indent=1))
