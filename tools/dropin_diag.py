"""Diagnostic: drop-in step time (CUDA events) in a fresh process, then after fused HotPathGraphs exist, then host cProfile."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synth_data as S
import pytorch_retinanet_b200 as P
from pytorch_retinanet_b200.graphs import HotPathGraph
from types import SimpleNamespace

dev = torch.device("cuda", 0)
cfg = S.CONFIGS[2]
n_img = 16
batch = S.make_batch(cfg, 0, n_img)
d_cls, d_box = batch["cls_preds"].to(dev), batch["bbox_preds"].to(dev)
targets = [{k: v.to(dev) for k, v in t.items()} for t in batch["targets"]]
gen = P.AnchorGenerator().to(dev)
fmaps = [torch.empty((n_img, 1, h, w), device=dev) for h, w in S.grid_sizes(cfg.padded_hw)]
images = SimpleNamespace(image_sizes=batch["im_szs"])
L = P.RetinaNetLosses(cfg.num_classes)
stub = SimpleNamespace(score_thres=0.05, nms_thres=0.5, detections_per_img=100)

def step():
    anchors = gen(images, fmaps)
    dets = P.process_detections(stub, {"cls_preds": d_cls, "bbox_preds": d_box}, anchors, batch["im_szs"])
    x, b = d_cls.detach().requires_grad_(True), d_box.detach().requires_grad_(True)
    out = L(targets, {"cls_preds": x, "bbox_preds": b}, anchors)
    (out["classification_loss"] + out["regression_loss"]).backward()
    return dets

def timed(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) / n * 1e3

print("fresh process: dropin step ms (events, wall)", timed(step), "mem GB", torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9)
for fused in (False, True):
    g = HotPathGraph(cfg.num_classes, d_cls, d_box, gen(images, fmaps)[0], batch["im_szs"], fused=fused)
    print("graph fused=%s: step ms" % fused, timed(lambda: g.step(targets).detections()))
    print("  after that graph exists: dropin step ms", timed(step), "mem GB", torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9)
    del g
    torch.cuda.synchronize()
    print("  after deleting it: dropin step ms", timed(step))
pr = cProfile.Profile(); pr.enable()
for _ in range(100): step()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
