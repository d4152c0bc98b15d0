"""Runs post-processing once with the instrumented library (RN_LAZY_TIMING=1) and prints per-image phase cycles."""
import os, sys
os.environ["RN_LAZY_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synth_data as S
from pytorch_retinanet_b200.detections import postprocess_batch
cfg = S.CONFIGS[2]
b = S.make_batch(cfg, 0, 16)
dev = torch.device("cuda", 0)
x, bb, anc = b["cls_preds"].to(dev), b["bbox_preds"].to(dev), b["anchors"].to(dev)
for i in range(2):
    postprocess_batch(x, bb, anc, 0, b["im_szs"], 0.05, 0.5, 100)
    torch.cuda.synchronize()
    print("----")
