"""Times post-processing (CUDA events) for the library selected by RN_LIB_SUFFIX."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synth_data as S
from pytorch_retinanet_b200.detections import postprocess_batch
from pytorch_retinanet_b200 import _native
cfg = S.CONFIGS[2]
b = S.make_batch(cfg, 0, 16)
dev = torch.device("cuda", 0)
x, bb, anc = b["cls_preds"].to(dev), b["bbox_preds"].to(dev), b["anchors"].to(dev)
for algo in ("auto", "general"):
    for i in range(5):
        postprocess_batch(x, bb, anc, 0, b["im_szs"], 0.05, 0.5, 100, algo=algo)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30):
        postprocess_batch(x, bb, anc, 0, b["im_szs"], 0.05, 0.5, 100, algo=algo)
    e1.record(); torch.cuda.synchronize()
    print(os.path.basename(_native.lib_path()), algo, "postprocess ms", e0.elapsed_time(e1) / 30)
