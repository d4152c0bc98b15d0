"""Times the row-N1 post-processing (rn_postprocess_levels) and, for comparison, rn_postprocess on the same data."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synth_data as S
from pytorch_retinanet_b200.detections import postprocess_levels_async, postprocess_batch
from pytorch_retinanet_b200 import _native
cfg = S.CONFIGS[2]
b = S.make_batch(cfg, 0, 16)
dev = torch.device("cuda", 0)
anc = b["anchors"].to(dev)
xs = [t.to(dev) for t in S.nac_to_levels(b["cls_preds"], cfg.padded_hw)]
bs = [t.to(dev) for t in S.nac_to_levels(b["bbox_preds"], cfg.padded_hw)]
x, bb = b["cls_preds"].to(dev), b["bbox_preds"].to(dev)
def t(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
lv = t(lambda: postprocess_levels_async(xs, bs, 80, anc, 0, b["im_szs"], 0.05, 0.5, 100).result())
nac = t(lambda: postprocess_batch(x, bb, anc, 0, b["im_szs"], 0.05, 0.5, 100))
print(os.path.basename(_native.lib_path()), f"postprocess levels {lv*1000:.1f} us   [N,A,C] {nac*1000:.1f} us")
