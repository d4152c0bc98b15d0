"""Times the row-N1 loss (match + loss_levels + finalize, fwd+grad) for the library selected by RN_LIB_SUFFIX."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synth_data as S
from pytorch_retinanet_b200.box_utils import PackedTargets
from pytorch_retinanet_b200.losses import fused_loss_forward_levels, fused_loss_forward
from pytorch_retinanet_b200 import _native
cfg = S.CONFIGS[2]
b = S.make_batch(cfg, 0, 16)
dev = torch.device("cuda", 0)
anc = b["anchors"].to(dev)
xs = [t.to(dev) for t in S.nac_to_levels(b["cls_preds"], cfg.padded_hw)]
bs = [t.to(dev) for t in S.nac_to_levels(b["bbox_preds"], cfg.padded_hw)]
tg = [{k: v.to(dev) for k, v in t.items()} for t in b["targets"]]
packed = PackedTargets([t["boxes"] for t in tg], [t["labels"] for t in tg], dev)
def run(): return fused_loss_forward_levels(xs, bs, anc, 0, packed, 80, .25, 2., .1, .5, .4, 16.0, True)
for i in range(5): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(30): run()
e1.record(); torch.cuda.synchronize()
print(os.path.basename(_native.lib_path()), "levels loss fwd+grad ms", e0.elapsed_time(e1) / 30)
def run2(): return fused_loss_forward_levels(xs, bs, anc, 0, packed, 80, .25, 2., .1, .5, .4, 16.0, False)
for i in range(5): run2()
torch.cuda.synchronize()
e0.record()
for i in range(30): run2()
e1.record(); torch.cuda.synchronize()
print(os.path.basename(_native.lib_path()), "levels loss fwd-only ms", e0.elapsed_time(e1) / 30)
