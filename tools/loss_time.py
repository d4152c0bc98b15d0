"""Times rn_train_loss (match + loss + finalize), fwd+grad and fwd-only, config 2, for the library selected by RN_LIB_SUFFIX."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synth_data as S
from pytorch_retinanet_b200 import _native
from pytorch_retinanet_b200.box_utils import PackedTargets
from pytorch_retinanet_b200.losses import fused_loss_forward
dev = torch.device("cuda", 0)
cfg = S.CONFIGS[2]; n = 16
b = S.make_batch(cfg, 0, n)
anc = b["anchors"].to(dev); x = b["cls_preds"].to(dev); bb = b["bbox_preds"].to(dev)
tg = [{k: v.to(dev) for k, v in t.items()} for t in b["targets"]]
packed = PackedTargets([t["boxes"] for t in tg], [t["labels"] for t in tg], dev)
REPS = int(os.environ.get("RN_REPS", "50"))
def t(fn, reps=REPS):
    for _ in range(min(5, reps)): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
g = t(lambda: fused_loss_forward(x, bb, anc, 0, packed, 0.25, 2.0, 0.1, 0.5, 0.4, float(n), True))
f = t(lambda: fused_loss_forward(x, bb, anc, 0, packed, 0.25, 2.0, 0.1, 0.5, 0.4, float(n), False))
print(os.path.basename(_native.lib_path()), f"train loss fwd+grad {g*1000:.1f} us   fwd-only {f*1000:.1f} us")
