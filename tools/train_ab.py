"""A/B timing of rn_train_loss: single-launch kernel vs the rn_match + rn_loss sequence (config 2 shape,
random N(-7,1.3) logits — the timing does not depend on the clustered structure)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synth_data as S
from pytorch_retinanet_b200 import _native
from pytorch_retinanet_b200.box_utils import PackedTargets
from pytorch_retinanet_b200.losses import fused_loss_forward
lib = _native.load()
dev = torch.device("cuda", 0)
cid, n = int(os.environ.get("CID", 2)), int(os.environ.get("NIMG", 16))
cfg = S.CONFIGS[cid]
anc = S.default_anchors(cfg.padded_hw).to(dev)
A, C = anc.shape[0], cfg.num_classes
boxes, labels = [], []
for i in range(n):
    gi = torch.Generator().manual_seed(1000 * cid + i)
    lo, hi = cfg.gt_range
    G = lo if lo == hi else int(torch.randint(lo, hi + 1, (1,), generator=gi))
    boxes.append(S._gt_boxes(gi, G, cfg.im_hw).to(dev)); labels.append(torch.randint(1, C + 1, (G,), generator=gi).to(dev))
packed = PackedTargets(boxes, labels, dev)
x = torch.randn((n, A, C), device=dev).mul_(1.3).add_(-7.0)
b = torch.randn((n, A, 4), device=dev).mul_(0.1)
bytes_fb = n * (2 * 4 * A * C + 2 * 16 * A) + 16 * A
def t(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for fused in (0, 1, 2, 0, 1, 2):
    lib.rn_train_loss_set_fused(fused)
    g = t(lambda: fused_loss_forward(x, b, anc, 0, packed, 0.25, 2.0, 0.1, 0.5, 0.4, float(n), True))
    f = t(lambda: fused_loss_forward(x, b, anc, 0, packed, 0.25, 2.0, 0.1, 0.5, 0.4, float(n), False))
    print(f"fused={fused}: fwd+grad {g*1000:.1f} us ({bytes_fb/g/1e6:.0f} GB/s)   fwd {f*1000:.1f} us")
lib.rn_train_loss_set_fused(1)
