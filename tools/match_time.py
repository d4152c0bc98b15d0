"""Times rn_match (codes + fg_count) on configs 2 (N=16, G<=100) and 5 (N=8, G=500)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synth_data as S
from pytorch_retinanet_b200.box_utils import PackedTargets, match_batch
dev = torch.device("cuda", 0)
for cid, n in ((2, 16), (5, 8)):
    cfg = S.CONFIGS[cid]
    anc = S.default_anchors(cfg.padded_hw).to(dev)
    g = torch.Generator().manual_seed(1)
    boxes, labels = [], []
    for i in range(n):
        gi = torch.Generator().manual_seed(1000 * cid + i)
        lo, hi = cfg.gt_range
        G = lo if lo == hi else int(torch.randint(lo, hi + 1, (1,), generator=gi))
        boxes.append(S._gt_boxes(gi, G, cfg.im_hw).to(dev)); labels.append(torch.randint(1, 81, (G,), generator=gi).to(dev))
    packed = PackedTargets(boxes, labels, dev)
    f = lambda: match_batch(anc, 0, packed, anc.shape[0], 0.5, 0.4, False, True)
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    pairs = anc.shape[0] * packed.total
    print(f"config {cid}: N={n} sumG={packed.total} match {ms*1000:.1f} us  {pairs/ms/1e6:.1f} G pairs/s")
