"""GPU tests (-m gpu) at the FULL batch sizes of BASELINE.json configs[2] (N=128 training) and configs[3]
(N=256 inference, 16.5 GB of logits), through size-independent properties: a batch made of R copies of a small
batch must reproduce the small batch's per-image results exactly (the small batch itself is pinned against the
oracle and the reference's golden vectors in test_gpu_parity.py), the batch loss is the mean of the per-image
losses, and gradients scale by the (power-of-two) ratio of the batch sizes bit-exactly."""
import pytest
import torch

import synth_data as S
from helpers import rel_close, to_cuda_targets

pytestmark = pytest.mark.gpu

BASE = 4          # distinct images


@pytest.fixture(scope="module")
def P():
    import pytorch_retinanet_b200 as pkg
    from pytorch_retinanet_b200 import _native
    _native.load()
    return pkg


@pytest.fixture(scope="module")
def base():
    cfg = S.CONFIGS[2]
    b = S.make_batch(cfg, 200, BASE, clustered=True)
    dev = torch.device("cuda")
    return dict(cfg=cfg, anc=b["anchors"].to(dev), x=b["cls_preds"].to(dev), bb=b["bbox_preds"].to(dev),
                tg=to_cuda_targets(b["targets"]), im_szs=b["im_szs"])


def _free_gb():
    free, _ = torch.cuda.mem_get_info()
    return free / 2 ** 30


def test_config3_scale_training_batch_128(P, base):
    if _free_gb() < 40:
        pytest.skip("needs ~35 GB of free device memory")
    cfg, anc = base["cfg"], base["anc"]
    L = P.RetinaNetLosses(cfg.num_classes)
    xs, bs = base["x"].clone().requires_grad_(True), base["bb"].clone().requires_grad_(True)
    o4 = L(base["tg"], {"cls_preds": xs, "bbox_preds": bs}, [anc] * BASE)
    (o4["classification_loss"] + o4["regression_loss"]).backward()
    img4 = L.last_per_image.clone()
    R = 128 // BASE
    xl = base["x"].repeat(R, 1, 1).requires_grad_(True)                  # 8.26 GB
    bl = base["bb"].repeat(R, 1, 1).requires_grad_(True)
    o128 = L(base["tg"] * R, {"cls_preds": xl, "bbox_preds": bl}, [anc] * 128)
    (o128["classification_loss"] + o128["regression_loss"]).backward()
    img128 = L.last_per_image
    assert torch.equal(img128, img4.repeat(R, 1))                                   # per-image values: bit-identical
    assert rel_close(o128["classification_loss"], img4[:, 0].double().mean(), 1e-6)  # losses.py:138-140
    assert rel_close(o128["regression_loss"], img4[:, 1].double().mean(), 1e-6)
    assert rel_close(o128["classification_loss"], o4["classification_loss"].detach(), 1e-6)
    scale = BASE / 128.0                                                             # power of two: exact
    for i in (0, 5, 64, 127):
        assert torch.equal(xl.grad[i], xs.grad[i % BASE] * scale), i
        assert torch.equal(bl.grad[i], bs.grad[i % BASE] * scale), i
    # foreground counts feed the multi-GPU statistics vector: [cls, reg, sum F, N]
    assert int(img128[:, 2].sum()) == R * int(img4[:, 2].sum())


@pytest.mark.parametrize("topk", [None, 1000])
def test_config4_scale_inference_batch_256(P, base, topk):
    if _free_gb() < 30:
        pytest.skip("needs ~25 GB of free device memory")
    from pytorch_retinanet_b200.detections import postprocess_batch
    cfg, anc = base["cfg"], base["anc"]
    offs = [0]
    for h, w in S.grid_sizes(cfg.padded_hw):
        offs.append(offs[-1] + 9 * h * w)
    kw = dict(pre_nms_topk=topk, level_offsets=offs if topk else None)
    ob4, os4, ol4, c4 = postprocess_batch(base["x"], base["bb"], anc, 0, base["im_szs"], 0.05, 0.5, 100, **kw)
    R = 256 // BASE
    xl = base["x"].repeat(R, 1, 1)                                       # 16.5 GB of logits
    bl = base["bb"].repeat(R, 1, 1)
    ob, os_, ol, c = postprocess_batch(xl, bl, anc, 0, base["im_szs"] * R, 0.05, 0.5, 100, **kw)
    assert c == c4 * R
    for i in range(256):
        k = c[i]
        j = i % BASE
        assert torch.equal(ob[i, :k], ob4[j, :k]) and torch.equal(os_[i, :k], os4[j, :k]) and torch.equal(ol[i, :k], ol4[j, :k]), i
    assert all(0 < k <= 100 for k in c)
    # idempotent and run-to-run deterministic at this size
    ob2, os2, ol2, c2 = postprocess_batch(xl, bl, anc, 0, base["im_szs"] * R, 0.05, 0.5, 100, **kw)
    assert c2 == c
    for i in (0, 17, 255):
        assert torch.equal(os2[i, :c[i]], os_[i, :c[i]]) and torch.equal(ol2[i, :c[i]], ol[i, :c[i]])
