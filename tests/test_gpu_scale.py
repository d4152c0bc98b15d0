"""GPU tests (-m gpu) at the FULL batch sizes of BASELINE.json configs[2..4] — N = 128 training, N = 256 inference
(16.5 GB of logits), N = 64 dense-crowd images with 500 GT boxes each — on DISTINCT images (drawn on the device,
synth_data.make_batch_device), plus oracle comparisons on batches of distinct host-generated images.

At full size the whole batch cannot go through the (eager, per-image, per-class) oracle in test time, so every test
combines: (a) the CUDA oracle (the reference's op sequence on CUDA tensors, oracle/torch_oracle.py) on a spread sample of
images of the batch — matches / labels / detections bit-exact, losses 1e-5; (b) size-independent properties for ALL
images: the per-image results of the big batch equal, bit for bit, those of the same images processed in small
sub-batches; the batch loss is the mean of the per-image losses (losses.py:138-140); gradients scale exactly with the
(power-of-two) ratio of the batch sizes."""
import pytest
import torch

import synth_data as S
from helpers import rel_close, to_cuda_targets
from oracle import torch_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import pytorch_retinanet_b200 as pkg
    from pytorch_retinanet_b200 import _native
    _native.load()
    return pkg


def _free_gb():
    free, _ = torch.cuda.mem_get_info()
    return free / 2 ** 30


def _level_offsets(padded_hw):
    offs = [0]
    for h, w in S.grid_sizes(padded_hw):
        offs.append(offs[-1] + 9 * h * w)
    return offs


def _dets_equal(ob, os_, ol, counts, i, want, ctx):
    k = counts[i]
    assert k == want["boxes"].shape[0], (ctx, k, want["boxes"].shape)
    assert torch.equal(ol[i, :k], want["labels"]), ctx
    assert torch.equal(os_[i, :k], want["scores"]), ctx
    assert torch.equal(ob[i, :k], want["boxes"]), ctx


def _oracle_image_loss(b, i, C):
    """(cls_i / max(1,F_i), reg_i / max(1,F_i), F_i, matches) of image i by the CUDA oracle."""
    t = b["targets"][i]
    reg, cls, m, nfg = O.image_loss(b["anchors"], b["cls_preds"][i], b["bbox_preds"][i], t["labels"], t["boxes"], C, chunk=32768)
    return float(cls), float(reg), int(nfg), m


# --------------------------------------------------------------------------------------------- oracle on distinct images
def test_config5_eight_distinct_images_vs_cuda_oracle(P):
    """configs[4] shape (1024x1024, 500 GT boxes per image, A = 196,416, C = 80), 8 DISTINCT host-generated images in one
    batch: matches bit-exact, losses 1e-5 and gradients 2e-5 against the CUDA oracle, detections bit-exact."""
    from pytorch_retinanet_b200.detections import postprocess_batch
    cfg = S.CONFIGS[5]
    n = 8
    b = S.make_batch(cfg, 300, n, clustered=True)
    dev = torch.device("cuda")
    anc = b["anchors"].to(dev)
    tg = to_cuda_targets(b["targets"])
    x = b["cls_preds"].to(dev).requires_grad_(True)
    bb = b["bbox_preds"].to(dev).requires_grad_(True)
    L = P.RetinaNetLosses(cfg.num_classes)
    out = L(tg, {"cls_preds": x, "bbox_preds": bb}, [anc] * n)
    (out["classification_loss"] + out["regression_loss"]).backward()
    xo = b["cls_preds"].to(dev).requires_grad_(True)
    bo = b["bbox_preds"].to(dev).requires_grad_(True)
    want = O.batch_loss(tg, xo, bo, [anc] * n, cfg.num_classes, chunk=32768)
    (want["classification_loss"] + want["regression_loss"]).backward()
    assert rel_close(out["classification_loss"], want["classification_loss"].detach(), 1e-5)
    assert rel_close(out["regression_loss"], want["regression_loss"].detach(), 1e-5)
    assert rel_close(x.grad, xo.grad, 2e-5, 1e-12) and rel_close(bb.grad, bo.grad, 2e-5, 1e-9)
    for i in range(n):
        m = P.matcher(anc, tg[i]["boxes"])
        assert torch.equal(m, O.match(anc, tg[i]["boxes"], chunk=32768)), i
        assert int(L.last_per_image[i, 2]) == int((m >= 0).sum())
    ob, os_, ol, counts = postprocess_batch(x.detach(), bb.detach(), anc, 0, b["im_szs"], 0.05, 0.5, 100)
    ref = O.postprocess(x.detach(), bb.detach(), [anc] * n, b["im_szs"])
    for i in range(n):
        _dets_equal(ob, os_, ol, counts, i, ref[i], f"config 5 image {i}")


@pytest.mark.parametrize("topk", [None, 1000])
def test_config4_sixteen_distinct_images_vs_cuda_oracle(P, topk):
    """configs[3] (inference, 800x1333, score 0.05, NMS 0.5, 100 dets/img) on 16 DISTINCT host-generated images in one
    batch, top-k extension off and 1000 per level: detections bit-exact against the CUDA oracle (the reference's loop,
    plus the documented per-level filter when top-k is on)."""
    from pytorch_retinanet_b200.detections import postprocess_batch
    cfg = S.CONFIGS[4]
    n = 16
    b = S.make_batch(cfg, 40, n, clustered=True)
    dev = torch.device("cuda")
    anc, x, bb = b["anchors"].to(dev), b["cls_preds"].to(dev), b["bbox_preds"].to(dev)
    offs = _level_offsets(cfg.padded_hw)
    ob, os_, ol, counts = postprocess_batch(x, bb, anc, 0, b["im_szs"], 0.05, 0.5, 100, pre_nms_topk=topk,
                                            level_offsets=offs if topk else None)
    ref = O.postprocess(x, bb, [anc] * n, b["im_szs"], pre_nms_topk=topk, level_offsets=offs if topk else None)
    for i in range(n):
        _dets_equal(ob, os_, ol, counts, i, ref[i], f"config 4 image {i} topk {topk}")
    assert all(0 < k <= 100 for k in counts)


# ------------------------------------------------------------------------------------------------ full stated sizes
def test_config3_training_batch_128_distinct(P):
    if _free_gb() < 40:
        pytest.skip("needs ~35 GB of free device memory")
    cfg = S.CONFIGS[3]
    dev = torch.device("cuda")
    N, sub = 128, 16
    b = S.make_batch_device(cfg, 0, N, dev)
    anc, C = b["anchors"], cfg.num_classes
    L = P.RetinaNetLosses(C)
    xl, bl = b["cls_preds"].requires_grad_(True), b["bbox_preds"].requires_grad_(True)      # 8.26 GB of logits
    o = L(b["targets"], {"cls_preds": xl, "bbox_preds": bl}, [anc] * N)
    (o["classification_loss"] + o["regression_loss"]).backward()
    img = L.last_per_image.clone()
    assert rel_close(o["classification_loss"], img[:, 0].double().mean(), 1e-6)              # losses.py:138-140
    assert rel_close(o["regression_loss"], img[:, 1].double().mean(), 1e-6)
    scale = sub / float(N)                                                                    # power of two: exact
    for s0 in range(0, N, sub):                                                               # ALL images, in sub-batches
        xs = xl.detach()[s0:s0 + sub].clone().requires_grad_(True)
        bs = bl.detach()[s0:s0 + sub].clone().requires_grad_(True)
        os_ = L(b["targets"][s0:s0 + sub], {"cls_preds": xs, "bbox_preds": bs}, [anc] * sub)
        (os_["classification_loss"] + os_["regression_loss"]).backward()
        assert torch.equal(L.last_per_image, img[s0:s0 + sub]), s0
        assert torch.equal(xl.grad[s0:s0 + sub], xs.grad * scale) and torch.equal(bl.grad[s0:s0 + sub], bs.grad * scale), s0
    data = {"anchors": anc, "cls_preds": xl.detach(), "bbox_preds": bl.detach(), "targets": b["targets"]}
    for i in (0, 37, 64, 127):                                                                # CUDA oracle on a spread sample
        c, r, f, m = _oracle_image_loss(data, i, C)
        assert abs(float(img[i, 0]) - c) <= 1e-5 * abs(c) and abs(float(img[i, 1]) - r) <= 1e-5 * abs(r) + 1e-12, i
        assert int(img[i, 2]) == f and torch.equal(P.matcher(anc, b["targets"][i]["boxes"]), m), i


@pytest.mark.parametrize("topk", [None, 1000])
def test_config4_inference_batch_256_distinct(P, topk):
    if _free_gb() < 30:
        pytest.skip("needs ~25 GB of free device memory")
    from pytorch_retinanet_b200.detections import postprocess_batch
    cfg = S.CONFIGS[4]
    dev = torch.device("cuda")
    N, sub = 256, 16
    b = S.make_batch_device(cfg, 0, N, dev)                              # 16.5 GB of logits
    anc, x, bb = b["anchors"], b["cls_preds"], b["bbox_preds"]
    offs = _level_offsets(cfg.padded_hw)
    kw = dict(pre_nms_topk=topk, level_offsets=offs if topk else None)
    ob, os_, ol, c = postprocess_batch(x, bb, anc, 0, b["im_szs"], 0.05, 0.5, 100, **kw)
    assert all(0 < k <= 100 for k in c)
    for s0 in range(0, N, sub):                                          # ALL images, in sub-batches of 16
        sb, ss, sl, sc = postprocess_batch(x[s0:s0 + sub], bb[s0:s0 + sub], anc, 0, b["im_szs"][s0:s0 + sub], 0.05, 0.5, 100, **kw)
        assert sc == c[s0:s0 + sub], s0
        for j, k in enumerate(sc):
            i = s0 + j
            assert torch.equal(ob[i, :k], sb[j, :k]) and torch.equal(os_[i, :k], ss[j, :k]) and torch.equal(ol[i, :k], sl[j, :k]), i
    for i in (0, 85, 170, 255):                                          # CUDA oracle on a spread sample
        ref = O.postprocess(x[i:i + 1], bb[i:i + 1], [anc], b["im_szs"][:1], **kw)[0]
        _dets_equal(ob, os_, ol, c, i, ref, f"image {i} topk {topk}")
    ob2, os2, ol2, c2 = postprocess_batch(x, bb, anc, 0, b["im_szs"], 0.05, 0.5, 100, **kw)   # run-to-run deterministic
    assert c2 == c and torch.equal(os2, os_) and torch.equal(ol2, ol)


def test_config5_crowd_batch_64_distinct(P):
    """configs[4] AS STATED: 64 distinct images, 500 GT boxes each (6.3 G nominal IoU pairs)."""
    if _free_gb() < 20:
        pytest.skip("needs ~15 GB of free device memory")
    from pytorch_retinanet_b200.detections import postprocess_batch
    cfg = S.CONFIGS[5]
    dev = torch.device("cuda")
    N, sub = 64, 8
    b = S.make_batch_device(cfg, 0, N, dev)
    anc, C = b["anchors"], cfg.num_classes
    assert all(t["boxes"].shape[0] == 500 for t in b["targets"])
    L = P.RetinaNetLosses(C)
    xl, bl = b["cls_preds"].requires_grad_(True), b["bbox_preds"].requires_grad_(True)
    o = L(b["targets"], {"cls_preds": xl, "bbox_preds": bl}, [anc] * N)
    (o["classification_loss"] + o["regression_loss"]).backward()
    img = L.last_per_image.clone()
    assert rel_close(o["classification_loss"], img[:, 0].double().mean(), 1e-6)
    assert rel_close(o["regression_loss"], img[:, 1].double().mean(), 1e-6)
    ob, os_, ol, c = postprocess_batch(xl.detach(), bl.detach(), anc, 0, b["im_szs"], 0.05, 0.5, 100)
    scale = sub / float(N)
    for s0 in range(0, N, sub):
        xs = xl.detach()[s0:s0 + sub].clone().requires_grad_(True)
        bs = bl.detach()[s0:s0 + sub].clone().requires_grad_(True)
        o8 = L(b["targets"][s0:s0 + sub], {"cls_preds": xs, "bbox_preds": bs}, [anc] * sub)
        (o8["classification_loss"] + o8["regression_loss"]).backward()
        assert torch.equal(L.last_per_image, img[s0:s0 + sub]), s0
        assert torch.equal(xl.grad[s0:s0 + sub], xs.grad * scale) and torch.equal(bl.grad[s0:s0 + sub], bs.grad * scale), s0
        sb, ss, sl, sc = postprocess_batch(xs.detach(), bs.detach(), anc, 0, b["im_szs"][s0:s0 + sub], 0.05, 0.5, 100)
        assert sc == c[s0:s0 + sub]
        for j, k in enumerate(sc):
            assert torch.equal(ob[s0 + j, :k], sb[j, :k]) and torch.equal(ol[s0 + j, :k], sl[j, :k]), s0 + j
    data = {"anchors": anc, "cls_preds": xl.detach(), "bbox_preds": bl.detach(), "targets": b["targets"]}
    for i in (0, 21, 42, 63):
        cl, r, f, m = _oracle_image_loss(data, i, C)
        assert abs(float(img[i, 0]) - cl) <= 1e-5 * abs(cl) and abs(float(img[i, 1]) - r) <= 1e-5 * abs(r), i
        assert int(img[i, 2]) == f and torch.equal(P.matcher(anc, b["targets"][i]["boxes"]), m), i
        ref = O.postprocess(xl.detach()[i:i + 1], bl.detach()[i:i + 1], [anc], b["im_szs"][:1])[0]
        _dets_equal(ob, os_, ol, c, i, ref, f"config 5 image {i}")
