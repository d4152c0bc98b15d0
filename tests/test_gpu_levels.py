"""GPU parity of the per-level NCHW entry points: the loss and the
post-processing computed directly on the head's raw conv outputs must equal the [N,A,C] path
(bit-exact detections, identical per-image loss values) and the oracle, and the gradients must be the
NCHW re-layout of the [N,A,C] gradients."""
import types
from types import SimpleNamespace

import pytest
import torch

import synth_data as S
from helpers import rel_close, to_cuda_targets
from oracle import torch_oracle as O

pytestmark = pytest.mark.gpu


def nac_to_levels(x, grid_sizes, na):
    """[N, A, K] -> list of [N, na*K, H, W] (inverse of layers.py:189-195)."""
    N, _, K = x.shape
    out, off = [], 0
    for h, w in grid_sizes:
        n = h * w * na
        out.append(x[:, off:off + n].reshape(N, h, w, na, K).permute(0, 3, 4, 1, 2).reshape(N, na * K, h, w).contiguous())
        off += n
    return out


def levels_to_nac(levels, na, K):
    """The reference's own re-layout: view -> permute(0,3,4,1,2) -> contiguous -> view, then cat."""
    outs = []
    for x in levels:
        N, _, H, W = x.shape
        outs.append(x.view(N, -1, K, H, W).permute(0, 3, 4, 1, 2).contiguous().view(N, -1, K))
    return torch.cat(outs, 1)


@pytest.mark.parametrize("hw,C,n_img", [((512, 512), 20, 2), ((800, 1344), 80, 2), ((200, 328), 5, 3)])
def test_levels_loss_and_detections_match_nac_path(hw, C, n_img):
    import pytorch_retinanet_b200 as P
    gen = torch.Generator().manual_seed(17)
    gs = S.grid_sizes(hw)                                   # (200,328): H*W not a multiple of 4 on most levels
    anc = S.default_anchors(hw)
    A = anc.shape[0]
    dev = torch.device("cuda")
    cls = (torch.randn((n_img, A, C), generator=gen) * 1.6 - 5.0)
    box = torch.randn((n_img, A, 4), generator=gen) * 0.2
    targets = []
    for i in range(n_img):
        g = S._gt_boxes(gen, 3 + 9 * i, hw)
        targets.append({"boxes": g, "labels": torch.randint(1, C + 1, (g.shape[0],), generator=gen)})
    fgcols = O.match(anc, targets[0]["boxes"])
    fg = torch.nonzero(fgcols >= 0).squeeze(1)
    cls[0, fg, targets[0]["labels"][fgcols[fg]] - 1] += 6.0   # some confident positives
    tg = to_cuda_targets(targets)
    anc_g = anc.to(dev)
    L = P.RetinaNetLosses(C)
    # reference layout
    x, b = cls.to(dev).requires_grad_(True), box.to(dev).requires_grad_(True)
    out = L(tg, {"cls_preds": x, "bbox_preds": b}, [anc_g] * n_img)
    (out["classification_loss"] + 2.0 * out["regression_loss"]).backward()
    per_image = L.last_per_image.clone()
    # raw per-level layout
    xl = [t.to(dev).requires_grad_(True) for t in nac_to_levels(cls, gs, 9)]
    bl = [t.to(dev).requires_grad_(True) for t in nac_to_levels(box, gs, 9)]
    assert torch.equal(levels_to_nac([t.detach() for t in xl], 9, C).cpu(), cls)          # the test's own re-layout is the reference's
    outl = L(tg, {"cls_levels": xl, "bbox_levels": bl}, [anc_g] * n_img)
    (outl["classification_loss"] + 2.0 * outl["regression_loss"]).backward()
    assert rel_close(outl["classification_loss"], out["classification_loss"].detach(), 1e-6)
    assert rel_close(outl["regression_loss"], out["regression_loss"].detach(), 1e-6)
    assert rel_close(L.last_per_image, per_image, 1e-6, 1e-9)
    assert torch.equal(L.last_per_image[:, 2], per_image[:, 2])
    gx = levels_to_nac([t.grad for t in xl], 9, C)
    gb = levels_to_nac([t.grad for t in bl], 9, 4)
    assert rel_close(gx, x.grad, 1e-6, 1e-12) and rel_close(gb, b.grad, 1e-6, 1e-12)
    # against the oracle on the reference layout
    xo, bo = cls.clone().requires_grad_(True), box.clone().requires_grad_(True)
    want = O.batch_loss(targets, xo, bo, [anc] * n_img, C)
    (want["classification_loss"] + 2.0 * want["regression_loss"]).backward()
    assert rel_close(outl["classification_loss"], want["classification_loss"].detach(), 1e-5)
    assert rel_close(outl["regression_loss"], want["regression_loss"].detach(), 1e-5)
    assert rel_close(gx, xo.grad, 2e-5, 1e-11)
    # post-processing: bit-identical detections to the [N,A,C] path, both algorithms
    sz = [hw] * n_img
    stub = SimpleNamespace(score_thres=0.05, nms_thres=0.5, detections_per_img=100, num_classes=C)
    ref = P.process_detections(stub, {"cls_preds": cls.to(dev), "bbox_preds": box.to(dev)}, [anc_g] * n_img, sz)
    outs = {"cls_levels": [t.detach() for t in xl], "bbox_levels": [t.detach() for t in bl]}
    got = P.process_detections(stub, outs, [anc_g] * n_img, sz)
    assert outs == {}
    from pytorch_retinanet_b200.detections import postprocess_levels_async
    gen_res = postprocess_levels_async([t.detach() for t in xl], [t.detach() for t in bl], C, anc_g, 0, sz, 0.05, 0.5, 100,
                                       algo="general").detections()
    for r, g, gg in zip(ref, got, gen_res):
        for k in ("boxes", "scores", "labels"):
            assert torch.equal(r[k], g[k]) and torch.equal(r[k], gg[k]), k


class _FakeSubnet(torch.nn.Module):
    def __init__(self, name, out_ch):
        super().__init__()
        setattr(self, name + "_subnet", torch.nn.Conv2d(8, 8, 3, padding=1))
        setattr(self, name + "_subnet_output", torch.nn.Conv2d(8, out_ch, 3, padding=1))
        self.name, self.k = name, out_ch // 9

    def forward(self, feature_maps):            # the reference's re-layout, layers.py:181-195 / 246-259
        outs = []
        for f in feature_maps:
            x = getattr(self, self.name + "_subnet_output")(getattr(self, self.name + "_subnet")(f))
            N, _, H, W = x.shape
            outs.append(x.view(N, -1, self.k, H, W).permute(0, 3, 4, 1, 2).contiguous().view(N, -1, self.k))
        return torch.cat(outs, dim=1)


class _FakeRetinanet(torch.nn.Module):
    """Mimics the call sites of retinanet/models.py:245-288 (the reference is absent on the GPU box)."""

    def __init__(self, C):
        super().__init__()
        import pytorch_retinanet_b200 as P
        self.num_classes, self.score_thres, self.nms_thres, self.detections_per_img = C, 0.05, 0.5, 100
        self.anchor_generator = P.AnchorGenerator()
        head = torch.nn.Module()
        head.classification_head = _FakeSubnet("class", 9 * C)
        head.regression_head = _FakeSubnet("box", 9 * 4)
        head.losses = None
        head.forward = types.MethodType(lambda h, fm: {"cls_preds": h.classification_head(fm), "bbox_preds": h.regression_head(fm)}, head)
        head.compute_loss = types.MethodType(lambda h, t, o, a: h.losses(t, o, a), head)       # layers.py:100-108
        self.retinanet_head = head

    def run(self, fmaps, targets, image_sizes):
        images = SimpleNamespace(image_sizes=image_sizes)
        outputs = self.retinanet_head(fmaps)
        anchors = self.anchor_generator(images, fmaps)                                         # models.py:284
        losses = self.retinanet_head.compute_loss(targets, outputs, anchors)                   # models.py:287
        outputs2 = self.retinanet_head(fmaps)
        dets = self.process_detections(outputs2, anchors, image_sizes)                         # models.py:270
        return losses, dets


def test_patch_retinanet_with_fused_head_layout():
    import pytorch_retinanet_b200 as P
    torch.manual_seed(0)
    C, hw = 6, (128, 160)
    dev = torch.device("cuda")
    model = _FakeRetinanet(C).to(dev)
    model.process_detections = None
    fmaps = [torch.randn((2, 8, h, w), device=dev) for h, w in S.grid_sizes(hw)]
    gen = torch.Generator().manual_seed(5)
    targets = [{"boxes": S._gt_boxes(gen, 4, hw).to(dev), "labels": torch.randint(1, C + 1, (4,), generator=gen).to(dev)} for _ in range(2)]
    with torch.no_grad():
        model.retinanet_head.classification_head.class_subnet_output.bias.fill_(-2.0)
    P.patch_retinanet(model)
    l0, d0 = model.run(fmaps, targets, [hw, hw])
    P.patch_retinanet(model, fuse_head_layout=True)
    l1, d1 = model.run(fmaps, targets, [hw, hw])
    (l1["classification_loss"] + l1["regression_loss"]).backward()                             # grads reach the convs
    assert model.retinanet_head.classification_head.class_subnet_output.weight.grad.abs().sum() > 0
    for k in l0:
        assert rel_close(l1[k], l0[k].detach(), 1e-6), k
    for a, b in zip(d0, d1):
        assert all(torch.equal(a[k], b[k]) for k in a)


def test_patch_retinanet_fold_box_resize():
    """Row N2 through the integration helper: the patched ``predict`` (models.py:245-272 flow) returns the boxes
    transform.postprocess would have produced, for both transform.training states."""
    import pytorch_retinanet_b200 as P
    torch.manual_seed(1)
    C, hw = 5, (96, 128)
    dev = torch.device("cuda")
    model = _FakeRetinanet(C).to(dev)
    fmaps = [torch.randn((2, 8, h, w), device=dev) for h, w in S.grid_sizes(hw)]
    with torch.no_grad():
        model.retinanet_head.classification_head.class_subnet_output.bias.fill_(-1.5)

    class FakeTransform:
        training = False

        def __call__(self, images, targets):
            return SimpleNamespace(tensors=torch.stack([torch.zeros((3,) + hw, device=dev)] * len(images)),
                                   image_sizes=[(90, 128), (96, 120)]), targets

    model.transform = FakeTransform()
    model.backbone = lambda t: fmaps
    model.fpn = lambda f: f
    model.process_detections = None
    P.patch_retinanet(model, fold_box_resize=True)
    images = [torch.zeros((3, 180, 256), device=dev), torch.zeros((3, 48, 60), device=dev)]
    got = model.predict(images)
    outputs = model.retinanet_head(fmaps)
    anchors = model.anchor_generator(SimpleNamespace(image_sizes=[(90, 128), (96, 120)]), fmaps)
    plain = model.process_detections(outputs, anchors, [(90, 128), (96, 120)])
    for g, d, s_, o in zip(got, plain, [(90, 128), (96, 120)], [(180, 256), (48, 60)]):
        assert torch.equal(g["boxes"], O.resize_boxes(d["boxes"], s_, o)) and torch.equal(g["labels"], d["labels"])
    model.transform.training = True                      # transform.postprocess is a no-op in training mode
    for g, d in zip(model.predict(images), plain):
        assert torch.equal(g["boxes"], d["boxes"])
