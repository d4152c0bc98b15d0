"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes), against the oracle
and the committed golden vectors of the unmodified reference.

Bars (BASELINE.json north_star): bit-exact for matched indices, labels and NMS keep lists; 1e-5
relative for losses and decoded boxes.  Gradients: 1e-5 relative + a 1e-9 absolute floor (values are
O(1e-7..1e-3) and fp32 SFU math is used).
"""
import numpy as np
import pytest
import torch

import synth_data as S
from helpers import config_image, golden, max_rel, rel_close, to_cuda_targets
from oracle import c_oracle as CO
from oracle import torch_oracle as O

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
BOX_RTOL = 1e-5


@pytest.fixture(scope="module")
def P():
    import pytorch_retinanet_b200 as pkg
    from pytorch_retinanet_b200 import _native
    _native.load()
    return pkg


def gpu_anchors(P, padded_hw, **kw):
    gen = P.AnchorGenerator(**kw).cuda()
    return gen.grid_anchors(S.grid_sizes(padded_hw), torch.device("cuda"))


# ------------------------------------------------------------------------------------------ anchors
@pytest.mark.parametrize("cid", [1, 2, 5])
def test_anchor_grid_bit_exact(P, cid):
    hw = S.CONFIGS[cid].padded_hw
    got = torch.cat(gpu_anchors(P, hw)).cpu()
    want = O.image_anchors(O.fpn_grid_sizes(*hw))
    assert got.shape == want.shape and torch.equal(got, want)


def test_anchor_generator_api(P):
    g = golden("known_answers.npz")
    gen = P.AnchorGenerator(offset=0.5).cuda()
    assert gen.num_anchors == [9] * 5 and gen.num_cell_anchors == [9] * 5
    assert sorted(gen.state_dict().keys()) == [f"cell_anchors.{i}" for i in range(5)]
    got = torch.cat(gen.grid_anchors(S.grid_sizes((64, 96)), torch.device("cuda"))).cpu().numpy()
    assert np.array_equal(got, g["anchors_off05_64x96"])
    # forward(): ImageList-like object, one entry per image, levels concatenated
    from types import SimpleNamespace
    fmaps = [torch.empty((2, 8, h, w), device="cuda") for h, w in S.grid_sizes((64, 96))]
    out = gen(SimpleNamespace(image_sizes=[(60, 90), (64, 96)]), fmaps)
    assert len(out) == 2 and np.array_equal(out[1].cpu().numpy(), g["anchors_off05_64x96"])
    # custom sizes / ratios / strides (3 levels, 2x2 cell anchors)
    gen2 = P.AnchorGenerator(sizes=[[16.0, 20.0]], aspect_ratios=[0.5, 2.0], strides=[4, 8, 16]).cuda()
    gs = [(5, 7), (3, 4), (2, 2)]
    got2 = torch.cat(gen2.grid_anchors(gs, torch.device("cuda"))).cpu()
    want2 = O.image_anchors(gs, strides=[4, 8, 16], sizes=[[16.0, 20.0]] * 3, ratios=[0.5, 2.0])
    assert torch.equal(got2, want2)


# ------------------------------------------------------------------------------------------ matcher
def test_matcher_known_answers(P):
    g = golden("known_answers.npz")
    anc, gt = torch.from_numpy(g["m_anchors"]).cuda(), torch.from_numpy(g["m_gt"]).cuda()
    assert P.matcher(anc, gt).cpu().tolist() == [0, -1, -2]
    assert P.matcher(anc, torch.zeros((0, 4), device="cuda")).cpu().tolist() == [-2, -2, -2]
    with pytest.raises(AssertionError):
        P.matcher(anc, gt, match_thr=0.3, back_thr=0.4)
    m = P.matcher(anc, gt)
    assert m.dtype == torch.int64 and m.shape == (3,)


@pytest.mark.parametrize("cid", [1, 2, 5])
def test_matcher_full_size_bit_exact(P, cid):
    b, g = config_image(cid)
    m = P.matcher(b["anchors"].cuda(), b["targets"][0]["boxes"].cuda()).cpu().numpy()
    assert np.array_equal(m, g["matches"])


def test_matcher_random_small_golden(P):
    g = golden("random_small.npz")
    for k in range(int(g["n_cases"])):
        p = f"c{k}_"
        m = P.matcher(torch.from_numpy(g[p + "anchors"]).cuda(), torch.from_numpy(g[p + "gt"]).cuda())
        assert np.array_equal(m.cpu().numpy(), g[p + "matches"]), k


def test_matcher_edge_cases_vs_oracle(P):
    gen = torch.Generator().manual_seed(5)
    anc = S.default_anchors((256, 320))
    cases = []
    gt_big = S._gt_boxes(gen, 1300, (256, 320))                 # > 2 GT tiles of 512
    cases.append(("g1300", gt_big, None, None))
    cases.append(("g1", S._gt_boxes(gen, 1, (256, 320)), None, None))
    dup = anc[torch.randint(0, anc.shape[0], (40,), generator=gen)]
    cases.append(("exact_copies", torch.cat([dup, dup]), None, None))  # IoU==1 ties -> lowest index
    bad = S._gt_boxes(gen, 12, (256, 320))
    bad[3, 0] = float("nan")                                    # NaN propagates -> everything ignore
    cases.append(("nan_gt", bad, None, None))
    neg = S._gt_boxes(gen, 12, (256, 320))
    neg[5] = torch.tensor([200., 200., 100., 120.])             # x2<x1: negative area
    neg[7] = torch.tensor([50., 60., 50., 60.])                 # degenerate, area 0
    cases.append(("malformed_gt", neg, None, None))
    cases.append(("thr_0.7_0.3", S._gt_boxes(gen, 30, (256, 320)), 0.7, 0.3))
    cases.append(("bg_thr_0", S._gt_boxes(gen, 30, (256, 320)), 0.5, 0.0))      # disables culling
    cases.append(("neg_thr", S._gt_boxes(gen, 30, (256, 320)), -0.1, -0.5))     # every anchor foreground
    for name, gt, ft, bt in cases:
        want = O.match(anc, gt, ft, bt)
        got = P.matcher(anc.cuda(), gt.cuda(), ft, bt).cpu()
        assert torch.equal(got, want), name
    # malformed anchors (zero area / NaN) take the generic path
    anc2 = anc[:4096].clone()
    anc2[17] = torch.tensor([10., 10., 10., 30.])
    anc2[99, 2] = float("nan")
    gt = S._gt_boxes(gen, 25, (256, 320))
    assert torch.equal(P.matcher(anc2.cuda(), gt.cuda()).cpu(), O.match(anc2, gt))


# ------------------------------------------------------------------------------------------ box coding
def test_box_coding(P):
    g = golden("known_answers.npz")
    dec = P.activ_2_bbox(torch.from_numpy(g["dec_act"]).cuda(), torch.from_numpy(g["dec_anchor"]).cuda())
    assert rel_close(dec, g["dec_out"], BOX_RTOL)
    enc = P.bbox_2_activ(torch.from_numpy(g["enc_gt"]).cuda(), torch.from_numpy(g["enc_anchor"]).cuda())
    assert rel_close(enc, g["enc_out"], BOX_RTOL)
    gen = torch.Generator().manual_seed(3)
    anc = S.default_anchors((128, 128))
    act = torch.randn((anc.shape[0], 4), generator=gen) * 0.5
    want = O.decode(act, anc)
    got = P.activ_2_bbox(act.cuda(), anc.cuda()).cpu()
    assert rel_close(got, want, BOX_RTOL, 1e-4)                  # cancellation near 0: tiny absolute floor
    # on the same device (CUDA expf) the decode is bit-identical to the eager op sequence
    assert torch.equal(got, O.decode(act.cuda(), anc.cuda()).cpu())
    gt = S._gt_boxes(gen, anc.shape[0], (128, 128))
    assert rel_close(P.bbox_2_activ(gt.cuda(), anc.cuda()), O.encode(gt, anc), BOX_RTOL, 1e-6)
    xywh = P.convert_xywh(anc.cuda())
    assert torch.equal(P.convert_x1y1x2y2(xywh).cpu(), O.to_corners(O.to_center_size(anc)))


# ------------------------------------------------------------------------------------------ losses
def run_gpu_loss(P, cls, bb, targets, anchors_list, C, grad=True):
    L = P.RetinaNetLosses(C)
    x = cls.cuda().requires_grad_(grad)
    b = bb.cuda().requires_grad_(grad)
    out = L(to_cuda_targets(targets), {"cls_preds": x, "bbox_preds": b}, anchors_list)
    if grad:
        (out["classification_loss"] + out["regression_loss"]).backward()
    return out, x, b, L


def test_loss_known_answers(P):
    g = golden("known_answers.npz")
    anc = torch.from_numpy(g["l3_anchors"]).cuda()
    tg = [{"boxes": torch.from_numpy(g["l3_gt"]), "labels": torch.from_numpy(g["l3_labels"])}]
    out, x, b, L = run_gpu_loss(P, torch.from_numpy(g["l3_cls"]), torch.from_numpy(g["l3_bb"]), tg, [anc], 3)
    assert rel_close(out["classification_loss"], g["l3_closs"], LOSS_RTOL)
    assert rel_close(out["regression_loss"], g["l3_rloss"], LOSS_RTOL)
    assert rel_close(x.grad, g["l3_gcls"], 1e-5, 1e-9) and rel_close(b.grad, g["l3_gbb"], 1e-5, 1e-9)
    assert L.last_per_image.cpu()[0, 2] == 2
    # empty GT: everything ignored, both losses exactly 0 and zero gradients
    tg0 = [{"boxes": torch.zeros((0, 4)), "labels": torch.zeros((0,), dtype=torch.int64)}]
    out0, x0, b0, _ = run_gpu_loss(P, torch.from_numpy(g["l3_cls"]), torch.from_numpy(g["l3_bb"]), tg0, [anc], 3)
    assert float(out0["classification_loss"]) == 0.0 and float(out0["regression_loss"]) == 0.0
    assert float(x0.grad.abs().sum()) == 0.0 and float(b0.grad.abs().sum()) == 0.0
    # calc_loss keeps the reference's (bb_loss, clas_loss) order
    bl, cl = L.calc_loss(anc, torch.from_numpy(g["l3_cls"][0]).cuda(), torch.from_numpy(g["l3_bb"][0]).cuda(),
                         tg[0]["labels"].cuda(), tg[0]["boxes"].cuda())
    assert rel_close(cl, g["l3_closs"], LOSS_RTOL) and rel_close(bl, g["l3_rloss"], LOSS_RTOL)


def test_loss_random_small_golden(P):
    """Ragged/empty GT, ties, C in {1,3,4,5,8,20} (scalar and 128-bit paths) vs the reference's outputs."""
    g = golden("random_small.npz")
    for k in range(int(g["n_cases"])):
        p = f"c{k}_"
        anc = torch.from_numpy(g[p + "anchors"]).cuda()
        tg = [{"boxes": torch.from_numpy(g[p + "gt"]), "labels": torch.from_numpy(g[p + "labels"])}]
        C = g[p + "cls"].shape[-1]
        out, x, b, _ = run_gpu_loss(P, torch.from_numpy(g[p + "cls"]), torch.from_numpy(g[p + "bb"]), tg, [anc], C)
        assert rel_close(out["classification_loss"], g[p + "closs"], LOSS_RTOL, 1e-7), (k, float(out["classification_loss"]), g[p + "closs"])
        assert rel_close(out["regression_loss"], g[p + "rloss"], LOSS_RTOL, 1e-7), k
        assert rel_close(x.grad, g[p + "gcls"], 2e-5, 1e-8), (k, max_rel(x.grad, g[p + "gcls"], 1e-6))
        assert rel_close(b.grad, g[p + "gbb"], 2e-5, 1e-8), k


@pytest.mark.parametrize("cid", [1, 2, 5])
def test_loss_full_size_vs_golden(P, cid):
    b, g = config_image(cid)
    anc = b["anchors"].cuda()
    out, x, bb, L = run_gpu_loss(P, b["cls_preds"], b["bbox_preds"], b["targets"], [anc], b["config"].num_classes)
    assert rel_close(out["classification_loss"], g["closs"], LOSS_RTOL), (float(out["classification_loss"]), g["closs"])
    assert rel_close(out["regression_loss"], g["rloss"], LOSS_RTOL), (float(out["regression_loss"]), g["rloss"])
    gx = x.grad.reshape(-1)[torch.from_numpy(g["g_idx"]).cuda()]
    assert rel_close(gx, g["gcls_at_idx"], 2e-5, 1e-12), max_rel(gx, g["gcls_at_idx"], 1e-9)
    rows = torch.from_numpy(g["g_fgrows"]).cuda()
    assert rel_close(x.grad[0, rows], g["gcls_fgrows"], 2e-5, 1e-12)
    assert rel_close(bb.grad[0, rows], g["gbb_fgrows"], 2e-5, 1e-9)
    assert rel_close(x.grad.double().abs().sum(), g["gcls_abs_sum"], 1e-5)
    assert rel_close(bb.grad.double().abs().sum(), g["gbb_abs_sum"], 1e-5)
    assert int(L.last_per_image[0, 2]) == int((g["matches"] >= 0).sum())


def test_loss_batch_precise_mode_and_determinism(P):
    """Batch of 3 images with different G (incl. one empty), fast vs precise math, bit-reproducibility,
    per-image anchor tensors (stacked path) and grad_output scaling."""
    from pytorch_retinanet_b200 import _native
    cfg = S.CONFIGS[1]
    b = S.make_batch(cfg, 10, 3, clustered=True)
    b["targets"][1] = {"boxes": torch.zeros((0, 4)), "labels": torch.zeros((0,), dtype=torch.int64)}
    anc = b["anchors"]
    xo = b["cls_preds"].clone().requires_grad_(True)
    bo = b["bbox_preds"].clone().requires_grad_(True)
    want = O.batch_loss(b["targets"], xo, bo, [anc] * 3, cfg.num_classes)
    (2.0 * want["classification_loss"] + 0.5 * want["regression_loss"]).backward()
    anc_g = anc.cuda()
    res = {}
    for mode in (0, 1):
        _native.load().rn_loss_set_math_mode(mode)
        L = P.RetinaNetLosses(cfg.num_classes)
        x = b["cls_preds"].cuda().requires_grad_(True)
        bb = b["bbox_preds"].cuda().requires_grad_(True)
        out = L(to_cuda_targets(b["targets"]), {"cls_preds": x, "bbox_preds": bb}, [anc_g] * 3)
        (2.0 * out["classification_loss"] + 0.5 * out["regression_loss"]).backward()
        res[mode] = (out, x.grad.clone(), bb.grad.clone())
        assert rel_close(out["classification_loss"], want["classification_loss"].detach(), LOSS_RTOL), mode
        assert rel_close(out["regression_loss"], want["regression_loss"].detach(), LOSS_RTOL), mode
        assert rel_close(x.grad, xo.grad, 2e-5, 1e-12), (mode, max_rel(x.grad, xo.grad, 1e-9))
        assert rel_close(bb.grad, bo.grad, 2e-5, 1e-9), mode
    _native.load().rn_loss_set_math_mode(0)
    # bit-reproducible run to run (fixed-order reductions, no float atomics)
    L = P.RetinaNetLosses(cfg.num_classes)
    outs = []
    for _ in range(3):
        o = L(to_cuda_targets(b["targets"]), {"cls_preds": b["cls_preds"].cuda(), "bbox_preds": b["bbox_preds"].cuda()},
              [anc_g] * 3)
        outs.append((float(o["classification_loss"]), float(o["regression_loss"])))
    assert outs[0] == outs[1] == outs[2]
    # distinct per-image anchor tensors -> stacked [N,A,4] path gives the same numbers
    o2 = L(to_cuda_targets(b["targets"]), {"cls_preds": b["cls_preds"].cuda(), "bbox_preds": b["bbox_preds"].cuda()},
           [anc_g.clone() for _ in range(3)])
    assert float(o2["classification_loss"]) == outs[0][0] and float(o2["regression_loss"]) == outs[0][1]


def test_loss_generic_gamma_and_l1(P):
    """gamma != 2 (pow path) and beta < 1e-5 (pure L1) against the oracle."""
    cfg = S.CONFIGS[1]
    b = S.make_batch(cfg, 20, 1, clustered=True)
    anc = b["anchors"]
    for gamma, beta in ((1.5, 0.1), (2.0, 0.0), (0.0, 0.5)):
        xo = b["cls_preds"].clone().requires_grad_(True)
        bo = b["bbox_preds"].clone().requires_grad_(True)
        want = O.batch_loss(b["targets"], xo, bo, [anc], cfg.num_classes, gamma=gamma, beta=beta)
        (want["classification_loss"] + want["regression_loss"]).backward()
        L = P.RetinaNetLosses(cfg.num_classes)
        L.gamma, L.beta = gamma, beta
        x = b["cls_preds"].cuda().requires_grad_(True)
        bb = b["bbox_preds"].cuda().requires_grad_(True)
        out = L(to_cuda_targets(b["targets"]), {"cls_preds": x, "bbox_preds": bb}, [anc.cuda()])
        (out["classification_loss"] + out["regression_loss"]).backward()
        assert rel_close(out["classification_loss"], want["classification_loss"].detach(), LOSS_RTOL), (gamma, beta)
        assert rel_close(out["regression_loss"], want["regression_loss"].detach(), LOSS_RTOL), (gamma, beta)
        assert rel_close(x.grad, xo.grad, 2e-5, 1e-11), (gamma, beta, max_rel(x.grad, xo.grad, 1e-8))
        assert rel_close(bb.grad, bo.grad, 2e-5, 1e-9), (gamma, beta)


def test_backward_twice_and_one_loss_at_a_time(P):
    """Every autograd usage the reference supports: the two losses back-propagated one after the other
    (retain_graph), a second backward through the same graph, and autograd.grad after the kernel-written gradient
    buffers were already handed over (they are recomputed from the saved inputs)."""
    cfg = S.CONFIGS[1]
    b = S.make_batch(cfg, 40, 2, clustered=True)
    anc, anc_g = b["anchors"], b["anchors"].cuda()
    tg = to_cuda_targets(b["targets"])
    xo = b["cls_preds"].clone().requires_grad_(True)
    bo = b["bbox_preds"].clone().requires_grad_(True)
    want = O.batch_loss(b["targets"], xo, bo, [anc] * 2, cfg.num_classes)
    want["classification_loss"].backward(retain_graph=True)
    assert bo.grad is None
    want["regression_loss"].backward()
    L = P.RetinaNetLosses(cfg.num_classes)
    x = b["cls_preds"].cuda().requires_grad_(True)
    bb = b["bbox_preds"].cuda().requires_grad_(True)
    out = L(tg, {"cls_preds": x, "bbox_preds": bb}, [anc_g] * 2)
    out["classification_loss"].backward(retain_graph=True)
    assert bb.grad is None and x.grad is not None           # nothing flows from the class loss to the boxes
    out["regression_loss"].backward(retain_graph=True)
    assert rel_close(x.grad, xo.grad, 2e-5, 1e-12) and rel_close(bb.grad, bo.grad, 2e-5, 1e-9)
    assert float(bb.grad.abs().sum()) > 0
    gx1, gb1 = x.grad.clone(), bb.grad.clone()
    # a second pass through the same graph accumulates the same gradients again (buffers recomputed)
    (out["classification_loss"] + out["regression_loss"]).backward(retain_graph=True)
    assert torch.equal(x.grad, gx1 + gx1) and torch.equal(bb.grad, gb1 + gb1)
    g3 = torch.autograd.grad(3.0 * out["regression_loss"], [bb])[0]
    assert rel_close(g3, 3.0 * bo.grad, 2e-5, 1e-9)
    # same for the per-level entry point
    cls_lv = [t.cuda().requires_grad_(True) for t in S.nac_to_levels(b["cls_preds"], cfg.padded_hw)]
    box_lv = [t.cuda().requires_grad_(True) for t in S.nac_to_levels(b["bbox_preds"], cfg.padded_hw)]
    ol = L(tg, {"cls_levels": cls_lv, "bbox_levels": box_lv}, [anc_g] * 2)
    ol["regression_loss"].backward(retain_graph=True)
    assert all(t.grad is None for t in cls_lv)
    ol["classification_loss"].backward(retain_graph=True)
    first = [t.grad.clone() for t in cls_lv + box_lv]
    (ol["classification_loss"] + ol["regression_loss"]).backward()
    for t, f in zip(cls_lv + box_lv, first):
        assert torch.equal(t.grad, f + f)
    want_lv = S.nac_to_levels(xo.grad, cfg.padded_hw)
    for t, w in zip(cls_lv, want_lv):
        assert rel_close(t.grad, 2.0 * w, 2e-5, 1e-12)


def test_labels_outside_one_to_C(P):
    """Label 0 is the reference's class id 0: the one-hot row is all zeros after the [:,1:] slice (losses.py:96-103),
    the anchor stays foreground for the regression term and the F count.  Labels the reference's one_hot rejects
    (> C) behave the same way here (documented: "no class column") and never corrupt the packed code."""
    cfg = S.CONFIGS[1]
    b = S.make_batch(cfg, 50, 2, clustered=True)
    anc = b["anchors"]
    t0 = [{"boxes": t["boxes"], "labels": t["labels"].clone()} for t in b["targets"]]
    for t in t0:
        t["labels"][::2] = 0
    xo = b["cls_preds"].clone().requires_grad_(True)
    bo = b["bbox_preds"].clone().requires_grad_(True)
    want = O.batch_loss(t0, xo, bo, [anc] * 2, cfg.num_classes)
    (want["classification_loss"] + want["regression_loss"]).backward()
    assert float(want["regression_loss"]) > 0
    L = P.RetinaNetLosses(cfg.num_classes)
    res = []
    for weird in (0, cfg.num_classes + 5, 2047, 2048, 5000, -3, 1 << 40):
        tw = [{"boxes": t["boxes"], "labels": t["labels"].clone()} for t in b["targets"]]
        for t in tw:
            t["labels"][::2] = weird
        x = b["cls_preds"].cuda().requires_grad_(True)
        bb = b["bbox_preds"].cuda().requires_grad_(True)
        out = L(to_cuda_targets(tw), {"cls_preds": x, "bbox_preds": bb}, [anc.cuda()] * 2)
        (out["classification_loss"] + out["regression_loss"]).backward()
        assert rel_close(out["classification_loss"], want["classification_loss"].detach(), LOSS_RTOL), weird
        assert rel_close(out["regression_loss"], want["regression_loss"].detach(), LOSS_RTOL), weird
        assert rel_close(x.grad, xo.grad, 2e-5, 1e-12) and rel_close(bb.grad, bo.grad, 2e-5, 1e-9), weird
        assert torch.equal(L.last_per_image[:, 2].cpu(), torch.tensor([float(O.match(anc, t["boxes"]).ge(0).sum()) for t in tw]))
        res.append((float(out["classification_loss"]), float(out["regression_loss"])))
    assert all(r == res[0] for r in res)


def test_dense_focal_and_smooth_l1_methods(P):
    """RetinaNetLosses.focal_loss / smooth_l1_loss (losses.py:19-47) on dense targets, with gradients."""
    gen = torch.Generator().manual_seed(8)
    x = torch.randn((3001, 7), generator=gen) * 3
    t = (torch.rand((3001, 7), generator=gen) < 0.1).float()
    t[5] = torch.rand(7, generator=gen)                            # soft targets are legal inputs too
    for gamma, alpha in ((2.0, 0.25), (1.5, 0.4)):
        xo = x.clone().requires_grad_(True)
        want = O.focal_sum(xo, t, alpha, gamma)
        want.backward()
        L = P.RetinaNetLosses(7)
        L.gamma, L.alpha = gamma, alpha
        xg = x.cuda().requires_grad_(True)
        got = L.focal_loss(xg, t.cuda())
        (got * 2.0).backward()
        assert rel_close(got, want.detach(), LOSS_RTOL), (gamma, float(got), float(want))
        assert rel_close(xg.grad, 2.0 * xo.grad, 2e-5, 1e-9)
    a, b = torch.randn((513, 4), generator=gen), torch.randn((513, 4), generator=gen) * 0.2
    for beta in (0.1, 0.0):
        ao = a.clone().requires_grad_(True)
        want = O.smooth_l1_sum(ao, b, beta)
        want.backward()
        L = P.RetinaNetLosses(3)
        L.beta = beta
        ag = a.cuda().requires_grad_(True)
        got = L.smooth_l1_loss(ag, b.cuda())
        got.backward()
        assert rel_close(got, want.detach(), LOSS_RTOL) and rel_close(ag.grad, ao.grad, 1e-5, 1e-9), beta


# ------------------------------------------------------------------------------------------ post-processing
def gpu_detect(P, cls, bb, anchors_list, im_szs, algo="auto", **kw):
    """algo="auto": through the drop-in process_detections(self, outputs, anchors, im_szs);
    "lazy"/"general": the same C-ABI call with the algorithm forced."""
    from types import SimpleNamespace
    score, nms, max_det = kw.get("score", 0.05), kw.get("nms", 0.5), kw.get("max_det", 100)
    if algo == "auto":
        stub = SimpleNamespace(score_thres=score, nms_thres=nms, detections_per_img=max_det)
        outputs = {"cls_preds": cls.cuda(), "bbox_preds": bb.cuda()}
        dets = P.process_detections(stub, outputs, anchors_list, im_szs)
        assert outputs == {}                                      # models.py:168-169 pops both keys
        return dets
    from pytorch_retinanet_b200.detections import postprocess_batch
    from pytorch_retinanet_b200.losses import _shared_anchors
    an, stride = _shared_anchors(anchors_list)
    ob, os_, ol, counts = postprocess_batch(cls.cuda(), bb.cuda(), an, stride, im_szs, score, nms, max_det, algo=algo)
    return [{"boxes": ob[i, :k], "scores": os_[i, :k], "labels": ol[i, :k]} for i, k in enumerate(counts)]


ALGOS = ["auto", "lazy", "general"]


def assert_dets_equal(got, want, exact=True, ctx=""):
    assert got["labels"].dtype == torch.int64 and got["boxes"].dtype == torch.float32
    assert got["boxes"].shape == want["boxes"].shape, (ctx, got["boxes"].shape, want["boxes"].shape)
    assert torch.equal(got["labels"].cpu(), want["labels"].cpu()), ctx
    if exact:
        assert torch.equal(got["scores"].cpu(), want["scores"].cpu()), ctx
        assert torch.equal(got["boxes"].cpu(), want["boxes"].cpu()), ctx
    else:
        assert rel_close(got["scores"], want["scores"], BOX_RTOL), ctx
        assert rel_close(got["boxes"], want["boxes"], BOX_RTOL, 1e-4), ctx


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("cid", [1, 2, 5])
def test_postprocess_full_size(P, cid, algo):
    """Bit-exact against the eager op sequence on the same device (CUDA oracle: same expf), and
    against the CPU reference's golden detections: same labels/order, floats within 1e-5."""
    b, g = config_image(cid)
    anc = b["anchors"].cuda()
    got = gpu_detect(P, b["cls_preds"], b["bbox_preds"], [anc], b["im_szs"], algo=algo)[0]
    want_cuda = O.postprocess(b["cls_preds"].cuda(), b["bbox_preds"].cuda(), [anc], b["im_szs"])[0]
    assert_dets_equal(got, want_cuda, exact=True, ctx=f"config{cid} vs CUDA oracle")
    want_cpu = {"boxes": torch.from_numpy(g["det_boxes"]), "scores": torch.from_numpy(g["det_scores"]),
                "labels": torch.from_numpy(g["det_labels"])}
    assert_dets_equal(got, want_cpu, exact=False, ctx=f"config{cid} vs CPU reference golden")


@pytest.mark.parametrize("algo", ["auto", "general"])
def test_postprocess_random_small_golden(P, algo):
    g = golden("random_small.npz")
    for k in range(int(g["n_cases"])):
        p = f"c{k}_"
        anc = torch.from_numpy(g[p + "anchors"]).cuda()
        cls, bb = torch.from_numpy(g[p + "cls"]), torch.from_numpy(g[p + "bb"])
        got = gpu_detect(P, cls, bb, [anc], [(90, 120)], algo=algo, max_det=20)[0]
        want = O.postprocess(cls.cuda(), bb.cuda(), [anc], [(90, 120)], max_det=20)[0]
        assert_dets_equal(got, want, exact=True, ctx=f"case {k} vs CUDA oracle")
        ref = {"boxes": torch.from_numpy(g[p + "det_boxes"]), "scores": torch.from_numpy(g[p + "det_scores"]),
               "labels": torch.from_numpy(g[p + "det_labels"])}
        if got["boxes"].shape == ref["boxes"].shape and torch.equal(got["labels"].cpu(), ref["labels"]):
            assert_dets_equal(got, ref, exact=False, ctx=f"case {k} vs reference golden")
        else:   # a 1-ulp CPU/GPU sigmoid/exp difference flipped a threshold/NMS decision: must be rare
            pytest.fail(f"case {k}: detections differ from the CPU reference golden")


@pytest.mark.parametrize("algo", ["auto", "general"])
def test_postprocess_batch_edge_cases(P, algo):
    """Batch of 5 with: an image without any candidate, huge segments (global-memory sort path and
    multi-chunk NMS), heavy score ties, max_det < kept, C not a multiple of 4, an image where nearly
    everything is suppressed (lazy algorithm runs out of rounds -> general fallback), pool overflow."""
    gen = torch.Generator().manual_seed(11)
    anc = S.default_anchors((128, 160))
    A, C = anc.shape[0], 6
    cls = torch.randn((5, A, C), generator=gen) * 1.5 - 6.0
    bb = torch.randn((5, A, 4), generator=gen) * 0.2
    cls[1] = -20.0                                                  # no candidates at all
    cls[2, :, 2] = torch.randn(A, generator=gen) * 0.5 + 1.0        # one class fires on every anchor (A > 2048)
    cls[3, :, 1] = 0.75                                             # thousands of exactly tied scores
    cls[3, :, 4] = torch.round(torch.randn(A, generator=gen) * 2) / 2  # few distinct values -> many ties
    cls[4, :, :3] = torch.randn((A, 3), generator=gen) + 2.0        # ~11k candidates, tiny image: almost all suppressed
    sz = [(128, 160), (100, 150), (128, 160), (120, 130), (6, 5)]
    ancs = [anc.cuda()] * 5
    want = O.postprocess(cls.cuda(), bb.cuda(), ancs, sz, max_det=50)
    got = gpu_detect(P, cls, bb, ancs, sz, algo=algo, max_det=50)
    assert got[1]["boxes"].shape == (0, 4) and got[1]["scores"].shape == (0,) and got[1]["labels"].shape == (0,)
    for i in range(5):
        assert_dets_equal(got[i], want[i], exact=True, ctx=f"image {i}")
    if algo == "auto":      # image 4 (most boxes clipped away) may need the general fallback; 0-3 never do
        got_l = gpu_detect(P, cls[:4], bb[:4], ancs[:4], sz[:4], algo="lazy", max_det=50)
        for i in range(4):
            assert_dets_equal(got_l[i], want[i], exact=True, ctx=f"lazy image {i}")
    # candidate pool overflow -> transparent re-run with the exact capacity
    from pytorch_retinanet_b200.detections import postprocess_batch
    ob, os_, ol, counts = postprocess_batch(cls.cuda(), bb.cuda(), anc.cuda(), 0, sz, 0.05, 0.5, 50, cand_capacity=100,
                                            algo=algo)
    for i in range(5):
        k = counts[i]
        assert torch.equal(ob[i, :k].cpu(), want[i]["boxes"].cpu()) and torch.equal(ol[i, :k].cpu(), want[i]["labels"].cpu())
    # other thresholds
    want2 = O.postprocess(cls.cuda(), bb.cuda(), ancs, sz, score_thr=0.3, nms_thr=0.35, max_det=100)
    got2 = gpu_detect(P, cls, bb, ancs, sz, algo=algo, score=0.3, nms=0.35, max_det=100)
    for i in range(5):
        assert_dets_equal(got2[i], want2[i], exact=True, ctx=f"thr image {i}")


def test_postprocess_lazy_fallback(P):
    """Every anchor decodes to the same box: one detection per class survives, so the lazy algorithm
    must look at ALL 3*6000 candidates, exhausts its round budget and the host falls back to the
    general algorithm (identical result)."""
    from pytorch_retinanet_b200 import _native
    A, C = 6000, 4
    gen = torch.Generator().manual_seed(4)
    anc = torch.tensor([[10., 12., 50., 60.]]).repeat(A, 1)
    cls = torch.full((2, A, C), -9.0)
    cls[0, :, :3] = torch.rand((A, 3), generator=gen) * 4.0
    cls[1, :500, 1] = torch.rand(500, generator=gen)               # image 1 is easy (one class, 500 candidates)
    bb = torch.zeros((2, A, 4))
    sz = [(100, 100), (100, 100)]
    want = O.postprocess(cls.cuda(), bb.cuda(), [anc.cuda()] * 2, sz)
    assert want[0]["boxes"].shape[0] == 3 and want[1]["boxes"].shape[0] == 1
    with pytest.raises(_native.NativeError):
        gpu_detect(P, cls, bb, [anc.cuda()] * 2, sz, algo="lazy")
    for algo in ("auto", "general"):
        got = gpu_detect(P, cls, bb, [anc.cuda()] * 2, sz, algo=algo)
        for i in range(2):
            assert_dets_equal(got[i], want[i], exact=True, ctx=f"{algo} image {i}")


@pytest.mark.parametrize("cid,topk", [(1, 50), (1, 1000), (2, 1000), (2, 7)])
def test_postprocess_pre_nms_topk_extension(P, cid, topk):
    """`pre_nms_topk` per (image, pyramid level) — BASELINE.json configs[3] lists top-k 1000/level; the
    reference has no such stage, so the oracle is the reference code + the documented filter
    (SURVEY.md §8 'Reconciling north_star with A19')."""
    from types import SimpleNamespace
    b, _ = config_image(cid)
    dev = torch.device("cuda")
    gen = P.AnchorGenerator().cuda()
    fmaps = [torch.empty((1, 1, h, w), device=dev) for h, w in S.grid_sizes(b["config"].padded_hw)]
    anchors = gen(SimpleNamespace(image_sizes=b["im_szs"]), fmaps)
    offs = gen.last_level_offsets
    assert offs[-1] == b["anchors"].shape[0] and len(offs) == 6
    model = SimpleNamespace(score_thres=0.05, nms_thres=0.5, detections_per_img=100, pre_nms_topk=topk,
                            anchor_generator=gen)
    got = P.process_detections(model, {"cls_preds": b["cls_preds"].cuda(), "bbox_preds": b["bbox_preds"].cuda()},
                               anchors, b["im_szs"])[0]
    want = O.postprocess(b["cls_preds"].cuda(), b["bbox_preds"].cuda(), anchors, b["im_szs"],
                         pre_nms_topk=topk, level_offsets=offs)[0]
    assert_dets_equal(got, want, exact=True, ctx=f"config{cid} topk={topk}")
    # a top-k that no level reaches is a no-op
    if topk >= 1000 and cid == 1:
        plain = P.process_detections(SimpleNamespace(score_thres=0.05, nms_thres=0.5, detections_per_img=100),
                                     {"cls_preds": b["cls_preds"].cuda(), "bbox_preds": b["bbox_preds"].cuda()},
                                     anchors, b["im_szs"])[0]
        assert_dets_equal(got, plain, exact=True, ctx="no-op topk")


@pytest.mark.parametrize("cid,n_img", [(2, 16), (5, 6)])
def test_full_batch_properties(P, cid, n_img):
    """Size-independent properties at the full per-image sizes of configs 2 (batch 16) and 5 (G=500):
    batch loss = mean of the single-image losses, batch gradients = single-image gradients / N,
    batch detections = single-image detections, invariance under a permutation of the images,
    linearity of backward in grad_output, bit-reproducibility, and loss consistency with the matcher
    (F_i from the loss == number of non-negative matches)."""
    from types import SimpleNamespace
    cfg = S.CONFIGS[cid]
    b = S.make_batch(cfg, 40, n_img)
    dev = torch.device("cuda")
    anc = b["anchors"].to(dev)
    tg = to_cuda_targets(b["targets"])
    x, bb = b["cls_preds"].to(dev), b["bbox_preds"].to(dev)
    L = P.RetinaNetLosses(cfg.num_classes)
    xb, bbb = x.clone().requires_grad_(True), bb.clone().requires_grad_(True)
    out = L(tg, {"cls_preds": xb, "bbox_preds": bbb}, [anc] * n_img)
    (out["classification_loss"] + 3.0 * out["regression_loss"]).backward()
    per_image = L.last_per_image.clone()
    stub = SimpleNamespace(score_thres=0.05, nms_thres=0.5, detections_per_img=100)
    dets = P.process_detections(stub, {"cls_preds": x, "bbox_preds": bb}, [anc] * n_img, b["im_szs"])
    cls_sum = reg_sum = 0.0
    for i in (0, n_img // 2, n_img - 1):
        xi, bi = x[i:i + 1].clone().requires_grad_(True), bb[i:i + 1].clone().requires_grad_(True)
        oi = L(tg[i:i + 1], {"cls_preds": xi, "bbox_preds": bi}, [anc])
        (oi["classification_loss"] + oi["regression_loss"]).backward()
        assert rel_close(oi["classification_loss"], per_image[i, 0], 1e-6) and rel_close(oi["regression_loss"], per_image[i, 1], 1e-6)
        assert rel_close(xb.grad[i] * n_img, xi.grad[0], 1e-6, 1e-12)
        assert rel_close(bbb.grad[i] * n_img / 3.0, bi.grad[0], 2e-6, 1e-12)       # linear in grad_output
        di = P.process_detections(stub, {"cls_preds": x[i:i + 1], "bbox_preds": bb[i:i + 1]}, [anc], b["im_szs"][i:i + 1])[0]
        assert_dets_equal(dets[i], di, exact=True, ctx=f"image {i}")
        m = P.matcher(anc, tg[i]["boxes"])
        assert int(per_image[i, 2]) == int((m >= 0).sum())
    assert rel_close(out["classification_loss"], per_image[:, 0].double().mean(), 1e-6)
    assert rel_close(out["regression_loss"], per_image[:, 1].double().mean(), 1e-6)
    perm = torch.randperm(n_img, generator=torch.Generator().manual_seed(1)).tolist()
    outp = L([tg[i] for i in perm], {"cls_preds": x[perm], "bbox_preds": bb[perm]}, [anc] * n_img)
    assert rel_close(outp["classification_loss"], out["classification_loss"].detach(), 1e-6)
    assert torch.equal(L.last_per_image[:, :], per_image[perm])                     # per-image values are bit-stable
    detp = P.process_detections(stub, {"cls_preds": x[perm], "bbox_preds": bb[perm]}, [anc] * n_img, [b["im_szs"][i] for i in perm])
    for j, i in enumerate(perm):
        assert_dets_equal(detp[j], dets[i], exact=True, ctx=f"perm {j}")
    for d in dets:                                                                  # sorted, capped, labelled 1..C
        assert d["boxes"].shape[0] <= 100 and bool((d["scores"][:-1] >= d["scores"][1:]).all())
        assert int(d["labels"].min()) >= 1 and int(d["labels"].max()) <= cfg.num_classes
        assert bool((d["boxes"][:, 0] >= 0).all()) and bool((d["boxes"][:, 2] <= b["im_szs"][0][1]).all())


def test_output_epilogue_resize_and_coco(P):
    """Rows N2 / N4: the box resize of transform.postprocess (torchvision resize_boxes) and the COCO xywh
    conversion folded into the detection write are bit-identical to applying them afterwards."""
    from types import SimpleNamespace
    from pytorch_retinanet_b200.detections import postprocess_batch_async
    b, _ = config_image(1)
    cfg = b["config"]
    dev = torch.device("cuda")
    x = torch.cat([b["cls_preds"], b["cls_preds"].flip(1)]).to(dev)
    bb = torch.cat([b["bbox_preds"], b["bbox_preds"].flip(1)]).to(dev)
    anc = b["anchors"].to(dev)
    sz = [(512, 480), (500, 512)]
    orig = [(375, 352), (1080, 1106)]
    stub = SimpleNamespace(score_thres=0.05, nms_thres=0.5, detections_per_img=100)
    plain = P.process_detections(stub, {"cls_preds": x, "bbox_preds": bb}, [anc] * 2, sz)
    resized = P.process_detections(stub, {"cls_preds": x, "bbox_preds": bb}, [anc] * 2, sz, orig)
    for d, r, s_, o in zip(plain, resized, sz, orig):
        assert torch.equal(r["boxes"], O.resize_boxes(d["boxes"], s_, o))       # transform.postprocess, models.py:271
        assert torch.equal(r["scores"], d["scores"]) and torch.equal(r["labels"], d["labels"])
    for algo in ("auto", "general"):
        h = postprocess_batch_async(x, bb, anc, 0, sz, 0.05, 0.5, 100, algo=algo, original_image_sizes=orig, box_format="xywh")
        res = h.coco_results([11, 22])
        want = []
        for img, r in zip([11, 22], resized):                                   # utils/coco/coco_eval.py:71-93, 159-161
            x0, y0, x1, y1 = r["boxes"].unbind(1)
            xywh = torch.stack((x0, y0, x1 - x0, y1 - y0), dim=1).tolist()
            want.extend({"image_id": img, "category_id": l, "bbox": bx, "score": sc}
                        for l, bx, sc in zip(r["labels"].tolist(), xywh, r["scores"].tolist()))
        assert res == want, algo
    with pytest.raises(ValueError):
        postprocess_batch_async(x, bb, anc, 0, sz, 0.05, 0.5, 100).coco_results([1, 2])


def test_pack_targets_one_launch(P):
    """Row N3: ragged targets packed by rn_pack_targets (no torch.cat, no H2D) == the torch packing; with
    per-image ratios == torchvision's resize_boxes applied to every image's boxes first."""
    import numpy as np
    from pytorch_retinanet_b200.box_utils import PackedTargets
    gen = torch.Generator().manual_seed(9)
    dev = torch.device("cuda")
    counts = [0, 3, 1, 0, 57, 200, 0] + [5] * 140                     # > 128 images -> two launches
    boxes = [S._gt_boxes(gen, c, (600, 900)) if c else torch.zeros((0, 4)) for c in counts]
    labels = [torch.randint(1, 81, (c,), generator=gen) for c in counts]
    bg, lg = [b.to(dev) for b in boxes], [l.to(dev) for l in labels]
    native = PackedTargets(bg, lg, dev)
    assert native.offsets.cpu().tolist() == [0] + list(np.cumsum(counts))
    assert torch.equal(native.boxes[:native.total].cpu(), torch.cat(boxes)) and torch.equal(native.labels[:native.total].cpu(), torch.cat(labels))
    generic = PackedTargets([b.double() for b in bg], lg, dev)          # dtype conversion -> torch path
    assert torch.equal(generic.boxes, native.boxes[:native.total]) and torch.equal(generic.offsets, native.offsets)
    orig = [(480 + 7 * i, 640 + 3 * i) for i in range(len(counts))]
    new = [(800, 1066 + (i % 5)) for i in range(len(counts))]
    ratios = [(float(np.float32(n[0]) / np.float32(o[0])), float(np.float32(n[1]) / np.float32(o[1]))) for o, n in zip(orig, new)]
    scaled = PackedTargets(bg, lg, dev, ratios_hw=ratios)
    want = torch.cat([O.resize_boxes(b, o, n) for b, o, n in zip(boxes, orig, new)])
    assert torch.equal(scaled.boxes[:scaled.total].cpu(), want)
    m = PackedTargets([torch.zeros((0, 4), device=dev)], [torch.zeros((0,), dtype=torch.int64, device=dev)], dev)
    assert m.total == 0 and m.offsets.cpu().tolist() == [0, 0]


def test_randomized_differential(P):
    """Seeded random problems (N 1-4, ragged G incl. 0, C 1-90 incl. odd, A 1-6000, random thresholds):
    matches bit-exact, losses/gradients 1e-5 vs the CPU oracle, detections bit-exact vs the oracle run on
    the same device."""
    from types import SimpleNamespace
    gen = torch.Generator().manual_seed(20260117)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=gen))
    for case in range(22):
        N, C = ri(1, 4), [1, 2, 3, 7, 20, 33, 80, 90][case % 8]
        H, W = 8 * ri(2, 30), 8 * ri(2, 30)
        if case % 5 == 0:
            anc = torch.rand((ri(1, 200), 2), generator=gen) * 100
            anc = torch.cat([anc, anc + 2 + torch.rand((anc.shape[0], 2), generator=gen) * 50], 1)
        else:
            anc = S.default_anchors((H, W))
        A = anc.shape[0]
        targets = []
        for i in range(N):
            G = [0, 1, ri(2, 60)][ri(0, 2)]
            if G:
                src = anc[torch.randint(0, A, (G,), generator=gen)]
                g = src + (torch.rand((G, 4), generator=gen) - 0.5) * (12.0 if i % 2 else 0.0)
                g = torch.stack([g[:, 0], g[:, 1], torch.maximum(g[:, 2], g[:, 0] + 1), torch.maximum(g[:, 3], g[:, 1] + 1)], 1)
            else:
                g = torch.zeros((0, 4))
            targets.append({"boxes": g, "labels": torch.randint(1, C + 1, (G,), generator=gen)})
        cls = torch.randn((N, A, C), generator=gen) * 2.5 - 3.0
        box = torch.randn((N, A, 4), generator=gen) * 0.4
        fg_t, bg_t = (0.5, 0.4) if case % 3 else (0.3 + 0.4 * float(torch.rand(1, generator=gen)), 0.25)
        dev = torch.device("cuda")
        anc_g = anc.to(dev)
        for i in range(N):
            m = P.matcher(anc_g, targets[i]["boxes"].to(dev), fg_t, bg_t).cpu()
            assert torch.equal(m, O.match(anc, targets[i]["boxes"], fg_t, bg_t)), (case, i)
        xo, bo = cls.clone().requires_grad_(True), box.clone().requires_grad_(True)
        want = O.batch_loss(targets, xo, bo, [anc] * N, C)
        tot = want["classification_loss"] + want["regression_loss"]
        if tot.requires_grad:
            tot.backward()
        out, x, b, _ = run_gpu_loss(P, cls, box, targets, [anc_g] * N, C)
        assert rel_close(out["classification_loss"], want["classification_loss"].detach(), LOSS_RTOL, 1e-7), case
        assert rel_close(out["regression_loss"], want["regression_loss"].detach(), LOSS_RTOL, 1e-7), case
        if xo.grad is not None:
            assert rel_close(x.grad, xo.grad, 2e-5, 1e-8), (case, max_rel(x.grad, xo.grad, 1e-6))
            assert rel_close(b.grad, bo.grad if bo.grad is not None else torch.zeros_like(bo), 2e-5, 1e-8), case
        sz = [(H - ri(0, 7), W - ri(0, 7)) for _ in range(N)]
        thr, nms, md = [0.05, 0.3, 0.01][case % 3], [0.5, 0.35, 0.7][case % 3], [100, 7, 300][case % 3]
        wantd = O.postprocess(cls.to(dev), box.to(dev), [anc_g] * N, sz, score_thr=thr, nms_thr=nms, max_det=md)
        for algo in ("auto", "general"):
            got = gpu_detect(P, cls, box, [anc_g] * N, sz, algo=algo, score=thr, nms=nms, max_det=md)
            for i in range(N):
                assert_dets_equal(got[i], wantd[i], exact=True, ctx=f"case {case} {algo} image {i}")


def test_nms_segments_vs_torchvision(P):
    import ctypes
    import torchvision
    from pytorch_retinanet_b200 import _native
    lib = _native.load()
    gen = torch.Generator().manual_seed(2)
    segs, offs = [], [0]
    for n in (0, 1, 5, 300, 1500):
        ctr = torch.rand((n, 2), generator=gen) * 60
        wh = torch.rand((n, 2), generator=gen) * 30 + 1
        bx = torch.cat([ctr, ctr + wh], 1)
        if n >= 5:
            bx[1] = bx[0]                                            # duplicates
        sc = torch.rand(n, generator=gen)
        order = torch.sort(sc, descending=True, stable=True)[1]
        segs.append((bx[order], sc[order]))
        offs.append(offs[-1] + n)
    boxes = torch.cat([s[0] for s in segs]).cuda().contiguous()
    off = torch.tensor(offs, dtype=torch.int32).cuda()
    keep = torch.zeros(boxes.shape[0], dtype=torch.uint8, device="cuda")
    ws = torch.empty(boxes.shape[0] * 16, dtype=torch.uint8, device="cuda")
    for thr in (0.5, 0.3, 0.7000000001):
        rc = lib.rn_nms_segments(boxes.data_ptr(), off.data_ptr(), len(segs), boxes.shape[0], ctypes.c_double(thr),
                                 keep.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
        _native.check(rc, "rn_nms_segments")
        kf = keep.cpu()
        for i, (bx, sc) in enumerate(segs):
            want = torchvision.ops.nms(bx, sc, thr)                  # CPU kernel, input already sorted
            gotk = torch.nonzero(kf[offs[i]:offs[i + 1]]).squeeze(1)
            assert torch.equal(gotk, torch.sort(want)[0]), (thr, i)
            assert CO.nms(bx.numpy(), sc.numpy(), thr).tolist() == want.tolist()


def test_no_cpu_fallback(P):
    from pytorch_retinanet_b200 import _native
    anc = S.default_anchors((64, 64))
    with pytest.raises(_native.NativeError):
        P.matcher(anc, anc[:3])
    with pytest.raises(_native.NativeError):
        P.AnchorGenerator().grid_anchors(S.grid_sizes((64, 64)), torch.device("cpu"))


def test_pack_targets_from_host_single_copy(P):
    """Targets still on the host (the reference's collate_fn output) are packed into one pinned buffer and shipped
    with ONE H2D copy: same packed tensors as the device-side packer, with and without the GT-box resize, and the
    loss accepts them directly."""
    from pytorch_retinanet_b200.box_utils import PackedTargets
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(3)
    boxes = [torch.rand((k, 4), generator=g) * 300 for k in (5, 0, 17, 1)]
    boxes[1] = torch.zeros((0, 4))
    labels = [torch.randint(1, 21, (b.shape[0],), generator=g) for b in boxes]
    ratios = [(1.5, 0.75), (1.0, 1.0), (0.5, 2.0), (1.25, 1.25)]
    for rat in (None, ratios):
        host = PackedTargets(boxes, labels, dev, ratios_hw=rat)
        devp = PackedTargets([b.to(dev) for b in boxes], [l.to(dev) for l in labels], dev, ratios_hw=rat)
        assert hasattr(host, "_staging") and not hasattr(devp, "_staging")
        assert host.total == devp.total == 23 and host.counts == devp.counts
        assert torch.equal(host.offsets, devp.offsets) and host.offsets.dtype == torch.int32
        assert torch.equal(host.boxes, devp.boxes[:23]) and torch.equal(host.labels, devp.labels[:23])
    cfg = S.CONFIGS[1]
    b = S.make_batch(cfg, 9, 2, clustered=True)
    anc = b["anchors"].to(dev)
    L = P.RetinaNetLosses(cfg.num_classes)
    x, bb = b["cls_preds"].to(dev), b["bbox_preds"].to(dev)
    a = L(b["targets"], {"cls_preds": x, "bbox_preds": bb}, [anc] * 2)                        # CPU targets
    c = L(to_cuda_targets(b["targets"]), {"cls_preds": x, "bbox_preds": bb}, [anc] * 2)       # CUDA targets
    for k in a:
        assert torch.equal(a[k], c[k]), k
