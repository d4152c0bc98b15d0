"""CPU tests (-m "not gpu"): the oracle (torch restatement + C restatement) against the golden
vectors produced by the unmodified reference, and — when /root/reference is present — against the
reference itself."""
import numpy as np
import pytest
import torch

import synth_data as S
from helpers import config_image, golden, rel_close
from oracle import c_oracle as CO
from oracle import torch_oracle as O
from oracle.ref_shim import reference_available


def test_known_answers_anchors():
    g = golden("known_answers.npz")
    for i in range(5):
        cell = O.cell_anchor_table(O.SIZES[i], O.RATIOS)
        assert np.array_equal(cell.numpy(), g[f"cell_{i}"])
        assert np.array_equal(CO.cell_anchors(O.SIZES[i], O.RATIOS), g[f"cell_{i}"])
    assert np.allclose(g["cell_0"][0], [-22.6274, -11.3137, 22.6274, 11.3137], atol=1e-4)
    a = O.image_anchors(O.fpn_grid_sizes(512, 512))
    assert a.shape[0] == 49104 and np.array_equal(a[-1].numpy(), g["anchors_512_last"])
    a05 = O.image_anchors(O.fpn_grid_sizes(64, 96), offset=0.5)
    assert np.array_equal(a05.numpy(), g["anchors_off05_64x96"])
    c05 = CO.image_anchors(O.fpn_grid_sizes(64, 96), O.STRIDES, O.SIZES, [O.RATIOS] * 5, offset=0.5)
    assert np.array_equal(c05, g["anchors_off05_64x96"])
    assert [S.num_anchors(S.CONFIGS[c].padded_hw) for c in (1, 2, 5)] == [49104, 201600, 196416]


def test_known_answers_matcher_and_coding():
    g = golden("known_answers.npz")
    anc, gt = torch.from_numpy(g["m_anchors"]), torch.from_numpy(g["m_gt"])
    assert g["m_out"].tolist() == [0, -1, -2]
    assert O.match(anc, gt).tolist() == [0, -1, -2]
    assert CO.match(anc, gt).tolist() == [0, -1, -2]
    assert O.match(anc, torch.zeros((0, 4))).tolist() == g["m_out_empty"].tolist() == [-2, -2, -2]
    assert CO.match(anc, np.zeros((0, 4))).tolist() == [-2, -2, -2]
    dec = O.decode(torch.from_numpy(g["dec_act"]), torch.from_numpy(g["dec_anchor"]))
    assert np.array_equal(dec.numpy(), g["dec_out"])
    assert np.allclose(g["dec_out"], [[0.4741, 1.7860, 11.5259, 26.2140]], atol=1e-4)       # the exp(dx,dy) quirk
    assert np.allclose(CO.decode(g["dec_act"], g["dec_anchor"]), g["dec_out"], rtol=1e-6)
    enc = O.encode(torch.from_numpy(g["enc_gt"]), torch.from_numpy(g["enc_anchor"]))
    assert np.array_equal(enc.numpy(), g["enc_out"])
    assert abs(float(g["enc_out"][1, 2]) - (-18.4207)) < 1e-3                               # log(1e-8)
    assert np.allclose(CO.encode(g["enc_gt"], g["enc_anchor"]), g["enc_out"], rtol=1e-6)


def test_known_answers_loss_and_nms():
    g = golden("known_answers.npz")
    cls = torch.from_numpy(g["l3_cls"]).requires_grad_(True)
    bb = torch.from_numpy(g["l3_bb"]).requires_grad_(True)
    anc = torch.from_numpy(g["l3_anchors"])
    tg = [{"boxes": torch.from_numpy(g["l3_gt"]), "labels": torch.from_numpy(g["l3_labels"])}]
    out = O.batch_loss(tg, cls, bb, [anc], 3)
    (out["classification_loss"] + out["regression_loss"]).backward()
    assert np.array_equal(out["classification_loss"].detach().numpy(), g["l3_closs"])
    assert np.array_equal(out["regression_loss"].detach().numpy(), g["l3_rloss"])
    assert abs(float(g["l3_closs"]) - 1.1412) < 1e-4 and abs(float(g["l3_rloss"]) - 0.6888) < 1e-4
    assert np.array_equal(cls.grad.numpy(), g["l3_gcls"]) and np.array_equal(bb.grad.numpy(), g["l3_gbb"])
    c, r, F, m = CO.image_loss(g["l3_anchors"], g["l3_cls"][0], g["l3_bb"][0], g["l3_gt"], g["l3_labels"])
    assert rel_close(c, g["l3_closs"], 1e-6) and rel_close(r, g["l3_rloss"], 1e-6) and F == 2
    e = O.batch_loss([{"boxes": torch.zeros((0, 4)), "labels": torch.zeros((0,), dtype=torch.int64)}],
                     cls.detach(), bb.detach(), [anc], 3)
    assert float(e["classification_loss"]) == 0.0 == float(g["l3_empty"][0])
    keep = g["nms_keep"].tolist()
    assert keep == [0, 2, 3]       # equal scores -> lower index kept; IoU exactly 0.5 is NOT suppressed (strict >)
    assert CO.nms(g["nms_boxes"], g["nms_scores"], 0.5).tolist() == keep
    assert O.nms_keep(torch.from_numpy(g["nms_boxes"]), torch.from_numpy(g["nms_scores"]), 0.5).tolist() == keep


def test_random_small_cases_both_oracles():
    g = golden("random_small.npz")
    for k in range(int(g["n_cases"])):
        p = f"c{k}_"
        anc, gt, lab = torch.from_numpy(g[p + "anchors"]), torch.from_numpy(g[p + "gt"]), torch.from_numpy(g[p + "labels"])
        cls = torch.from_numpy(g[p + "cls"]).requires_grad_(True)
        bb = torch.from_numpy(g[p + "bb"]).requires_grad_(True)
        C = cls.shape[-1]
        assert np.array_equal(O.match(anc, gt).numpy(), g[p + "matches"]), k
        assert np.array_equal(CO.match(anc, gt), g[p + "matches"]), k
        out = O.batch_loss([{"boxes": gt, "labels": lab}], cls, bb, [anc], C)
        tot = out["classification_loss"] + out["regression_loss"]
        if tot.requires_grad:
            tot.backward()
        assert np.array_equal(out["classification_loss"].detach().numpy(), g[p + "closs"]), k
        assert np.array_equal(out["regression_loss"].detach().numpy(), g[p + "rloss"]), k
        if cls.grad is not None:
            assert np.array_equal(cls.grad.numpy(), g[p + "gcls"]), k
        c, r, F, _ = CO.image_loss(anc, g[p + "cls"][0], g[p + "bb"][0], gt, lab)
        assert rel_close(c, g[p + "closs"], 1e-5, 1e-7) and rel_close(r, g[p + "rloss"], 1e-5, 1e-7), k
        det = O.postprocess(cls.detach(), bb.detach(), [anc], [(90, 120)], max_det=20)[0]
        assert np.array_equal(det["labels"].numpy(), g[p + "det_labels"]), k
        assert np.array_equal(det["scores"].numpy(), g[p + "det_scores"]), k
        assert np.array_equal(det["boxes"].numpy(), g[p + "det_boxes"]), k
        ob, os_, ol = CO.postprocess_image(g[p + "cls"][0], g[p + "bb"][0], anc, (90, 120), max_det=20)
        assert ob.shape == g[p + "det_boxes"].shape, k
        assert np.allclose(os_, g[p + "det_scores"], rtol=1e-5) and np.allclose(ob, g[p + "det_boxes"], rtol=1e-5, atol=1e-4), k


@pytest.mark.parametrize("cid", [1, 2, 5])
def test_full_size_configs(cid):
    b, g = config_image(cid)
    anc, t = b["anchors"], b["targets"][0]
    m = O.match(anc, t["boxes"], chunk=32768)
    assert np.array_equal(m.numpy(), g["matches"])
    assert np.array_equal(CO.match(anc, t["boxes"]), g["matches"])
    if cid == 5:
        return   # loss/detections at G=500 are covered on the GPU box; keep the CPU suite short
    cls = b["cls_preds"].clone().requires_grad_(True)
    bb = b["bbox_preds"].clone().requires_grad_(True)
    out = O.batch_loss(b["targets"], cls, bb, [anc], b["config"].num_classes, chunk=32768)
    (out["classification_loss"] + out["regression_loss"]).backward()
    assert np.array_equal(out["classification_loss"].detach().numpy(), g["closs"])
    assert np.array_equal(out["regression_loss"].detach().numpy(), g["rloss"])
    assert np.array_equal(cls.grad.reshape(-1)[torch.from_numpy(g["g_idx"])].numpy(), g["gcls_at_idx"])
    assert np.array_equal(bb.grad[0, torch.from_numpy(g["g_fgrows"])].numpy(), g["gbb_fgrows"])
    c, r, F, _ = CO.image_loss(anc, b["cls_preds"][0], b["bbox_preds"][0], t["boxes"], t["labels"])
    assert rel_close(c, g["closs"], 1e-5) and rel_close(r, g["rloss"], 1e-5)
    ob, os_, ol = CO.postprocess_image(b["cls_preds"][0], b["bbox_preds"][0], anc, b["im_szs"][0])
    assert np.array_equal(ol, g["det_labels"])
    assert np.allclose(os_, g["det_scores"], rtol=1e-5) and np.allclose(ob, g["det_boxes"], rtol=1e-5, atol=1e-4)
    if cid == 1:
        det = O.postprocess(b["cls_preds"], b["bbox_preds"], [anc], b["im_szs"])[0]
        assert np.array_equal(det["boxes"].numpy(), g["det_boxes"]) and np.array_equal(det["labels"].numpy(), g["det_labels"])


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
def test_oracle_matches_live_reference():
    """Differential test against the imported reference on fresh random inputs (build container only)."""
    from types import SimpleNamespace
    from oracle.ref_shim import load_reference
    ref = load_reference()
    gen = torch.Generator().manual_seed(99)
    anc = torch.cat(ref.anchors.AnchorGenerator().grid_anchors(S.grid_sizes((128, 160)), "cpu"))
    assert torch.equal(anc, O.image_anchors(O.fpn_grid_sizes(128, 160)))
    for G in (0, 3, 40):
        gt = S._gt_boxes(gen, G, (128, 160)) if G else torch.zeros((0, 4))
        lab = torch.randint(1, 8, (G,), generator=gen)
        cls = torch.randn((2, anc.shape[0], 7), generator=gen) - 3
        bb = torch.randn((2, anc.shape[0], 4), generator=gen) * 0.2
        tg = [{"boxes": gt, "labels": lab}] * 2
        assert torch.equal(ref.box_utils.matcher(anc, gt), O.match(anc, gt))
        lr = ref.losses.RetinaNetLosses(7)(tg, {"cls_preds": cls, "bbox_preds": bb}, [anc, anc])
        lo = O.batch_loss(tg, cls, bb, [anc, anc], 7)
        assert all(torch.equal(lr[k], lo[k]) for k in lr)
        stub = SimpleNamespace(score_thres=0.05, nms_thres=0.5, detections_per_img=100)
        dr = ref.models.Retinanet.process_detections(stub, {"cls_preds": cls.clone(), "bbox_preds": bb.clone()},
                                                     [anc, anc], [(128, 150)] * 2)
        do = O.postprocess(cls, bb, [anc, anc], [(128, 150)] * 2, stable_final_sort=False)
        for a, b_ in zip(dr, do):
            assert torch.equal(a["scores"], b_["scores"]) and torch.equal(a["labels"], b_["labels"])
            assert torch.equal(a["boxes"], b_["boxes"])


@pytest.mark.skipif(not reference_available(), reason="no reference tree (baseline/_ref, /root/reference)")
def test_label_zero_and_separate_backward_match_live_reference():
    """Pins two behaviours the GPU tests check against the oracle: (1) a label 0 is the reference's class id 0 — all-zero
    class targets after the [:,1:] slice, regression term and foreground count kept (losses.py:96-103); a label > C makes
    the reference's one_hot raise; (2) the two losses can be back-propagated one after the other (retain_graph), the
    class loss sends no gradient to bbox_preds."""
    from oracle.ref_shim import load_reference
    ref = load_reference()
    gen = torch.Generator().manual_seed(5)
    anc = O.image_anchors(O.fpn_grid_sizes(128, 160))
    gt = S._gt_boxes(gen, 12, (128, 160))
    lab = torch.randint(1, 8, (12,), generator=gen)
    lab[::2] = 0
    cls = (torch.randn((1, anc.shape[0], 7), generator=gen) - 3).requires_grad_(True)
    bb = (torch.randn((1, anc.shape[0], 4), generator=gen) * 0.2).requires_grad_(True)
    tg = [{"boxes": gt, "labels": lab}]
    lr = ref.losses.RetinaNetLosses(7)(tg, {"cls_preds": cls, "bbox_preds": bb}, [anc])
    co = cls.detach().clone().requires_grad_(True)
    bo = bb.detach().clone().requires_grad_(True)
    lo = O.batch_loss(tg, co, bo, [anc], 7)
    assert all(torch.equal(lr[k], lo[k]) for k in lr) and float(lr["regression_loss"]) > 0
    lr["classification_loss"].backward(retain_graph=True)
    lo["classification_loss"].backward(retain_graph=True)
    assert bb.grad is None and bo.grad is None
    lr["regression_loss"].backward()
    lo["regression_loss"].backward()
    assert torch.equal(cls.grad, co.grad) and torch.equal(bb.grad, bo.grad)
    lab_bad = torch.full_like(lab, 9)                               # > C: the reference cannot build the one-hot rows
    with pytest.raises(Exception):
        ref.losses.RetinaNetLosses(7)([{"boxes": gt, "labels": lab_bad}], {"cls_preds": cls.detach(), "bbox_preds": bb.detach()}, [anc])


def test_degenerate_inputs_both_oracles_agree():
    """Adversarial inputs for the two restatements (torch op-for-op and plain C): exact-threshold IoUs, ties, duplicate
    and zero-area boxes, NaN / inf coordinates, equal NMS scores.  The torch oracle is the reference's own op
    sequence (bit-identical to it, see test_oracle_matches_live_reference); the C oracle must agree with it."""
    g = torch.Generator().manual_seed(11)
    base = torch.tensor([[0., 0., 10., 10.], [100., 100., 110., 110.], [0., 0., 10., 20.], [5., 5., 5., 5.],
                         [0., 0., 10., 12.5], [2., 2., 8., 8.]])
    extra = torch.rand((40, 2), generator=g) * 50
    wh = torch.rand((40, 2), generator=g) * 30 + 1
    anchors = torch.cat([base, torch.cat([extra, extra + wh], 1), base[:2]])            # duplicates at the end
    cases = [
        torch.tensor([[0., 0., 10., 10.], [0., 0., 10., 10.]]),                         # tie -> first GT
        torch.tensor([[0., 0., 10., 8.], [0., 0., 10., 25.]]),                          # IoU 0.8 / exactly 0.4 with anchor 0
        torch.tensor([[3., 3., 3., 3.], [0., 0., 10., 10.]]),                           # zero-area GT
        torch.tensor([[float("nan"), 0., 10., 10.], [0., 0., 10., 10.]]),               # NaN propagates through max
        torch.tensor([[0., 0., float("inf"), 10.], [20., 20., 40., 45.]]),              # inf extent
        torch.tensor([[10., 10., 0., 0.], [0., 0., 10., 10.]]),                         # inverted box (negative extent)
        torch.zeros((0, 4)),                                                            # no GT -> all ignore
        torch.cat([extra[:25], extra[:25] + wh[:25]], 1),                               # GT identical to some anchors
    ]
    for k, gt in enumerate(cases):
        want = O.match(anchors, gt).numpy()
        got = CO.match(anchors, gt)
        assert np.array_equal(got, want), (k, got[:12], want[:12])
    # NMS: equal scores keep the lower index, strict threshold, many exact duplicates
    boxes = torch.cat([base[[0, 0, 2, 4, 5]], torch.cat([extra[:20], extra[:20] + wh[:20]], 1)])
    for scores in (torch.ones(boxes.shape[0]), torch.linspace(1, 0.1, boxes.shape[0]).round(decimals=1),
                   torch.rand(boxes.shape[0], generator=g)):
        for thr in (0.5, 0.4, 0.0, 1.0):
            want = O.nms_keep(boxes, scores, thr).numpy()
            got = CO.nms(boxes, scores, thr)
            assert np.array_equal(np.asarray(got), want), (thr, got, want)
            try:
                import torchvision
                assert np.array_equal(torchvision.ops.nms(boxes, scores, thr).numpy(), want), thr
            except ImportError:
                pass
