"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden.py
The reference ships no tests or golden vectors of its own, so these files — outputs
of the reference's own functions on seeded inputs — are what pins the oracle and the CUDA path.
Large inputs are NOT stored: they are regenerated from seeds by synth_data.py and guarded by a
checksum stored next to the expected outputs.
"""
import hashlib
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import synth_data as S  # noqa: E402
from oracle.ref_shim import load_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ref = load_reference()
torch.set_num_threads(8)


def digest(*tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def ref_postprocess(cls, box, anchors_list, im_szs, score=0.05, nms=0.5, max_det=100):
    stub = SimpleNamespace(score_thres=score, nms_thres=nms, detections_per_img=max_det)
    return ref.models.Retinanet.process_detections(
        stub, {"cls_preds": cls.clone(), "bbox_preds": box.clone()}, anchors_list, im_szs)


def known_answers():
    d = {}
    gen = ref.anchors.AnchorGenerator()
    for i, c in enumerate(gen.cell_anchors):
        d[f"cell_{i}"] = c.numpy()
    d["anchors_512_last"] = torch.cat(gen.grid_anchors(S.grid_sizes((512, 512)), "cpu"))[-1].numpy()
    d["anchors_off05_64x96"] = torch.cat(
        ref.anchors.AnchorGenerator(offset=0.5).grid_anchors(S.grid_sizes((64, 96)), "cpu")).numpy()
    # matcher micro case: tie -> first GT, IoU 0 -> background, IoU exactly 0.5 -> ignore
    anc = torch.tensor([[0., 0, 10, 10], [100, 100, 110, 110], [0, 0, 10, 20]])
    gt = torch.tensor([[0., 0, 10, 10], [0, 0, 10, 10]])
    d["m_anchors"], d["m_gt"] = anc.numpy(), gt.numpy()
    d["m_out"] = ref.box_utils.matcher(anc, gt).numpy()
    d["m_out_empty"] = ref.box_utils.matcher(anc, torch.zeros((0, 4))).numpy()
    # decode quirk
    act, a1 = torch.tensor([[.1, .2, .3, .4]]), torch.tensor([[0., 0, 10, 20]])
    d["dec_act"], d["dec_anchor"] = act.numpy(), a1.numpy()
    d["dec_out"] = ref.box_utils.activ_2_bbox(act.clone(), a1).numpy()
    # encode incl. a degenerate GT (log(1e-8))
    g2 = torch.tensor([[1., 2, 9, 14], [5, 5, 5, 5]])
    a2 = torch.tensor([[0., 0, 10, 20], [0, 0, 10, 10]])
    d["enc_gt"], d["enc_anchor"] = g2.numpy(), a2.numpy()
    d["enc_out"] = ref.box_utils.bbox_2_activ(g2, a2).numpy()
    # 3-anchor loss case
    L = ref.losses.RetinaNetLosses(3)
    cls = torch.tensor([[[.3, -1, 2], [.1, .2, .3], [.5, .5, .5]]], requires_grad=True)
    bb = torch.tensor([[[.1, .2, .3, .4], [0, 0, 0, 0], [0, 0, 0, 0]]], requires_grad=True)
    anc3 = torch.tensor([[0., 0, 10, 10], [100, 100, 110, 110], [0, 0, 10, 20]])
    tg = [{"boxes": torch.tensor([[0., 0, 10, 11]]), "labels": torch.tensor([2])}]
    out = L(tg, {"cls_preds": cls, "bbox_preds": bb}, [anc3])
    (out["classification_loss"] + out["regression_loss"]).backward()
    d["l3_cls"], d["l3_bb"], d["l3_anchors"] = cls.detach().numpy(), bb.detach().numpy(), anc3.numpy()
    d["l3_gt"], d["l3_labels"] = tg[0]["boxes"].numpy(), tg[0]["labels"].numpy()
    d["l3_closs"], d["l3_rloss"] = out["classification_loss"].detach().numpy(), out["regression_loss"].detach().numpy()
    d["l3_gcls"], d["l3_gbb"] = cls.grad.numpy(), bb.grad.numpy()
    # empty-GT image: both losses 0
    out0 = L([{"boxes": torch.zeros((0, 4)), "labels": torch.zeros((0,), dtype=torch.int64)}],
             {"cls_preds": cls.detach(), "bbox_preds": bb.detach()}, [anc3])
    d["l3_empty"] = np.array([float(out0["classification_loss"]), float(out0["regression_loss"])], dtype=np.float32)
    # nms ties / strictness via the reference's own dependency
    import torchvision
    nb = torch.tensor([[0., 0, 10, 10], [0, 0, 10, 10], [0, 0, 10, 20], [50, 50, 60, 60]])
    d["nms_boxes"] = nb.numpy()
    d["nms_scores"] = np.array([.9, .9, .8, .7], dtype=np.float32)
    d["nms_keep"] = torchvision.ops.nms(nb, torch.tensor([.9, .9, .8, .7]), 0.5).numpy()
    np.savez_compressed(os.path.join(OUT, "known_answers.npz"), **d)


def random_small_cases(n_cases=24):
    """Tiny random problems stored WITH their inputs (ragged/empty GT, ties, several C)."""
    d = {}
    g = torch.Generator().manual_seed(20261017)
    for k in range(n_cases):
        A = int(torch.randint(1, 400, (1,), generator=g))
        C = [1, 3, 4, 5, 8, 20][k % 6]
        G = [0, 1, 2, 7, 33][k % 5]
        anchors = torch.rand((A, 2), generator=g) * 100
        anchors = torch.cat([anchors, anchors + 4 + torch.rand((A, 2), generator=g) * 60], 1)
        if G:
            pick = torch.randint(0, A, (G,), generator=g)
            gt = anchors[pick] + (torch.rand((G, 4), generator=g) - 0.5) * (8.0 if k % 2 else 0.0)  # exact copies -> ties
        else:
            gt = torch.zeros((0, 4))
        labels = torch.randint(1, C + 1, (G,), generator=g)
        cls = (torch.randn((1, A, C), generator=g) * 2.0 - 2.0).requires_grad_(True)
        bb = (torch.randn((1, A, 4), generator=g) * 0.3).requires_grad_(True)
        m = ref.box_utils.matcher(anchors, gt)
        L = ref.losses.RetinaNetLosses(C)
        out = L([{"boxes": gt, "labels": labels}], {"cls_preds": cls, "bbox_preds": bb}, [anchors])
        tot = out["classification_loss"] + out["regression_loss"]
        if tot.requires_grad:
            tot.backward()
        gcls = cls.grad if cls.grad is not None else torch.zeros_like(cls)
        gbb = bb.grad if bb.grad is not None else torch.zeros_like(bb)
        det = ref_postprocess(cls.detach(), bb.detach(), [anchors], [(90, 120)], score=0.05, nms=0.5, max_det=20)[0]
        p = f"c{k}_"
        d[p + "anchors"], d[p + "gt"], d[p + "labels"] = anchors.numpy(), gt.numpy(), labels.numpy()
        d[p + "cls"], d[p + "bb"] = cls.detach().numpy(), bb.detach().numpy()
        d[p + "matches"] = m.numpy()
        d[p + "closs"], d[p + "rloss"] = out["classification_loss"].detach().numpy(), out["regression_loss"].detach().numpy()
        d[p + "gcls"], d[p + "gbb"] = gcls.numpy(), gbb.numpy()
        d[p + "det_boxes"], d[p + "det_scores"], d[p + "det_labels"] = (det["boxes"].numpy(), det["scores"].numpy(),
                                                                       det["labels"].numpy())
    d["n_cases"] = np.array(n_cases)
    np.savez_compressed(os.path.join(OUT, "random_small.npz"), **d)


def config_case(cid: int, image_index: int = 0):
    """One full-size image of a BASELINE config: expected outputs + input checksum."""
    cfg = S.CONFIGS[cid]
    b = S.make_batch(cfg, image_index, 1)
    anchors = torch.cat(ref.anchors.AnchorGenerator().grid_anchors(S.grid_sizes(cfg.padded_hw), "cpu"))
    assert torch.equal(anchors, b["anchors"])
    tg = b["targets"]
    cls = b["cls_preds"].clone().requires_grad_(True)
    bb = b["bbox_preds"].clone().requires_grad_(True)
    d = {"input_sha256": np.array(digest(b["cls_preds"], b["bbox_preds"], tg[0]["boxes"], tg[0]["labels"])),
         "anchors_sha256": np.array(digest(anchors)), "num_anchors": np.array(anchors.shape[0])}
    m = ref.box_utils.matcher(anchors, tg[0]["boxes"])
    d["matches"] = m.numpy()
    L = ref.losses.RetinaNetLosses(cfg.num_classes)
    out = L(tg, {"cls_preds": cls, "bbox_preds": bb}, [anchors])
    (out["classification_loss"] + out["regression_loss"]).backward()
    d["closs"], d["rloss"] = out["classification_loss"].detach().numpy(), out["regression_loss"].detach().numpy()
    gi = torch.Generator().manual_seed(7)
    idx = torch.randint(0, cls.numel(), (4096,), generator=gi)
    fgrows = torch.nonzero(m >= 0).squeeze(1)[:256]
    d["g_idx"], d["gcls_at_idx"] = idx.numpy(), cls.grad.reshape(-1)[idx].numpy()
    d["g_fgrows"], d["gcls_fgrows"], d["gbb_fgrows"] = fgrows.numpy(), cls.grad[0, fgrows].numpy(), bb.grad[0, fgrows].numpy()
    d["gcls_abs_sum"] = np.array(float(cls.grad.double().abs().sum()))
    d["gbb_abs_sum"] = np.array(float(bb.grad.double().abs().sum()))
    det = ref_postprocess(b["cls_preds"], b["bbox_preds"], [anchors], b["im_szs"])[0]
    d["det_boxes"], d["det_scores"], d["det_labels"] = det["boxes"].numpy(), det["scores"].numpy(), det["labels"].numpy()
    d["num_candidates"] = np.array(int((torch.sigmoid(b["cls_preds"]) > 0.05).sum()))
    np.savez_compressed(os.path.join(OUT, f"config{cid}_img{image_index}.npz"), **d)
    print(f"config {cid}: A={anchors.shape[0]} fg={int((m >= 0).sum())} closs={float(out['classification_loss']):.6f} "
          f"cands={int(d['num_candidates'])} dets={det['boxes'].shape[0]}")


if __name__ == "__main__":
    known_answers()
    random_small_cases()
    for cid in (1, 2, 5):
        config_case(cid)
    print("golden written to", OUT)
