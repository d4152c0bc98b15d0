"""GPU tests (-m gpu) of ``HotPathGraph``: the CUDA graph of the path must give bit-identical results to
the drop-in calls (same kernels, same arguments), which the parity tests pin against the oracle; it is also
checked against the oracle directly at the north-star tolerances."""
from types import SimpleNamespace

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


def _dropin(P, cfg, x, bb, anc, tg, im_szs):
    n = x.shape[0]
    xg, bg = x.clone().requires_grad_(True), bb.clone().requires_grad_(True)
    L = P.RetinaNetLosses(cfg.num_classes)
    out = L(tg, {"cls_preds": xg, "bbox_preds": bg}, [anc] * n)
    (out["classification_loss"] + out["regression_loss"]).backward()
    stub = SimpleNamespace(score_thres=0.05, nms_thres=0.5, detections_per_img=100)
    dets = P.process_detections(stub, {"cls_preds": x, "bbox_preds": bb}, [anc] * n, im_szs)
    return out, xg.grad, bg.grad, L.last_per_image.clone(), dets


def _assert_same_dets(a, b):
    assert len(a) == len(b)
    for d, r in zip(a, b):
        for k in ("boxes", "scores", "labels"):
            assert d[k].dtype == r[k].dtype and d[k].shape == r[k].shape, k
            assert torch.equal(d[k], r[k]), k


@pytest.mark.parametrize("mode", ["fused", "concurrent", "serial"])
@pytest.mark.parametrize("cid,n_img", [(1, 3), (2, 4)])
def test_graph_step_equals_dropin_calls(P, cid, n_img, mode):
    """fused = rn_train_detect (one pass over the logits for both halves); concurrent / serial = rn_train_loss and
    rn_postprocess as two branches / back to back."""
    concurrent, fused = mode == "concurrent", mode == "fused"
    from pytorch_retinanet_b200.graphs import HotPathGraph
    cfg = S.CONFIGS[cid]
    dev = torch.device("cuda")
    b = S.make_batch(cfg, 7, n_img, clustered=True)
    anc = b["anchors"].to(dev)
    x, bb = b["cls_preds"].to(dev), b["bbox_preds"].to(dev)
    tg = to_cuda_targets(b["targets"])
    g = HotPathGraph(cfg.num_classes, x, bb, anc, b["im_szs"], max_targets=2048, concurrent=concurrent, fused=fused)
    assert g.fused == fused
    for rep in range(3):                      # replays; the 2nd with other targets and inputs written into the static buffers
        if rep == 1:
            b2 = S.make_batch(cfg, 100, n_img, clustered=True)
            x.copy_(b2["cls_preds"])
            bb.copy_(b2["bbox_preds"])
            tg = to_cuda_targets(b2["targets"])
            tg[1] = {"boxes": torch.zeros((0, 4), device=dev), "labels": torch.zeros((0,), dtype=torch.int64, device=dev)}
        res = g.step(tg)
        want_out, want_gx, want_gb, want_img, want_dets = _dropin(P, cfg, x, bb, anc, tg, b["im_szs"])
        got = res.losses
        for k in want_out:
            assert torch.equal(got[k], want_out[k].detach()), (k, float(got[k]), float(want_out[k]))
        assert torch.equal(res.per_image, want_img)
        gx, gb = res.grads
        assert torch.equal(gx, want_gx) and torch.equal(gb, want_gb)
        _assert_same_dets(res.detections(), want_dets)
    # and against the oracle (CPU restatement of the reference), north-star tolerances
    xo, bo = x.cpu().requires_grad_(True), bb.cpu().requires_grad_(True)
    cpu_t = [{k: v.cpu() for k, v in t.items()} for t in tg]
    want = O.batch_loss(cpu_t, xo, bo, [b["anchors"]] * n_img, cfg.num_classes)
    (want["classification_loss"] + want["regression_loss"]).backward()
    for k in want:
        assert rel_close(res.losses[k], want[k].detach(), 1e-5), k
    assert torch.allclose(res.grads[0].cpu(), xo.grad, rtol=2e-5, atol=1e-9)
    assert torch.allclose(res.grads[1].cpu(), bo.grad, rtol=2e-5, atol=1e-9)


def test_graph_train_only_and_detect_only(P):
    from pytorch_retinanet_b200.graphs import HotPathGraph
    cfg = S.CONFIGS[1]
    dev = torch.device("cuda")
    b = S.make_batch(cfg, 3, 2, clustered=True)
    anc = b["anchors"].to(dev)
    x, bb = b["cls_preds"].to(dev), b["bbox_preds"].to(dev)
    tg = to_cuda_targets(b["targets"])
    want_out, want_gx, _, _, want_dets = _dropin(P, cfg, x, bb, anc, tg, b["im_szs"])
    gt = HotPathGraph(cfg.num_classes, x, bb, anc, train=True, detect=False)
    r = gt.step(tg)
    assert torch.equal(r.losses["classification_loss"], want_out["classification_loss"].detach())
    assert torch.equal(r.grads[0], want_gx)
    gt.grad_cls_preds.zero_()
    r = gt.step(b["targets"])                 # host-resident targets: uploaded with the pinned single-buffer route
    assert torch.equal(r.losses["classification_loss"], want_out["classification_loss"].detach())
    assert torch.equal(r.losses["regression_loss"], want_out["regression_loss"].detach())
    assert torch.equal(r.grads[0], want_gx)
    with pytest.raises(RuntimeError):
        r.detections()
    gd = HotPathGraph(cfg.num_classes, x, bb, anc, b["im_szs"], train=False, detect=True)
    _assert_same_dets(gd.step().detections(), want_dets)
    # a tiny candidate pool overflows: the result must come from the eager re-run and still be identical
    gs = HotPathGraph(cfg.num_classes, x, bb, anc, b["im_szs"], train=False, detect=True, cand_capacity=8)
    _assert_same_dets(gs.step().detections(), want_dets)
    with pytest.raises(ValueError):
        HotPathGraph(cfg.num_classes, x, bb, anc, train=True, detect=False, max_targets=4).step(tg)
    from pytorch_retinanet_b200._native import NativeError
    with pytest.raises(NativeError):
        HotPathGraph(cfg.num_classes, x.cpu(), bb.cpu(), anc.cpu(), train=True, detect=False)


@pytest.mark.parametrize("case", ["cfg1", "cfg2", "crowd_tiles", "odd_A"])
def test_train_loss_call_equals_match_then_loss(P, case):
    """rn_train_loss (one C call) must equal rn_match followed by rn_loss called separately — losses, per-image
    values, gradients, codes — and be bit-reproducible run to run; covers G > GT_TILE (two shared-memory tiles),
    A not a multiple of the CTA span, an image without GT, generic gamma."""
    from pytorch_retinanet_b200 import _native
    from pytorch_retinanet_b200.box_utils import _REG_WEIGHTS_C, PackedTargets, match_batch
    from pytorch_retinanet_b200.losses import fused_loss_forward
    lib = _native.load()
    dev = torch.device("cuda")
    if case == "cfg1":
        cfg, n_img = S.CONFIGS[1], 3
        b = S.make_batch(cfg, 11, n_img, clustered=True)
    elif case == "cfg2":
        cfg, n_img = S.CONFIGS[2], 5
        b = S.make_batch(cfg, 11, n_img, clustered=True)
    elif case == "crowd_tiles":               # G = 500 and one image with G > GT_TILE (512): two shared-memory tiles
        cfg, n_img = S.CONFIGS[5], 3
        b = S.make_batch(cfg, 2, n_img, clustered=True)
        t0, t1 = b["targets"][0], b["targets"][1]
        b["targets"][0] = {"boxes": torch.cat([t0["boxes"], t1["boxes"][:200]]), "labels": torch.cat([t0["labels"], t1["labels"][:200]])}
    else:                                      # A not a multiple of the CTA span, C = 20 (vec4), an image without GT
        cfg, n_img = S.CONFIGS[1], 4
        b = S.make_batch(cfg, 30, n_img, clustered=True)
        keep = 49104 - 77
        b["anchors"] = b["anchors"][:keep].contiguous()
        b["cls_preds"] = b["cls_preds"][:, :keep].contiguous()
        b["bbox_preds"] = b["bbox_preds"][:, :keep].contiguous()
        b["targets"][2] = {"boxes": torch.zeros((0, 4)), "labels": torch.zeros((0,), dtype=torch.int64)}
    anc = b["anchors"].to(dev)
    x, bb = b["cls_preds"].to(dev), b["bbox_preds"].to(dev)
    tg = to_cuda_targets(b["targets"])
    packed = PackedTargets([t["boxes"] for t in tg], [t["labels"] for t in tg], dev)
    N, A, C = x.shape

    def separate(gamma, want_grad):
        _, codes, fg = match_batch(anc, 0, packed, A, 0.5, 0.4, False, True)
        total = torch.empty((4,), device=dev)
        image = torch.empty((N, 3), device=dev)
        gl = torch.empty_like(x) if want_grad else None
        gb = torch.empty_like(bb) if want_grad else None
        nb = lib.rn_loss_workspace_bytes(N, A, C)
        ws = torch.empty((nb,), dtype=torch.uint8, device=dev)
        rc = lib.rn_loss(x.data_ptr(), bb.data_ptr(), anc.data_ptr(), 0, packed.boxes.data_ptr(), packed.offsets.data_ptr(),
                         codes.data_ptr(), fg.data_ptr(), N, A, C, 0.25, gamma, 0.1, _REG_WEIGHTS_C, float(n_img), image.data_ptr(),
                         total.data_ptr(), None if gl is None else gl.data_ptr(), None if gb is None else gb.data_ptr(),
                         ws.data_ptr(), nb, torch.cuda.current_stream().cuda_stream, None)
        _native.check(rc, "rn_loss")
        return total, image, gl, gb, codes

    for gamma in (2.0, 1.5):
        ref = separate(gamma, True)
        for rep in range(3):
            got = fused_loss_forward(x, bb, anc, 0, packed, 0.25, gamma, 0.1, 0.5, 0.4, float(n_img), True)
            for g, r, name in zip(got, ref, ("total", "per_image", "grad_logits", "grad_bbox", "codes")):
                assert torch.equal(g, r), (case, gamma, rep, name)
    got_ng = fused_loss_forward(x, bb, anc, 0, packed, 0.25, 2.0, 0.1, 0.5, 0.4, float(n_img), False)
    ref_ng = separate(2.0, False)
    assert torch.equal(got_ng[0], ref_ng[0]) and torch.equal(got_ng[1], ref_ng[1]) and got_ng[2] is None


def test_graph_on_raw_level_outputs(P):
    """HotPathGraph on the head's raw per-level conv outputs (row N1) equals the drop-in calls on the same lists
    (bit-identical) and the [N,A,C] path (losses 1e-6, detections bit-identical)."""
    from pytorch_retinanet_b200.graphs import HotPathGraph
    cfg = S.CONFIGS[1]
    dev = torch.device("cuda")
    n_img = 3
    b = S.make_batch(cfg, 21, n_img, clustered=True)
    anc = b["anchors"].to(dev)
    tg = to_cuda_targets(b["targets"])
    xs = [t.to(dev) for t in S.nac_to_levels(b["cls_preds"], cfg.padded_hw)]
    bs = [t.to(dev) for t in S.nac_to_levels(b["bbox_preds"], cfg.padded_hw)]
    g = HotPathGraph(cfg.num_classes, xs, bs, anc, b["im_szs"])
    res = g.step(tg)
    L = P.RetinaNetLosses(cfg.num_classes)
    xg, bg = [t.clone().requires_grad_(True) for t in xs], [t.clone().requires_grad_(True) for t in bs]
    out = L(tg, {"cls_levels": xg, "bbox_levels": bg}, [anc] * n_img)
    (out["classification_loss"] + out["regression_loss"]).backward()
    for k in out:
        assert torch.equal(res.losses[k], out[k].detach()), k
    for got, want in zip(res.grads[0] + res.grads[1], xg + bg):
        assert torch.equal(got, want.grad)
    stub = SimpleNamespace(score_thres=0.05, nms_thres=0.5, detections_per_img=100, num_classes=cfg.num_classes)
    dets = P.process_detections(stub, {"cls_levels": xs, "bbox_levels": bs}, [anc] * n_img, b["im_szs"])
    _assert_same_dets(res.detections(), dets)
    want_out, _, _, _, want_dets = _dropin(P, cfg, b["cls_preds"].to(dev), b["bbox_preds"].to(dev), anc, tg, b["im_szs"])
    for k in want_out:
        assert rel_close(res.losses[k], want_out[k].detach(), 1e-6), k
    _assert_same_dets(res.detections(), want_dets)


def test_graph_pre_nms_topk_extension(P):
    """pre_nms_topk (per-level top-k before NMS, an opt-in extension) through the graph equals the drop-in call."""
    from pytorch_retinanet_b200.detections import postprocess_batch, slice_detections
    from pytorch_retinanet_b200.graphs import HotPathGraph
    cfg = S.CONFIGS[1]
    dev = torch.device("cuda")
    b = S.make_batch(cfg, 5, 2, clustered=True)
    anc, x, bb = b["anchors"].to(dev), b["cls_preds"].to(dev), b["bbox_preds"].to(dev)
    offs = [0]
    for h, w in S.grid_sizes(cfg.padded_hw):
        offs.append(offs[-1] + 9 * h * w)
    want = slice_detections(*postprocess_batch(x, bb, anc, 0, b["im_szs"], 0.05, 0.5, 100, pre_nms_topk=50, level_offsets=offs))
    plain = slice_detections(*postprocess_batch(x, bb, anc, 0, b["im_szs"], 0.05, 0.5, 100))
    g = HotPathGraph(cfg.num_classes, x, bb, anc, b["im_szs"], train=False, pre_nms_topk=50, level_offsets=offs)
    got = g.step().detections()
    _assert_same_dets(got, want)
    assert any(not torch.equal(a["scores"], p["scores"]) for a, p in zip(got, plain) if a["scores"].shape == p["scores"].shape) or \
        any(a["scores"].shape != p["scores"].shape for a, p in zip(got, plain))      # the filter is active
    xs = [t.to(dev) for t in S.nac_to_levels(b["cls_preds"], cfg.padded_hw)]
    bs = [t.to(dev) for t in S.nac_to_levels(b["bbox_preds"], cfg.padded_hw)]
    gl = HotPathGraph(cfg.num_classes, xs, bs, anc, b["im_szs"], train=False, pre_nms_topk=50)
    _assert_same_dets(gl.step().detections(), want)


def test_dropin_calls_in_graph_mode(P):
    """``RetinaNetLosses(graph=True)`` / ``process_detections`` with ``rn_graph`` replay cached CUDA graphs keyed on the
    inputs' addresses: bit-identical to the eager drop-in calls, step after step, with inputs that move, change shape
    and change content; the reference's call pattern (one backward per forward) is what the mode supports."""
    cfg = S.CONFIGS[1]
    dev = torch.device("cuda")
    anc = S.default_anchors(cfg.padded_hw).to(dev)
    Lg = P.RetinaNetLosses(cfg.num_classes, graph=True)
    stub_g = SimpleNamespace(score_thres=0.05, nms_thres=0.5, detections_per_img=100, rn_graph=True)
    bufs = {}
    for step, (first, n_img) in enumerate([(7, 3), (40, 3), (80, 2), (7, 3), (41, 3)]):
        b = S.make_batch(cfg, first, n_img, clustered=True)
        tg = to_cuda_targets(b["targets"])
        if step == 1:                                               # one image without ground truth
            tg[1] = {"boxes": torch.zeros((0, 4), device=dev), "labels": torch.zeros((0,), dtype=torch.int64, device=dev)}
        # the SAME device buffers per shape (what the caching allocator gives a model in steady state), new contents
        x, bb = bufs.setdefault(n_img, (torch.empty_like(b["cls_preds"], device=dev), torch.empty_like(b["bbox_preds"], device=dev)))
        x.copy_(b["cls_preds"])
        bb.copy_(b["bbox_preds"])
        want_out, want_gx, want_gb, want_img, want_dets = _dropin(P, cfg, x, bb, anc, tg, b["im_szs"])
        xg, bg = x.detach().requires_grad_(True), bb.detach().requires_grad_(True)     # same storage, fresh autograd leaves
        out = Lg(tg, {"cls_preds": xg, "bbox_preds": bg}, [anc] * n_img)
        if step % 2:
            (out["classification_loss"] + out["regression_loss"]).backward()
            got_gx, got_gb = xg.grad, bg.grad
        else:       # what a model's backward does: the gradients are passed on, not accumulated into leaves
            got_gx, got_gb = torch.autograd.grad(out["classification_loss"] + out["regression_loss"], (xg, bg))
        assert torch.equal(out["classification_loss"].detach(), want_out["classification_loss"].detach()), step
        assert torch.equal(out["regression_loss"].detach(), want_out["regression_loss"].detach()), step
        assert torch.equal(got_gx, want_gx) and torch.equal(got_gb, want_gb), step
        assert torch.equal(Lg.last_per_image, want_img), step
        outputs = {"cls_preds": x, "bbox_preds": bb}
        dets = P.process_detections(stub_g, outputs, [anc] * n_img, b["im_szs"])
        assert outputs == {}
        _assert_same_dets(dets, want_dets)
    assert len(Lg._graphs) == 2 and len(stub_g._rn_det_graphs) == 2    # one graph per (addresses, shape)
    # a second backward through the same forward is refused loudly in this mode
    xg = x.detach().requires_grad_(True)
    out = Lg(tg, {"cls_preds": xg, "bbox_preds": bb}, [anc] * n_img)
    out["classification_loss"].backward(retain_graph=True)
    with pytest.raises(RuntimeError):
        out["classification_loss"].backward()
    # scaled grad_output
    xg = x.detach().requires_grad_(True)
    out = Lg(tg, {"cls_preds": xg, "bbox_preds": bb}, [anc] * n_img)
    (2.0 * out["classification_loss"]).backward()
    assert torch.equal(xg.grad, 2.0 * want_gx)


def test_pipeline_of_split_graphs_equals_single_graph(P):
    """HotPathPipeline (front and tail of the fused step as two graphs on two streams, two input buffers) gives the
    single graph's results, step after step, with results consumed one step late."""
    from pytorch_retinanet_b200.graphs import HotPathGraph, HotPathPipeline
    cfg = S.CONFIGS[1]
    dev = torch.device("cuda")
    n_img = 3
    anc = S.default_anchors(cfg.padded_hw).to(dev)
    bufs = [(torch.empty((n_img, anc.shape[0], cfg.num_classes), device=dev), torch.empty((n_img, anc.shape[0], 4), device=dev))
            for _ in range(2)]
    ref_x, ref_b = torch.empty_like(bufs[0][0]), torch.empty_like(bufs[0][1])
    im_szs = [cfg.im_hw] * n_img
    pipe = HotPathPipeline(cfg.num_classes, bufs, anc, im_szs, max_targets=1024)
    single = HotPathGraph(cfg.num_classes, ref_x, ref_b, anc, im_szs, max_targets=1024, fused=False)
    pending, want = [], []
    for step in range(7):
        b = S.make_batch(cfg, 20 * step, n_img, clustered=True)
        tg = to_cuda_targets(b["targets"])
        x, bb = bufs[step % 2]
        x.copy_(b["cls_preds"])
        bb.copy_(b["bbox_preds"])
        ref_x.copy_(b["cls_preds"])
        ref_b.copy_(b["bbox_preds"])
        r = single.step(tg)
        want.append((r.losses["classification_loss"].clone(), r.losses["regression_loss"].clone(), r.grads[0].clone(),
                     r.grads[1].clone(), [{k: v.clone() for k, v in d.items()} for d in r.detections()]))
        pending.append(pipe.step(tg))
        if len(pending) > 1:                       # consume step-1 after launching this step
            got, w = pending.pop(0), want.pop(0)
            assert torch.equal(got.losses["classification_loss"], w[0]) and torch.equal(got.losses["regression_loss"], w[1])
            # (the gradients of that step are still intact: the other buffer set is in use now)
            assert torch.equal(got.grads[0], w[2]) and torch.equal(got.grads[1], w[3])
            _assert_same_dets(got.detections(), w[4])
    got, w = pending.pop(0), want.pop(0)
    assert torch.equal(got.losses["classification_loss"], w[0]) and torch.equal(got.grads[0], w[2])
    _assert_same_dets(got.detections(), w[4])
