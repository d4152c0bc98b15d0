"""Multi-GPU parity (-m gpu, needs >= 2 devices): the image-sharded loss equals the single-process oracle on the full
batch — with the in-kernel peer-memory exchange (default) and with the NCCL all-reduce — including an EMPTY shard,
the per-level entry point, the DDP-style gradient scaling and the CUDA-graph form; sharded post-processing equals the
unsharded one; the exchange itself is stress-tested over a few hundred back-to-back steps."""
import os
import sys
import time

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret, mode):
    try:
        _worker_body(rank, world, port, ret, mode)
    except BaseException:
        import traceback
        ret[rank] = ["exception: " + traceback.format_exc()[-1500:]]
        raise


def _worker_body(rank, world, port, ret, mode):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))      # the CPU oracle runs in every rank: do not oversubscribe
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sys.path.insert(0, ROOT)
    import synth_data as S
    from oracle import torch_oracle as O
    import pytorch_retinanet_b200 as P
    from pytorch_retinanet_b200.distributed import (PeerExchange, ShardedRetinaNetLosses, close_exchanges, get_exchange,
                                                    shard_range)

    fails = []

    def check(cond, what):
        if not cond:
            fails.append(what)

    dev = torch.device("cuda", rank)
    xch = get_exchange(None, mode)
    check((xch is not None) == (mode == "peer"), "exchange kind")

    # ---- the exchange by itself: 300 back-to-back steps, values that identify (rank, step); same bits on every rank ----
    if xch is not None:
        tot = torch.empty((300, 4), device=dev)
        for s in range(300):
            tot[s] = torch.tensor([rank + 1.0, s * 0.5, (rank + 1.0) * s, 1.0], device=dev)
        for s in range(300):
            xch.exchange(tot[s])
        rs = torch.arange(1, world + 1, dtype=torch.float64)
        want = torch.stack([torch.stack([rs.sum(), torch.tensor(s * 0.5 * world, dtype=torch.float64), rs.sum() * s,
                                         torch.tensor(float(world), dtype=torch.float64)]) for s in range(300)]).float()
        check(torch.equal(tot.cpu(), want), "exchange stress values")
        check(not xch.error(), "exchange time-out flag")

    cfg = S.CONFIGS[1]
    for n_total in (world + 2, world - 1):      # uneven shards; then FEWER images than ranks: one shard is empty
        b = S.make_batch(cfg, 100, n_total, clustered=True)
        lo, hi = shard_range(n_total, rank, world)
        anc = b["anchors"].to(dev)
        xf = b["cls_preds"].clone().requires_grad_(True)
        bf = b["bbox_preds"].clone().requires_grad_(True)
        full = O.batch_loss(b["targets"], xf, bf, [b["anchors"]] * n_total, cfg.num_classes)
        (full["classification_loss"] + full["regression_loss"]).backward()
        tg = [{k: v.to(dev) for k, v in t.items()} for t in b["targets"][lo:hi]]
        for red in ("sum", "mean"):
            L = ShardedRetinaNetLosses(cfg.num_classes, exchange=mode, grad_reduction=red)
            x = b["cls_preds"][lo:hi].to(dev).requires_grad_(True)
            bb = b["bbox_preds"][lo:hi].to(dev).requires_grad_(True)
            out = L(tg, {"cls_preds": x, "bbox_preds": bb}, [anc] * (hi - lo))
            (out["classification_loss"] + out["regression_loss"]).backward()
            gs = float(world) if red == "mean" else 1.0          # DDP averages: local gradients are world x larger
            for k in full:
                check(abs(float(out[k]) - float(full[k])) <= 1e-5 * abs(float(full[k])), f"{k} n={n_total} {red}")
            check(torch.allclose(x.grad.cpu(), gs * xf.grad[lo:hi], rtol=2e-5, atol=1e-12), f"grad logits n={n_total} {red}")
            check(torch.allclose(bb.grad.cpu(), gs * bf.grad[lo:hi], rtol=2e-5, atol=1e-9), f"grad bbox n={n_total} {red}")
            check(int(L.last_stats[3]) == n_total, f"N in stats n={n_total} {red}")
            # every rank holds the same bits
            mine = torch.stack([out["classification_loss"].detach(), out["regression_loss"].detach()])
            allv = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allv, mine)
            check(all(torch.equal(v, allv[0]) for v in allv), f"identical on all ranks n={n_total} {red}")
        if n_total < world:          # some rank holds an empty shard: the remaining (collective) checks need every rank
            continue
        # the per-level entry point through the sharded loss (fuse_head_layout=True heads)
        cls_lv = [t.to(dev).requires_grad_(True) for t in S.nac_to_levels(b["cls_preds"][lo:hi], cfg.padded_hw)]
        box_lv = [t.to(dev).requires_grad_(True) for t in S.nac_to_levels(b["bbox_preds"][lo:hi], cfg.padded_hw)]
        L = ShardedRetinaNetLosses(cfg.num_classes, exchange=mode)
        ol = L(tg, {"cls_levels": cls_lv, "bbox_levels": box_lv}, [anc] * (hi - lo))
        for k in full:
            check(abs(float(ol[k]) - float(full[k])) <= 1e-5 * abs(float(full[k])), f"levels {k} n={n_total}")
        # inference needs no communication: the shard's detections equal the same images of the full batch
        from types import SimpleNamespace
        stub = SimpleNamespace(score_thres=0.05, nms_thres=0.5, detections_per_img=100)
        dets = P.process_detections(stub, {"cls_preds": x.detach(), "bbox_preds": bb.detach()}, [anc] * (hi - lo),
                                    b["im_szs"][lo:hi])
        want = O.postprocess(x.detach(), bb.detach(), [anc] * (hi - lo), b["im_szs"][lo:hi])
        for d, w in zip(dets, want):
            check(torch.equal(d["labels"], w["labels"]) and torch.equal(d["boxes"], w["boxes"]), f"detections n={n_total}")

    # ---- the graph form of the sharded step (every rank needs >= 1 image: static shapes): bit-identical to the
    # sharded drop-in calls on every rank, step after step (the exchange is captured inside the graph) ----
    n_total = 2 * world
    b = S.make_batch(cfg, 200, n_total, clustered=True)
    lo, hi = 2 * rank, 2 * rank + 2
    anc = b["anchors"].to(dev)
    tg = [{k: v.to(dev) for k, v in t.items()} for t in b["targets"][lo:hi]]
    x, bb = b["cls_preds"][lo:hi].to(dev), b["bbox_preds"][lo:hi].to(dev)
    L = ShardedRetinaNetLosses(cfg.num_classes, global_batch=n_total, exchange=mode)
    xg, bg = x.clone().requires_grad_(True), bb.clone().requires_grad_(True)
    out = L(tg, {"cls_preds": xg, "bbox_preds": bg}, [anc] * 2)
    (out["classification_loss"] + out["regression_loss"]).backward()
    from pytorch_retinanet_b200.graphs import HotPathGraph
    g = HotPathGraph(cfg.num_classes, x, bb, anc, b["im_szs"][lo:hi], global_batch=n_total, exchange=mode)
    check((g._xch is not None) == (mode == "peer"), "graph exchange kind")
    for it in range(5):
        res = g.step(tg)
        check(torch.equal(res.losses["classification_loss"], out["classification_loss"].detach()), f"graph cls step {it}")
        check(torch.equal(res.losses["regression_loss"], out["regression_loss"].detach()), f"graph reg step {it}")
        check(torch.equal(res.grads[0], xg.grad) and torch.equal(res.grads[1], bg.grad), f"graph grads step {it}")
        check(len(res.detections()) == 2, "graph detections")
    full = O.batch_loss(b["targets"], b["cls_preds"], b["bbox_preds"], [b["anchors"]] * n_total, cfg.num_classes)
    check(abs(float(out["classification_loss"]) - float(full["classification_loss"])) <= 1e-5 * float(full["classification_loss"]),
          "graph-shaped batch vs oracle")
    if xch is not None:
        check(not xch.error(), "exchange time-out flag (end)")
        check(isinstance(xch, PeerExchange), "type")
    torch.cuda.synchronize()
    close_exchanges()
    ret[rank] = fails
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["peer", "nccl"])
@pytest.mark.parametrize("ranks", [2, 8])
def test_sharded_loss_multi_gpu(ranks, mode):
    have = torch.cuda.device_count()
    if have < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(have, ranks)
    if ranks == 8 and have <= 2:
        pytest.skip("covered by the 2-rank case")
    mgr = mp.Manager()
    ret = mgr.dict()
    ctx = mp.spawn(_worker, args=(world, 29700 + (os.getpid() + 7 * ranks + (3 if mode == "peer" else 0)) % 1000, ret, mode),
                   nprocs=world, join=False)
    deadline = time.time() + 900                    # a rank that stops taking part must not hang the suite (8 ranks: ~3 min)
    try:
        while not ctx.join(timeout=5):
            if time.time() > deadline:
                raise TimeoutError("multi-GPU worker processes did not finish within 900 s")
    finally:
        for p in ctx.processes:
            if p.is_alive():
                p.kill()
    assert dict(ret) == {r: [] for r in range(world)}
