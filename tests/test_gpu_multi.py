"""Multi-GPU parity (-m gpu, needs >= 2 devices): the image-sharded loss over NCCL equals the
single-process oracle on the full batch, and sharded post-processing equals the unsharded one."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sys.path.insert(0, ROOT)
    import synth_data as S
    from oracle import torch_oracle as O
    import pytorch_retinanet_b200 as P
    from pytorch_retinanet_b200.distributed import ShardedRetinaNetLosses, shard_range

    cfg = S.CONFIGS[1]
    n_total = world + 2            # uneven shards; with 8 ranks some hold a single image
    b = S.make_batch(cfg, 100, n_total, clustered=True)
    lo, hi = shard_range(n_total, rank, world)
    dev = torch.device("cuda", rank)
    anc = b["anchors"].to(dev)
    L = ShardedRetinaNetLosses(cfg.num_classes)
    x = b["cls_preds"][lo:hi].to(dev).requires_grad_(True)
    bb = b["bbox_preds"][lo:hi].to(dev).requires_grad_(True)
    tg = [{k: v.to(dev) for k, v in t.items()} for t in b["targets"][lo:hi]]
    out = L(tg, {"cls_preds": x, "bbox_preds": bb}, [anc] * (hi - lo))
    (out["classification_loss"] + out["regression_loss"]).backward()
    xf = b["cls_preds"].clone().requires_grad_(True)
    bf = b["bbox_preds"].clone().requires_grad_(True)
    full = O.batch_loss(b["targets"], xf, bf, [b["anchors"]] * n_total, cfg.num_classes)
    (full["classification_loss"] + full["regression_loss"]).backward()
    ok = True
    for k in full:
        ok &= abs(float(out[k]) - float(full[k])) <= 1e-5 * abs(float(full[k]))
    ok &= torch.allclose(x.grad.cpu(), xf.grad[lo:hi], rtol=2e-5, atol=1e-12)
    ok &= torch.allclose(bb.grad.cpu(), bf.grad[lo:hi], rtol=2e-5, atol=1e-9)
    ok &= int(L.last_stats[3]) == n_total
    # inference needs no communication: the shard's detections equal the same images of the full batch
    from types import SimpleNamespace
    stub = SimpleNamespace(score_thres=0.05, nms_thres=0.5, detections_per_img=100)
    dets = P.process_detections(stub, {"cls_preds": x.detach(), "bbox_preds": bb.detach()}, [anc] * (hi - lo),
                                b["im_szs"][lo:hi])
    want = O.postprocess(x.detach(), bb.detach(), [anc] * (hi - lo), b["im_szs"][lo:hi])
    for d, w in zip(dets, want):
        ok &= torch.equal(d["labels"], w["labels"]) and torch.equal(d["boxes"], w["boxes"])
    # the graph form of the sharded step: bit-identical to the sharded drop-in calls on every rank
    from pytorch_retinanet_b200.graphs import HotPathGraph
    res = HotPathGraph(cfg.num_classes, x.detach(), bb.detach(), anc, b["im_szs"][lo:hi], global_batch=n_total).step(tg)
    ok &= torch.equal(res.losses["classification_loss"], out["classification_loss"].detach())
    ok &= torch.equal(res.losses["regression_loss"], out["regression_loss"].detach())
    ok &= torch.equal(res.grads[0], x.grad) and torch.equal(res.grads[1], bb.grad)
    for d, w in zip(res.detections(), dets):
        ok &= torch.equal(d["labels"], w["labels"]) and torch.equal(d["boxes"], w["boxes"]) and torch.equal(d["scores"], w["scores"])
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_loss_nccl():
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29700 + os.getpid() % 1000, ret), nprocs=world, join=True)
    assert dict(ret) == {r: True for r in range(world)}
