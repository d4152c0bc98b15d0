"""CPU tests (-m "not gpu") of the host-side logic: C-ABI symbol coverage, argument checking that
needs no GPU, packing of ragged targets, the integration swap, and the 2-rank (gloo) logic of the
image-sharded loss with the CUDA kernel replaced by the oracle."""
import ctypes
import os
import re
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    from pytorch_retinanet_b200 import _native
    hdr = open(os.path.join(ROOT, "include", "retinanet_b200.h")).read()
    declared = set(re.findall(r"\b(rn_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    lib = _native.load()                                   # loads without a GPU
    raw = ctypes.CDLL(_native.lib_path())
    for name in declared:
        assert hasattr(raw, name), name
    assert lib.rn_abi_version() == _native.header_abi_version() == 2
    # argument errors are reported through return codes + rn_last_error (no GPU touched)
    assert lib.rn_match(None, 10, 0, None, None, None, 1, 0, 0.5, 0.4, None, None, None, None) == -1
    assert b"null" in lib.rn_last_error()
    # the packed match codes keep the GT index in 20 bits: too many boxes is an error, not silent corruption
    one = ctypes.c_void_p(16)                               # never dereferenced: the size check comes first
    assert lib.rn_match(one, 10, 0, one, one, one, 1, 1 << 20, 0.5, 0.4, None, one, one, None) == -2
    assert b"2^20" in lib.rn_last_error()
    assert lib.rn_match(one, 10, 0, one, None, one, 1, 1 << 20, 0.3, 0.4, one, None, None, None) == -1   # box_utils.py:66
    bad = _native.RnExchange()
    bad.rank, bad.world = 3, 2
    assert lib.rn_exchange_total(one, ctypes.byref(bad), None) == -1
    assert lib.rn_comm_bytes() == 2048
    # rn_train_detect: argument errors before any launch (phases mask, null pointers, workspace size)
    args = [one] * 3 + [0] + [one] * 3 + [2, 10, 100, 4, 0.5, 0.4, 0.25, 2.0, 0.1, one, 2.0] + [one] * 7 + \
           [0.05, 0.5, 100, 0, None, 0, 1 << 16] + [one] * 6 + [0, one, 1 << 30, None, None, 9]
    assert lib.rn_train_detect(*args) == -1 and b"phases" in lib.rn_last_error()
    args[-1] = 7
    args[-4] = 16                                           # workspace far too small
    assert lib.rn_train_detect(*args) == -3
    assert lib.rn_train_detect_workspace_bytes(16, 201600, 80, 1 << 20, 100) > \
        lib.rn_loss_workspace_bytes(16, 201600, 80) + lib.rn_postprocess_workspace_bytes(16, 201600, 80, 1 << 20, 100) - 512
    assert lib.rn_postprocess_workspace_bytes(16, 201600, 80, 1 << 20, 100) > (1 << 20) * 8
    assert lib.rn_loss_workspace_bytes(16, 201600, 80) == (16 * 788 * 2 + 16 * 2 + 2 + 8) * 8


def test_no_cpu_fallback_and_loud_failure():
    import pytorch_retinanet_b200 as P
    from pytorch_retinanet_b200 import _native
    a = torch.rand(8, 4)
    with pytest.raises(_native.NativeError):
        P.matcher(a, a[:2])
    with pytest.raises(_native.NativeError):
        P.AnchorGenerator().grid_anchors([(4, 4)] * 5, torch.device("cpu"))
    with pytest.raises(AssertionError):
        P.matcher(a, a[:2], match_thr=0.3, back_thr=0.4)   # box_utils.py:66
    src = "".join(open(os.path.join(ROOT, "pytorch_retinanet_b200", f)).read()
                  for f in os.listdir(os.path.join(ROOT, "pytorch_retinanet_b200")) if f.endswith(".py"))
    assert "oracle" not in src                              # the product never imports the oracle


def test_packed_targets_and_anchor_generator_host_side():
    from pytorch_retinanet_b200.box_utils import PackedTargets
    from pytorch_retinanet_b200 import AnchorGenerator
    from oracle import torch_oracle as O
    dev = torch.device("cpu")
    boxes = [torch.rand(3, 4), torch.zeros(0, 4), torch.rand(5, 4).double()]
    labels = [torch.tensor([1, 2, 3]), torch.zeros(0, dtype=torch.int64), torch.tensor([4, 5, 6, 7, 8], dtype=torch.int32)]
    p = PackedTargets(boxes, labels, dev)
    assert p.offsets.tolist() == [0, 3, 3, 8] and p.boxes.shape == (8, 4) and p.boxes.dtype == torch.float32
    assert p.labels.dtype == torch.int64 and p.labels.tolist() == [1, 2, 3, 4, 5, 6, 7, 8]
    g = AnchorGenerator()
    assert g.num_anchors == [9] * 5 and list(g.state_dict().keys()) == [f"cell_anchors.{i}" for i in range(5)]
    for i, c in enumerate(g.cell_anchors):
        assert torch.equal(c, O.cell_anchor_table(O.SIZES[i], O.RATIOS))
    with pytest.raises(AssertionError):
        AnchorGenerator(sizes=[[32.0], [64.0]], strides=[8, 16, 32])


def test_patch_retinanet_on_the_real_reference():
    from oracle.ref_shim import load_reference, reference_available
    if not reference_available():
        pytest.skip("/root/reference not present")
    ref = load_reference()
    import pytorch_retinanet_b200 as P
    model = ref.Retinanet(num_classes=7, backbone_kind="resnet18", pretrained=False)
    keys_before = list(model.state_dict().keys())
    cells_before = [c.clone() for c in model.anchor_generator.cell_anchors]
    P.patch_retinanet(model)
    assert isinstance(model.anchor_generator, P.AnchorGenerator)
    assert isinstance(model.retinanet_head.losses, P.RetinaNetLosses) and model.retinanet_head.losses.n_c == 7
    assert model.process_detections.__func__ is P.process_detections
    assert list(model.state_dict().keys()) == keys_before
    assert all(torch.equal(a, b) for a, b in zip(model.anchor_generator.cell_anchors, cells_before))


# ---- 2-rank gloo test of the sharded loss: the CUDA call is replaced by the oracle -------------------
def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth_data as S
    from oracle import torch_oracle as O
    import pytorch_retinanet_b200.distributed as D

    class FakeFused:                                         # same contract as _FusedRetinaNetLoss.apply
        @staticmethod
        def apply(cls, box, anchors, stride, packed, hp):
            rows = []
            off = packed.offsets.tolist()
            for i in range(cls.shape[0]):
                gt, lab = packed.boxes[off[i]:off[i + 1]], packed.labels[off[i]:off[i + 1]]
                r, c, _, f = O.image_loss(anchors, cls[i], box[i], lab, gt if off[i + 1] > off[i] else gt[:0], cls.shape[-1])
                rows.append(torch.stack([c, r, f.float()]))
            image = torch.stack(rows) if rows else torch.zeros((0, 3))
            c, r = image[:, 0].sum() / hp["batch_div"], image[:, 1].sum() / hp["batch_div"]
            total = torch.stack([c.detach(), r.detach(), image[:, 2].sum().detach(), torch.tensor(float(cls.shape[0]))])
            if hp.get("all_reduce_group", False) is not False:
                dist.all_reduce(total, group=hp["all_reduce_group"])
            total[:2] *= hp.get("total_scale", 1.0)           # grad_reduction="mean": local values carry a factor world
            # value of the global batch, gradient of the local shard (what the CUDA autograd function does)
            return c + (total[0] - c.detach()), r + (total[1] - r.detach()), image.detach(), total

    D._FusedRetinaNetLoss = FakeFused
    cfg = S.CONFIGS[1]
    n_total = 5                                              # uneven shards: 2 + 3
    b = S.make_batch(cfg, 0, n_total, clustered=True)
    b["targets"][1] = {"boxes": torch.zeros((0, 4)), "labels": torch.zeros((0,), dtype=torch.int64)}
    lo, hi = D.shard_range(n_total, rank, world)
    L = D.ShardedRetinaNetLosses(cfg.num_classes)           # global batch discovered by all-reduce
    x = b["cls_preds"][lo:hi].clone().requires_grad_(True)
    out = L(b["targets"][lo:hi], {"cls_preds": x, "bbox_preds": b["bbox_preds"][lo:hi]}, [b["anchors"]] * (hi - lo))
    (out["classification_loss"] + out["regression_loss"]).backward()
    full = O.batch_loss(b["targets"], b["cls_preds"], b["bbox_preds"], [b["anchors"]] * n_total, cfg.num_classes)
    xf = b["cls_preds"].clone().requires_grad_(True)
    ff = O.batch_loss(b["targets"], xf, b["bbox_preds"], [b["anchors"]] * n_total, cfg.num_classes)
    (ff["classification_loss"] + ff["regression_loss"]).backward()
    ok = (abs(float(out["classification_loss"]) - float(full["classification_loss"])) <= 1e-5 * float(full["classification_loss"])
          and abs(float(out["regression_loss"]) - float(full["regression_loss"])) <= 1e-5 * float(full["regression_loss"])
          and torch.allclose(x.grad, xf.grad[lo:hi], rtol=1e-5, atol=1e-12)
          and int(L.last_stats[3]) == n_total)
    # fewer images than ranks: rank 0's shard is EMPTY and must still take part in the exchange (no anchors / targets to
    # look at on that rank), in both gradient conventions
    b1 = S.make_batch(cfg, 9, 1, clustered=True)
    lo, hi = D.shard_range(1, rank, world)
    full1 = O.batch_loss(b1["targets"], b1["cls_preds"], b1["bbox_preds"], [b1["anchors"]], cfg.num_classes)
    for red in ("sum", "mean"):
        L1 = D.ShardedRetinaNetLosses(cfg.num_classes, grad_reduction=red)
        x1 = b1["cls_preds"][lo:hi].clone().requires_grad_(True)
        out1 = L1(b1["targets"][lo:hi], {"cls_preds": x1, "bbox_preds": b1["bbox_preds"][lo:hi]}, [b1["anchors"]] * (hi - lo))
        ok = ok and (hi - lo) == (0 if rank == 0 else 1)
        ok = ok and abs(float(out1["classification_loss"]) - float(full1["classification_loss"])) <= 1e-5 * float(full1["classification_loss"])
        ok = ok and int(L1.last_stats[3]) == 1
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_loss_two_ranks_gloo():
    from pytorch_retinanet_b200.distributed import shard_range
    assert [shard_range(5, r, 2) for r in range(2)] == [(0, 2), (2, 5)]
    assert [shard_range(128, r, 8) for r in range(8)] == [(16 * r, 16 * r + 16) for r in range(8)]
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_graph_api_rejects_cpu_tensors_and_bad_arguments():
    """HotPathGraph validates its arguments before touching the GPU: CPU tensors are refused (no CPU path)."""
    import pytest
    import torch
    from pytorch_retinanet_b200 import _native
    from pytorch_retinanet_b200.graphs import HotPathGraph
    x, b, a = torch.zeros((2, 9, 4)), torch.zeros((2, 9, 4)), torch.zeros((9, 4))
    with pytest.raises(ValueError):
        HotPathGraph(4, x, b, a, train=False, detect=False)
    try:
        _native.load()
    except _native.NativeError:
        pytest.skip("CUDA library not built on this machine")
    with pytest.raises(_native.NativeError):
        HotPathGraph(4, x, b, a, [(8, 8), (8, 8)])
    with pytest.raises(_native.NativeError):
        HotPathGraph(4, [x.view(2, 36, 1, 1)], [b.view(2, 36, 1, 1)], a, [(8, 8), (8, 8)])


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the reference's CPU path on the host cores) must print exactly ONE JSON line on
    stdout with the contract's keys — it runs without a GPU."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("images/sec") and d["steps"] == 1 and d["n_gpus"] == 1
    from baseline.reference import reference_available
    assert d["cpu_baseline"]["kind"] == ("reference" if reference_available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"] and d["warmup"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
