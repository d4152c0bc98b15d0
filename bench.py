#!/usr/bin/env python
"""Benchmark of the dense per-anchor hot path (BASELINE.json metric: images/sec for
match + focal/smooth-L1 loss + decode + NMS).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the whole hot path over one batch of synthetic COCO-shaped inputs
(BASELINE.json configs[1]: 800x1333 -> A=201,600 anchors, 80 classes, batch 16 per GPU, <=100 GT/img):
  matcher + focal + smooth-L1 loss with gradients (rn_train_loss) and post-processing (sigmoid/threshold/decode/
  clip/NMS/top-100, rn_postprocess) of the same batch.  `value` runs it as HotPathGraph.step (one CUDA graph, double
  buffered); `api.*` holds the same step through the drop-in calls (RetinaNetLosses.forward + backward +
  process_detections) and the unpipelined graph; `e2e` is the same graph step fed from pinned HOST buffers.
N > 1: every rank owns its own 16 images (weak scaling; 8 GPUs = configs[2], batch 128); the ranks' 16-byte loss
vectors are summed inside the loss's final reduction kernel over peer-mapped memory (NVLink).
`extra.configs` adds the other BASELINE configs at their stated sizes (config 3: global batch 128 split over the
ranks; config 4: batch 256 inference with and without the top-k extension; config 5: batch 64, 500 GT/img).

Prints ONE JSON line (see README / DESIGN.md for the keys).  `--impl reference` times the UNMODIFIED reference's
own CPU implementation of the same step (baseline/_ref; the oracle port only if no reference tree is reachable) on
all host threads, every step a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import synth_data as S  # noqa: E402

METRIC = "images/sec (match+focal loss+decode+NMS)"
WORKLOAD_CFG = 2


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU every few ms (NVML) while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int, period_s: float = 0.004):
        self.index, self.period, self.rows, self.stop_flag, self.th = index, period_s, [], False, None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(i):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[i])
            except Exception:
                return i
        return i

    def _loop(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
                try:
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, rs))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.nvml is None:
            return
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()

    def stop(self):
        if self.nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.stop_flag = True
        self.th.join(timeout=1.0)
        n = self.nvml
        try:
            mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        except Exception:
            mx = None
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for r in self.rows:
            bits |= r[1]
        reasons = sorted(name for bit, name in self.REASONS.items() if bits & bit)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def workload_config(world):
    """The `config` object — identical in both arms (the driver compares them)."""
    cfg = S.CONFIGS[WORKLOAD_CFG]
    return {"workload": "BASELINE.json configs[1]: COCO-shaped 800x1333 (padded 800x1344), 80 classes, 9 anchors/loc "
                        "P3-P7 (A=201600), batch 16 per GPU, <=100 GT/img; step = loss fwd+bwd + post-process",
            "images_per_gpu": cfg.batch, "global_batch": cfg.batch * world, "anchors": S.num_anchors(cfg.padded_hw),
            "classes": cfg.num_classes, "parallelism": f"image-sharded dp{world}",
            "l2_policy": "inputs (1.09 GB/step/GPU) larger than L2 (126 MB); no explicit flush",
            "logits": "clustered N(-7,1.3^2) + N(5,1.5^2) on matched anchors (SURVEY.md 8d)"}


# ---- the reference's path, on whatever device the tensors live ------------------------------------------------------
def load_reference_or_port():
    """(kind, step function).  kind "reference" = the unmodified reference package driven through its own API
    (baseline/reference.py); "port" = the op-for-op oracle port, used only when no reference tree is reachable."""
    try:
        from baseline.reference import load_reference, reference_root, reference_step
        ref = load_reference()
        load_reference_or_port.where = os.path.relpath(reference_root(), ROOT) if reference_root().startswith(ROOT) else reference_root()

        def step(batch, n_images, backward=True, device=None):
            cfg = batch["config"]
            x, b, tg = batch["cls_preds"][:n_images], batch["bbox_preds"][:n_images], batch["targets"][:n_images]
            if device is not None:
                x, b = x.to(device), b.to(device)
                tg = [{k: v.to(device) for k, v in t.items()} for t in tg]
            return reference_step(ref, x, b, tg, cfg.padded_hw, batch["im_szs"][:n_images], cfg.num_classes, backward)
        return "reference", step
    except Exception as e:                                   # no reference tree on this machine
        sys.stderr.write(f"bench: reference tree unavailable ({e}); using the oracle port\n")
        from oracle import torch_oracle as O
        try:
            import torchvision
            nms_fn = torchvision.ops.nms
        except Exception:
            nms_fn = None

        def step(batch, n_images, backward=True, device=None):
            cfg = batch["config"]
            dev = torch.device("cpu") if device is None else device
            x = batch["cls_preds"][:n_images].to(dev).clone().requires_grad_(backward)
            b = batch["bbox_preds"][:n_images].to(dev).clone().requires_grad_(backward)
            tg = [{k: v.to(dev) for k, v in t.items()} for t in batch["targets"][:n_images]]
            anchors = [O.image_anchors(O.fpn_grid_sizes(*cfg.padded_hw), device=dev) for _ in range(n_images)]   # anchors.py:223-226
            out = O.batch_loss(tg, x, b, anchors, cfg.num_classes)
            if backward:
                (out["classification_loss"] + out["regression_loss"]).backward()
            dets = O.postprocess(x.detach(), b.detach(), anchors, batch["im_szs"][:n_images], nms_fn=nms_fn)
            return out, dets
        return "port", step


def run_reference(args, rank, world):
    """`--impl reference`: rank 0 alone times the reference's CPU path on all host threads, K steps after W warm-ups,
    each step a bounded sample of the workload's batch sized so that the whole run ends within a few minutes."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)              # every host thread the box has
    cfg = S.CONFIGS[WORKLOAD_CFG]
    kind, step = load_reference_or_port()
    batch = S.make_batch(cfg, 0, cfg.batch)
    step(batch, 1)                                          # untimed: first-touch / thread-pool start-up
    t0 = time.perf_counter()
    step(batch, 2)
    t_img = (time.perf_counter() - t0) / 2
    budget_s = 150.0
    sample = int(max(1, min(cfg.batch, budget_s / ((args.steps + args.warmup) * t_img))))
    for _ in range(args.warmup):
        step(batch, sample)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step(batch, sample)
        times.append(time.perf_counter() - t0)
    ms = 1000.0 * sum(times) / len(times)
    value = sample / (ms / 1000.0)
    cores = torch.get_num_threads()
    what = (f"the unmodified reference ({load_reference_or_port.where}: AnchorGenerator.forward, RetinaNetLosses.forward + "
            "backward, Retinanet.process_detections)" if kind == "reference" else "torch CPU eager port of the reference (oracle/torch_oracle.py)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world),
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": kind,
                         "sample": f"{sample} of the batch's {cfg.batch} images per step (reference cost is linear in images: "
                                   f"python loop per image, losses.py:126, models.py:181), {what}, torch CPU eager, "
                                   f"{cores} threads, os.cpu_count={os.cpu_count()}"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ---- rank placement (e2e ingest): cores and host memory next to the rank's GPU --------------------------------------
def _parse_cpulist(text):
    out = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


def place_rank(local_rank, local_world):
    """Pins this process to cores of the NUMA node its GPU hangs off (a disjoint slice per rank) and prefers that
    node for host allocations made from now on (the pinned staging buffers).  Best effort; returns what was done."""
    info = {"numa_node": None, "cpus": None, "mempolicy": None}
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        info["pci"] = bus
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read())
    except Exception as e:
        info["error"] = f"numa node of the GPU unknown: {e}"
        node = -1
    info["numa_node"] = node
    try:
        allowed = sorted(os.sched_getaffinity(0))
        cpus = allowed
        if node >= 0:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                near = [c for c in _parse_cpulist(f.read()) if c in set(allowed)]
            if near:
                cpus = near
        k = max(1, len(cpus) // max(1, local_world))
        mine = cpus[(local_rank * k) % len(cpus):][:k] or cpus
        os.sched_setaffinity(0, mine)
        info["cpus"] = f"{mine[0]}-{mine[-1]} ({len(mine)} of {len(allowed)} allowed)"
    except Exception as e:
        info["error"] = f"affinity: {e}"
    if node >= 0:
        try:                                                # set_mempolicy(MPOL_PREFERRED, {node})
            mask = ctypes.c_ulong(1 << node)
            rc = ctypes.CDLL(None, use_errno=True).syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))
            info["mempolicy"] = "preferred node %d" % node if rc == 0 else "set_mempolicy errno %d" % ctypes.get_errno()
        except Exception as e:
            info["mempolicy"] = f"unavailable: {e}"
    return info


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    import pytorch_retinanet_b200 as P
    from pytorch_retinanet_b200 import _native
    from pytorch_retinanet_b200.box_utils import PackedTargets
    from pytorch_retinanet_b200.detections import postprocess_batch
    from pytorch_retinanet_b200.distributed import ShardedRetinaNetLosses, close_exchanges, get_exchange
    from pytorch_retinanet_b200.graphs import HotPathGraph
    from pytorch_retinanet_b200.losses import fused_loss_forward
    from types import SimpleNamespace

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    placement = place_rank(local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    lib = _native.load()
    peak, peak_src = measured_peak_gbs()
    cfg = S.CONFIGS[WORKLOAD_CFG]
    n_img = cfg.batch
    first = rank * n_img                                   # per-image seeds are global image indices
    batch = S.make_batch(cfg, first, n_img, pin=True)
    A, C = batch["anchors"].shape[0], cfg.num_classes
    h_cls, h_box = batch["cls_preds"], batch["bbox_preds"]
    d_cls, d_box = h_cls.to(dev), h_box.to(dev)
    targets = [{k: v.to(dev) for k, v in t.items()} for t in batch["targets"]]
    gen = P.AnchorGenerator().to(dev)
    fmaps = [torch.empty((n_img, 1, h, w), device=dev) for h, w in S.grid_sizes(cfg.padded_hw)]
    images = SimpleNamespace(image_sizes=batch["im_szs"])
    exchange_mode = os.environ.get("RN_BENCH_EXCHANGE", "auto")       # auto = in-kernel peer exchange, NCCL if unavailable
    xch = get_exchange(None, exchange_mode) if world > 1 else None
    exchange_used = "none (1 GPU)" if world == 1 else ("peer memory, inside the loss's final reduction kernel" if xch is not None
                                                       else "NCCL all_reduce after the graph")
    losses = ShardedRetinaNetLosses(C, global_batch=n_img * world, exchange=xch if xch is not None else "nccl")
    if os.environ.get("RN_BENCH_NO_EXCHANGE"):              # DIAGNOSTIC ONLY (not a valid bench line): isolates the
        losses = P.RetinaNetLosses(C)                       # cost of the per-step exchange at N > 1
        exchange_used = "DISABLED (diagnostic run, losses are per-rank)"
    stub = SimpleNamespace(score_thres=0.05, nms_thres=0.5, detections_per_img=100)

    def step(cls, box):
        # Both halves of the path on the same batch, through the drop-in API.  The inference half goes first:
        # its one host sync (detection counts) then falls before the host-heavy enqueue of the training half
        # (packing, autograd), which overlaps with the GPU instead of preceding an idle wait.
        anchors = gen(images, fmaps)
        dets = P.process_detections(stub, {"cls_preds": cls, "bbox_preds": box}, anchors, batch["im_szs"])
        x, b = cls.detach().requires_grad_(True), box.detach().requires_grad_(True)
        out = losses(targets, {"cls_preds": x, "bbox_preds": b}, anchors)
        (out["classification_loss"] + out["regression_loss"]).backward()
        return out, dets, x.grad

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, streams=()):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        for s_ in streams:                                   # work the step put on its own streams is inside the timed region
            torch.cuda.current_stream(dev).wait_stream(s_)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    # ---- multi-GPU parity, where the driver can see it: the sharded step (graph + in-kernel exchange, and the drop-in
    # ShardedRetinaNetLosses) on 2 small images per rank against the CPU oracle on the whole 2*world-image batch ----
    parity = parity_check(P, HotPathGraph, ShardedRetinaNetLosses, xch, rank, world, dev, dist)
    if not parity["ok"]:
        if rank == 0:
            sys.stderr.write("bench: multi-GPU parity check FAILED: %s\n" % json.dumps(parity))
        raise SystemExit(3)

    # ---- device-resident throughput: the step as ONE CUDA graph (same C-ABI calls, two concurrent branches).
    # Two graphs on two input buffers alternate (double buffering): step i+1 is launched, then the losses, gradients
    # and detections of step i are read — every step's results are consumed, one step late, so the host work of a
    # step (GT packing, graph launch, count copy, slicing) overlaps with the GPU work of the other buffer. ----
    gsum_max = sum(int(t["boxes"].shape[0]) for t in targets)
    gb = (n_img * world) if (world > 1 and not os.environ.get("RN_BENCH_NO_EXCHANGE")) else None
    anc0 = gen(images, fmaps)[0]
    barrier()                                               # ranks leave data generation seconds apart
    gkw = dict(max_targets=max(4096, gsum_max), global_batch=gb, exchange=xch if xch is not None else "nccl")
    if os.environ.get("RN_BENCH_UNFUSED"):                  # DIAGNOSTIC: the round-1 form of the step (two concurrent branches)
        gkw["fused"] = False
    graph = HotPathGraph(C, d_cls, d_box, anc0, batch["im_szs"], **gkw)
    d_cls2, d_box2 = d_cls.clone(), d_box.clone()
    graph2 = HotPathGraph(C, d_cls2, d_box2, anc0, batch["im_szs"], **gkw)
    gpend = []
    use_pipe = not os.environ.get("RN_BENCH_NO_PIPELINE") and graph.fused
    if use_pipe:
        # the two buffers' graphs split into front (matcher, loss + filter) and tail (NMS) on two streams: the NMS of step
        # i runs under the front of step i+1 (graphs.HotPathPipeline)
        from pytorch_retinanet_b200.graphs import HotPathPipeline
        pkw = {k: v for k, v in gkw.items() if k != "fused"}
        if os.environ.get("RN_BENCH_EXCHANGE_ON_TAIL"):      # EXPERIMENTAL, see HotPathGraph(exchange_on_tail=...)
            pkw["exchange_on_tail"] = True
        pipe = HotPathPipeline(C, [(d_cls, d_box), (d_cls2, d_box2)], anc0, batch["im_szs"], **pkw)
        pipe_streams = pipe.streams
        nodes_per_step = pipe.kernel_nodes

        def step_graph_pipelined():
            gpend.append((pipe, pipe.step(targets)))        # pack GT + two graph launches (exchange inside the front)
            if len(gpend) > 1:
                r = gpend.pop(0)[1]
                return r.losses, r.detections(), r.grads    # detections() waits for that step's counts
    else:
        pipe_streams = ()
        nodes_per_step = graph.kernel_nodes

        def step_graph_pipelined():
            g = graph if (len(gpend) == 0 or gpend[-1][0] is graph2) else graph2
            gpend.append((g, g.step(targets)))              # pack GT + graph launch (exchange inside the graph)
            if len(gpend) > 1:
                r = gpend.pop(0)[1]
                return r.losses, r.detections(), r.grads    # detections() waits for that step's counts

    for _ in range(max(args.warmup, 4)):
        step_graph_pipelined()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.rn_launch_count()
    ms_step = timed(step_graph_pipelined, args.steps, pipe_streams)
    eager_launches = lib.rn_launch_count() - launches0      # launches our library issued directly (rn_pack_targets)
    gpu_launches = int(eager_launches + args.steps * nodes_per_step)
    while gpend:
        gpend.pop(0)[1].detections()
    if args.only_step:
        if rank == 0:
            sampler.stop()
            _emit({"metric": METRIC, "value": n_img * world / (ms_step * 1e-3), "unit": "images/s", "n_gpus": world,
                   "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "only_step": True,
                   "gpu_launches": gpu_launches})
        close_exchanges()
        return

    # ---- one graph, results read in the same step (one host sync per step) ----
    def step_graph():
        r = graph.step(targets)
        return r.losses, r.detections(), r.grads

    for _ in range(3):
        step_graph()
    ms_graph_sync = timed(step_graph, args.steps)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(20):
        graph.step(targets)
    host_us_graph = (time.perf_counter() - t0) / 20 * 1e6   # host time to enqueue one step (no wait)
    torch.cuda.synchronize(dev)

    # ---- the same step through the drop-in calls (reference signatures, autograd), synchronous ----
    for _ in range(args.warmup):
        step(d_cls, d_box)
    ms_dropin = timed(lambda: step(d_cls, d_box), args.steps)

    # ---- drop-in calls in graph mode (patch_retinanet(model, graph=True)): reference signatures + autograd, the kernels
    # replayed from CUDA graphs cached on the inputs' addresses ----
    losses_g = P.RetinaNetLosses(C, graph=True)
    stub_g = SimpleNamespace(score_thres=0.05, nms_thres=0.5, detections_per_img=100, rn_graph=True)
    ms_dropin_graph = None
    if world == 1:                                          # (single-GPU mode: the sharded loss keeps the eager path)
        def step_dropin_graph():
            # graph mode makes the host side of both calls cheap, so the training half goes FIRST (no host sync in it) and
            # the inference half — whose result read is the step's one host sync — is enqueued behind it while it runs
            anchors = gen(images, fmaps)
            x, b = d_cls.detach().requires_grad_(True), d_box.detach().requires_grad_(True)
            out = losses_g(targets, {"cls_preds": x, "bbox_preds": b}, anchors)
            # autograd.grad hands the gradients on as a model's backward would (to the head's CatBackward); .backward()
            # on these LEAF tensors would make AccumulateGrad clone the graph's static 1 GB buffer
            gx, gb_ = torch.autograd.grad(out["classification_loss"] + out["regression_loss"], (x, b))
            dets = P.process_detections(stub_g, {"cls_preds": d_cls, "bbox_preds": d_box}, anchors, batch["im_szs"])
            return out, dets, gx

        for _ in range(3):
            step_dropin_graph()
        ms_dropin_graph = timed(step_dropin_graph, args.steps)
        del losses_g, stub_g

    # ---- same work, software-pipelined: detections of step i are collected while step i+1 is enqueued ----
    pending = []

    def step_pipelined():
        anchors = gen(images, fmaps)
        x, b = d_cls.detach().requires_grad_(True), d_box.detach().requires_grad_(True)
        out = losses(targets, {"cls_preds": x, "bbox_preds": b}, anchors)
        (out["classification_loss"] + out["regression_loss"]).backward()
        pending.append(P.process_detections_async(stub, {"cls_preds": d_cls, "bbox_preds": d_box}, anchors, batch["im_szs"]))
        if len(pending) > 1:
            pending.pop(0).detections()

    for _ in range(3):
        step_pipelined()
    ms_pipe = timed(step_pipelined, args.steps)
    while pending:
        pending.pop(0).detections()
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end: pinned HOST inputs -> H2D every step -> graph step -> results copied back to the host.
    # The two graphs' input buffers alternate: the copy stream fills buffer (i+1)%2 while the graph of step i runs,
    # so the PCIe link never waits for the GPU and the GPU work hides behind the next step's ingest. ----
    graphs = (graph, graph2)
    copy_stream = torch.cuda.Stream(device=dev)
    host_targets = batch["targets"]                          # CPU tensors, as the reference's collate_fn hands them over
    M = 100
    h_out = [dict(boxes=torch.empty((n_img, M, 4), pin_memory=True), scores=torch.empty((n_img, M), pin_memory=True),
                  labels=torch.empty((n_img, M), dtype=torch.int64, pin_memory=True),
                  total=torch.empty((4,), pin_memory=True)) for _ in range(2)]
    free_ev = [None, None]                                    # buffer k's previous graph step has finished
    e2e_pending = []
    counter = [0]

    def e2e_step():
        k = counter[0] % 2
        counter[0] += 1
        g = graphs[k]
        main = torch.cuda.current_stream(dev)
        if free_ev[k] is not None:
            copy_stream.wait_event(free_ev[k])
        with torch.cuda.stream(copy_stream):
            g.cls_preds.copy_(h_cls, non_blocking=True)
            g.bbox_preds.copy_(h_box, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(copy_stream)
        main.wait_event(ready)
        r = g.step(host_targets)                            # GT: one pinned staging block, one async copy (no packing launch)
        ho = h_out[k]
        ho["total"].copy_(g.total, non_blocking=True)
        ho["boxes"].copy_(g.out_boxes, non_blocking=True)
        ho["scores"].copy_(g.out_scores, non_blocking=True)
        ho["labels"].copy_(g.out_labels, non_blocking=True)
        done = torch.cuda.Event()
        done.record(main)
        free_ev[k] = done
        e2e_pending.append((r, done))
        if len(e2e_pending) > 1:                             # consume step i-1 on the host while step i runs
            pr, pdone = e2e_pending.pop(0)
            counts = pr.result()[3]
            pdone.synchronize()
            return counts

    for _ in range(3):
        e2e_step()
    e2e_steps = max(3, min(args.steps, 10))
    ms_e2e = timed(e2e_step, e2e_steps)
    while e2e_pending:
        pr, pdone = e2e_pending.pop(0)
        pr.result()
        pdone.synchronize()
    gt_bytes = graph._off_bytes + gsum_max * 24
    h2d = h_cls.numel() * 4 + h_box.numel() * 4 + gt_bytes
    d2h = 16 + n_img * M * 28 + (n_img + 4) * 4

    # raw H2D rate of the same pinned buffers (all ranks at once): the PCIe / host-memory roof e2e is measured against
    def h2d_probe():
        d_cls2.copy_(h_cls, non_blocking=True)
        d_box2.copy_(h_box, non_blocking=True)

    for _ in range(2):
        h2d_probe()
    ms_probe = timed(h2d_probe, 5)
    probe_gbs = (h_cls.numel() + h_box.numel()) * 4 / (ms_probe * 1e-3) / 1e9
    del graph2, d_cls2, d_box2, graphs

    # ---- dominant kernel (fused loss fwd+grad) timed alone, live, for the roofline ----
    packed = PackedTargets([t["boxes"] for t in targets], [t["labels"] for t in targets], dev)
    anc = gen(images, fmaps)[0]

    def loss_only(want_grad):
        return fused_loss_forward(d_cls, d_box, anc, 0, packed, 0.25, 2.0, 0.1, 0.5, 0.4, float(n_img), want_grad)

    g_det = HotPathGraph(C, d_cls, d_box, anc, batch["im_szs"], train=False, detect=True)

    kern = {}
    for name, fn in (("loss_fwd_bwd", lambda: loss_only(True)), ("loss_fwd", lambda: loss_only(False)),
                     ("postprocess_graph", lambda: g_det.step().result()),
                     ("loss_kernel_alone", graph._enqueue_loss),     # loss_kernel<4,grad> + finalize on precomputed codes
                     ("loss_fwd_kernel_alone", lambda: graph._enqueue_loss(False)),
                     ("match_alone", graph._enqueue_match),
                     ("postprocess", lambda: postprocess_batch(d_cls, d_box, anc, 0, batch["im_szs"], 0.05, 0.5, 100))):
        for _ in range(3):
            fn()
        kern[name] = timed(fn, max(5, args.steps))

    # ---- row N1: the same work on the head's raw per-level conv outputs, and the cost of the
    # reference head's view/permute/contiguous/cat pass that it makes unnecessary (layers.py:189-195, 253-259)
    n1 = None
    if not args.no_levels:
        n1 = levels_leg(S, HotPathGraph, cfg, h_cls, h_box, dev, anc, packed, batch, targets, C, n_img, gsum_max, timed, args)

    gsum = sum(int(t["boxes"].shape[0]) for t in batch["targets"])
    bytes_fb = n_img * (2 * 4 * A * C + 2 * 16 * A) + 16 * A + 24 * gsum + 12 * n_img     # B_fb
    bytes_f = n_img * (4 * A * C + 16 * A) + 16 * A + 24 * gsum + 12 * n_img               # B_f
    bytes_p = n_img * (4 * A * C + 16 * A + 100 * 28) + 16 * A                             # B_p

    def bw(nbytes, ms, note=None):
        d = {"GBps": nbytes / (ms * 1e-3) / 1e9, "ms": ms, "frac": nbytes / (ms * 1e-3) / 1e9 / peak}
        if note:
            d["note"] = note
        return d

    # The dominant kernel of the timed step is the fused front of rn_train_detect (prep + matcher + loss_kernel<4,grad,FILTER>:
    # loss, gradients, score filter and final reduction in ONE pass over the logits): timed alone, live, by replaying the
    # front graph of the pipeline.  Its algorithmic bytes are B_fb — the filter rides on the same pass and adds none.
    ms_front = None
    if use_pipe:
        try:
            g_front = pipe.graphs[0].graph_front
            for _ in range(3):
                g_front.replay()
            ms_front = timed(g_front.replay, max(5, args.steps))
        except Exception as e:                               # never fatal: the entry falls back to the rn_train_loss call
            sys.stderr.write(f"bench: timing of the fused front failed ({e})\n")
            ms_front = None
    ms_dom = ms_front if ms_front else kern["loss_fwd_bwd"]
    ach_fb = bytes_fb / (ms_dom * 1e-3) / 1e9
    roofline = {"bound": "hbm",
                "kernel": ("rn_train_detect front = prep_kernel + match_kernel + loss_kernel<4,grad,FILTER> (training loss fwd+grad, the "
                           "post-processing's score filter and the final reduction in one pass over the logits; 82 % of the "
                           "step in the ncu launch list)" if ms_front else
                           "rn_train_loss = match_kernel + loss_kernel<4,grad> + finalize (training loss, fwd+grad in one pass over "
                           "the logits)"),
                "achieved": ach_fb, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": ach_fb / peak,
                "traffic": 2.094e9 + 3.3e6 + 0.2e6,
                "traffic_source": "NOT measured in this run: one `ncu --set full` capture of the unfused twin of the same launches "
                                  "(loss_kernel 1.059 GB read + 1.035 GB write, match_kernel 3.3 MB, finalize 0.2 MB), "
                                  "profiles/r02_ncu_loss_summary.txt; the fused kernel adds the candidate keys (~2 MB)",
                "bytes_per_launch": bytes_fb, "ms_per_launch": ms_dom,
                "others": {"train_loss_call": bw(bytes_fb, kern["loss_fwd_bwd"], "rn_train_loss = match_kernel + loss_kernel<4,grad> + "
                                                 "finalize (the unfused training half, what RetinaNetLosses.forward runs; round 1's "
                                                 "roofline entry)"),
                           "loss_fwd": bw(bytes_f, kern["loss_fwd"], "matcher + forward-only loss kernel + finalize (the reference's "
                                                                      "validation_step path)"),
                           "postprocess": bw(bytes_p, kern["postprocess"], "whole synchronous call: streaming score filter + lazy NMS "
                                                                           "+ the count copy/sync"),
                           "postprocess_graph": bw(bytes_p, kern["postprocess_graph"], "the same post-processing replayed from a CUDA "
                                                   "graph (HotPathGraph(train=False) / process_detections in graph mode): one launch + the "
                                                   "count copy/sync per call"),
                           "loss_fwd_kernel_alone": bw(bytes_f, kern["loss_fwd_kernel_alone"], "forward-only streaming kernel "
                                                                                                "(+ finalize) on precomputed codes"),
                           "loss_kernel_alone": bw(bytes_fb, kern["loss_kernel_alone"], "loss_kernel<4,grad> + finalize, codes "
                                                                                         "precomputed by rn_match"),
                           "match_alone": {"ms": kern["match_alone"], "iou_pairs_per_s": float(A) * gsum / (kern["match_alone"] * 1e-3),
                                           "note": "ALU-bound (SURVEY 8d): nominal anchor x GT pairs per second, not a bandwidth"},
                           "graph_step": bw(bytes_fb + bytes_p, ms_step, "B_fb + B_p (the algorithmic bytes of the two reference calls) "
                                                                         "over the whole timed step, target packing and result "
                                                                         "read-back included; the fused step streams the logits "
                                                                         "ONCE for both halves (actual traffic ~B_fb + candidates), "
                                                                         "so this fraction can exceed 1")}}

    # ---- the other BASELINE configs at their stated sizes ----
    extra = {"configs": other_configs(S, P, HotPathGraph, lib, dev, rank, world, timed, peak, xch, args)}

    # ---- reference eager code on CUDA tensors (the secondary comparator of SURVEY 8d), rank 0 ----
    cuda_eager = None
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        kind, ref_step = load_reference_or_port()
        sub = 4
        ref_step(batch, 1, True, dev)                        # warm-up (cuDNN-free, but allocator + kernels' first use)
        torch.cuda.synchronize(dev)
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            ref_step(batch, sub, True, dev)
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        cuda_eager = {"value": sub / best, "unit": "images/s", "ms_per_step": best * 1e3, "images": sub, "kind": kind,
                      "note": "the reference's own eager torch code (AnchorGenerator.forward, RetinaNetLosses.forward + backward, "
                              "process_detections) with CUDA tensors on this GPU, first %d images of the batch, best of 2 "
                              "(python loop per image and per class: cost is linear in images)" % sub}
        if world == 1:
            # ---- CPU baseline on the box's host cores (rank 0, N=1 only) ----
            torch.set_num_threads(os.cpu_count() or 1)
            ref_step(batch, 1)
            best = None
            for _ in range(2):
                t0 = time.perf_counter()
                ref_step(batch, n_img)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            cpu = {"value": n_img / best, "unit": "images/s", "cores": torch.get_num_threads(), "kind": kind,
                   "sample": f"the same {n_img}-image batch, best of 2 passes ({best:.2f} s per pass), "
                             + (f"the unmodified reference ({load_reference_or_port.where})" if kind == "reference" else "oracle port of the reference")
                             + f", torch CPU eager, os.cpu_count={os.cpu_count()}"}
    if world > 1:
        tp = torch.tensor([probe_gbs], device=dev)
        dist.all_reduce(tp)
        probe_total = float(tp.item())
    else:
        probe_total = probe_gbs
    if rank == 0:
        total = n_img * world
        e2e_gbs = h2d * world / (ms_e2e * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": total / (ms_step * 1e-3), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "clocks": clocks,
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e, "steps": e2e_steps,
                    "through": "HotPathGraph.step on two alternating input buffers: pinned host logits / boxes / GT -> H2D on a copy "
                               "stream -> graph -> losses + detection slabs copied back to pinned host memory, every step",
                    "ingest_GBps_all_gpus": e2e_gbs, "h2d_roof_GBps_all_gpus": probe_total,
                    "frac_of_h2d_roof": e2e_gbs / probe_total if probe_total else None,
                    "h2d_roof_note": "raw cudaMemcpyAsync rate of the same pinned buffers with all ranks copying at once, measured in "
                                     "this run: e2e is bound by host->device ingest (PCIe / host memory), not by a kernel",
                    "placement": placement},
            "api": {"value_through": "HotPathPipeline.step: rn_train_detect (matcher, then ONE pass over the logits for the loss with "
                                     "gradients, the score filter and the final reduction incl. the multi-GPU exchange, then the "
                                     "lazy NMS) captured as CUDA graphs on two alternating input buffers; every step's losses, "
                                     "gradients and detections are read one step late; anchors generated once (cached), outside "
                                     "the timed region" if use_pipe else "HotPathGraph.step on two alternating input buffers",
                    "pipeline": "front (matcher + loss/filter) and tail (NMS) of a step are two graphs on two streams: the NMS of "
                                "step i runs under the front of step i+1; results of step i are read after step i+1 is launched "
                                "(graph_sync = one graph, no pipelining)",
                    "graph_sync": {"value": total / (ms_graph_sync * 1e-3), "unit": "images/s", "ms_per_step": ms_graph_sync,
                                   "note": "one graph, results read in the same step (one host sync per step)"},
                    "graph_host_enqueue_us": host_us_graph,
                    "dropin_sync": {"value": total / (ms_dropin * 1e-3), "unit": "images/s", "ms_per_step": ms_dropin,
                                    "note": "RetinaNetLosses.forward + backward + process_detections (reference signatures, "
                                            "autograd), one host sync per step"},
                    "dropin_graph_sync": None if ms_dropin_graph is None else {
                        "value": total / (ms_dropin_graph * 1e-3), "unit": "images/s", "ms_per_step": ms_dropin_graph,
                        "note": "the same drop-in calls with graph=True (patch_retinanet(model, graph=True)): reference "
                                "signatures + autograd, kernels replayed from CUDA graphs cached on the inputs' addresses"},
                    "dropin_pipelined": {"value": total / (ms_pipe * 1e-3), "unit": "images/s", "ms_per_step": ms_pipe,
                                         "note": "drop-in calls with process_detections_async: results of step i read while "
                                                 "step i+1 is enqueued"}},
            "gpu_launches": gpu_launches,
            "gpu_launches_note": f"counted: {nodes_per_step} kernel nodes recorded in the captured graph(s) of a step x {args.steps} replays + "
                                 f"{eager_launches} launches issued directly by the library in the timed region (rn_launch_count)",
            "exchange": exchange_used, "parity_check": parity,
            "roofline": roofline, "cpu_baseline": cpu, "cuda_eager_baseline": cuda_eager, "n1_levels": n1, "extra": extra,
        }
        _emit(line)
    close_exchanges()


def parity_check(P, HotPathGraph, losses_cls, xch, rank, world, dev, dist):
    """Sharded loss of a 2*world-image batch (config-1 shape) through the graph step and the drop-in sharded loss, each
    rank holding 2 images, against the CPU oracle on the WHOLE batch (rank 0 computes it and broadcasts the numbers):
    global losses within 1e-5, this rank's gradient slice within 2e-5 (retinanet/losses.py:138-140)."""
    from oracle import torch_oracle as O
    cfg = S.CONFIGS[1]
    per = 2
    n_tot = per * world
    b = S.make_batch(cfg, 700, n_tot, clustered=True)
    lo = rank * per
    anc = b["anchors"].to(dev)
    x = b["cls_preds"][lo:lo + per].to(dev)
    bb = b["bbox_preds"][lo:lo + per].to(dev)
    tg = [{k: v.to(dev) for k, v in t.items()} for t in b["targets"][lo:lo + per]]
    if world > 1:
        dist.barrier()                                       # the first exchange follows: the ranks start it together
    g = HotPathGraph(cfg.num_classes, x, bb, anc, b["im_szs"][:per], global_batch=n_tot if world > 1 else None,
                     exchange=xch if xch is not None else "nccl")
    r = g.step(tg)
    got = torch.stack([r.losses["classification_loss"], r.losses["regression_loss"]]).cpu()
    L = losses_cls(cfg.num_classes, global_batch=n_tot, exchange=xch if xch is not None else "nccl") if world > 1 else P.RetinaNetLosses(cfg.num_classes)
    xg, bg = x.clone().requires_grad_(True), bb.clone().requires_grad_(True)
    out = L(tg, {"cls_preds": xg, "bbox_preds": bg}, [anc] * per)
    (out["classification_loss"] + out["regression_loss"]).backward()
    got2 = torch.stack([out["classification_loss"].detach(), out["regression_loss"].detach()]).cpu()
    same = bool(torch.equal(got, got2) and torch.equal(r.grads[0], xg.grad) and torch.equal(r.grads[1], bg.grad))
    # the oracle on the whole batch is cheap at this shape (49k anchors, 20 classes): every rank computes it
    xo = b["cls_preds"].clone().requires_grad_(True)
    bo = b["bbox_preds"].clone().requires_grad_(True)
    want = O.batch_loss(b["targets"], xo, bo, [b["anchors"]] * n_tot, cfg.num_classes)
    (want["classification_loss"] + want["regression_loss"]).backward()
    w = torch.stack([want["classification_loss"].detach(), want["regression_loss"].detach()])
    loss_err = float(((got - w).abs() / w.abs()).max())

    def close_err(a, ref, rtol, atol):       # max |a-ref| / (atol/rtol + |ref|): <= rtol  <=>  torch.allclose(a, ref, rtol, atol)
        return float(((a - ref).abs() / (atol / rtol + ref.abs())).max())

    grad_err = close_err(xg.grad.cpu(), xo.grad[lo:lo + per], 2e-5, 1e-12)
    # d(smooth-L1)/d(pred) = (pred - target) / beta / (F * N): the target's fp32 log differs by 1 ulp between the CPU
    # and the GPU libm, which the subtraction turns into an ABSOLUTE error of ~1e-8 on gradients of ~1e-4
    gbox_err = close_err(bg.grad.cpu(), bo.grad[lo:lo + per], 2e-5, 2e-8)
    ok = same and loss_err <= 1e-5 and grad_err <= 2e-5 and gbox_err <= 2e-5
    stats = torch.tensor([loss_err, grad_err, gbox_err, 0.0 if ok else 1.0], device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    if xch is not None and xch.error():
        stats[3] = 1.0
    s = stats.tolist()
    return {"ok": s[3] == 0.0, "loss_rel_err": s[0], "grad_logits_rel_err": s[1], "grad_bbox_rel_err": s[2],
            "graph_equals_dropin": same, "images": n_tot,
            "what": "global loss of 2 images/rank (config-1 shape) through HotPathGraph and ShardedRetinaNetLosses vs the CPU "
                    "oracle on the whole batch; max over ranks; losses: relative error <= 1e-5; gradients: |a-ref| / (atol/rtol + |ref|) "
                    "<= rtol = 2e-5 with atol 1e-12 (logits) / 2e-8 (boxes), i.e. torch.allclose"}


def levels_leg(S, HotPathGraph, cfg, h_cls, h_box, dev, anc, packed, batch, targets, C, n_img, gsum_max, timed, args):
    from pytorch_retinanet_b200.detections import postprocess_levels_async
    from pytorch_retinanet_b200.losses import fused_loss_forward_levels
    cls_lv = [t.to(dev) for t in S.nac_to_levels(h_cls, cfg.padded_hw)]
    box_lv = [t.to(dev) for t in S.nac_to_levels(h_box, cfg.padded_hw)]

    def relayout():
        outs = []
        for x in cls_lv:
            Nn, _, H, W = x.shape
            outs.append(x.view(Nn, -1, C, H, W).permute(0, 3, 4, 1, 2).contiguous().view(Nn, -1, C))
        return torch.cat(outs, dim=1)

    lv = {}
    for name, fn in (("loss_fwd_bwd", lambda: fused_loss_forward_levels(cls_lv, box_lv, anc, 0, packed, C, 0.25, 2.0, 0.1,
                                                                       0.5, 0.4, float(n_img), True)),
                     ("postprocess", lambda: postprocess_levels_async(cls_lv, box_lv, C, anc, 0, batch["im_szs"], 0.05, 0.5,
                                                                      100).result()),
                     ("reference_head_relayout_cls_fwd", relayout)):
        for _ in range(3):
            fn()
        lv[name] = timed(fn, max(5, args.steps))
    glv = HotPathGraph(C, cls_lv, box_lv, anc, batch["im_szs"], max_targets=max(4096, gsum_max))

    def step_graph_levels():
        r = glv.step(targets)
        return r.losses, r.detections(), r.grads

    for _ in range(3):
        step_graph_levels()
    lv["graph_step_sync"] = timed(step_graph_levels, max(5, args.steps))
    return {"ms": lv, "note": "loss / post-processing on raw [N, 9*C, H_l, W_l] conv outputs (no permute+cat); graph_step_sync = "
                              "HotPathGraph on the level lists, results read in the same step; "
                              "reference_head_relayout_cls_fwd = torch time of the re-layout pass this removes (forward only; "
                              "its backward costs the same again)"}


def other_configs(S, P, HotPathGraph, lib, dev, rank, world, timed, peak, xch, args):
    """BASELINE.json configs[2..4] at their stated sizes, on device-generated inputs (synth_data.make_batch_device).
    config 3: global batch 128 split over the ranks (strong scaling; the whole batch on one GPU at N = 1);
    config 4 (N = 1 only): batch 256 inference, top-k extension off and 1000/level;
    config 5 (N = 1 only): batch 64, 500 GT boxes per image: loss fwd+grad, matcher alone (IoU pairs/s), post-processing."""
    from pytorch_retinanet_b200.box_utils import PackedTargets, match_batch
    from pytorch_retinanet_b200.detections import postprocess_batch
    from pytorch_retinanet_b200.losses import fused_loss_forward
    out = {}
    reps = max(5, min(args.steps, 10))
    if args.no_extra:
        return None

    # ---- config 3 ----
    c3 = S.CONFIGS[3]
    per = c3.batch // world
    b = S.make_batch_device(c3, rank * per, per, dev)
    A, C = b["anchors"].shape[0], c3.num_classes
    gsum = sum(int(t["boxes"].shape[0]) for t in b["targets"])
    if world > 1:
        import torch.distributed as dist
        dist.barrier()                                       # the graph's warm-up exchanges: the ranks start together
    g = HotPathGraph(C, b["cls_preds"], b["bbox_preds"], b["anchors"], b["im_szs"], max_targets=max(4096, gsum),
                     global_batch=c3.batch if world > 1 else None, exchange=xch if xch is not None else "nccl")

    def step3():
        r = g.step(b["targets"])
        return r.losses, r.detections(), r.grads

    for _ in range(3):
        step3()
    ms = timed(step3, reps)
    nbytes = per * (2 * 4 * A * C + 2 * 16 * A) + per * (4 * A * C + 16 * A + 100 * 28) + 32 * A + 24 * gsum
    out["config3"] = {"what": "BASELINE.json configs[2]: batch 128 image-sharded, step = loss fwd+grad + post-process "
                              "(HotPathGraph.step, results read in the same step)", "global_batch": c3.batch,
                      "images_per_gpu": per, "n_gpus": world, "ms_per_step": ms, "images_per_s": c3.batch / (ms * 1e-3),
                      "hbm_frac_per_gpu": nbytes / (ms * 1e-3) / 1e9 / peak, "scaling": "strong", "steps": reps,
                      "hbm_frac_note": "B_fb + B_p of the two reference calls over the step time; the fused step reads the logits once, "
                                       "so the fraction can exceed 1"}
    del g, b
    torch.cuda.empty_cache()
    if world > 1:
        return out

    # ---- config 4: inference, batch 256 ----
    c4 = S.CONFIGS[4]
    b = S.make_batch_device(c4, 0, c4.batch, dev)
    A, C, N = b["anchors"].shape[0], c4.num_classes, c4.batch
    offs = [0]
    for h, w in S.grid_sizes(c4.padded_hw):
        offs.append(offs[-1] + 9 * h * w)
    bytes_p = N * (4 * A * C + 16 * A + 100 * 28) + 16 * A
    res4 = {"what": "BASELINE.json configs[3]: inference post-processing, batch 256, score 0.05, NMS 0.5, 100 dets/img "
                    "(synchronous postprocess_batch call incl. the count copy)", "batch": N, "steps": reps}
    for name, topk in (("topk_none", None), ("topk_1000_per_level", 1000)):
        def pp():
            return postprocess_batch(b["cls_preds"], b["bbox_preds"], b["anchors"], 0, b["im_szs"], 0.05, 0.5, 100,
                                     pre_nms_topk=topk, level_offsets=offs if topk else None)
        for _ in range(2):
            r = pp()
        ms = timed(pp, reps)
        res4[name] = {"ms": ms, "images_per_s": N / (ms * 1e-3), "hbm_frac": bytes_p / (ms * 1e-3) / 1e9 / peak,
                      "detections": int(sum(r[3]))}
    out["config4"] = res4
    del b
    torch.cuda.empty_cache()

    # ---- config 5: dense crowd, batch 64, 500 GT/img ----
    c5 = S.CONFIGS[5]
    b = S.make_batch_device(c5, 0, c5.batch, dev)
    A, C, N = b["anchors"].shape[0], c5.num_classes, c5.batch
    packed = PackedTargets([t["boxes"] for t in b["targets"]], [t["labels"] for t in b["targets"]], dev)
    gsum = packed.total

    def loss5(want):
        return fused_loss_forward(b["cls_preds"], b["bbox_preds"], b["anchors"], 0, packed, 0.25, 2.0, 0.1, 0.5, 0.4, float(N), want)

    def match5():
        return match_batch(b["anchors"], 0, packed, A, 0.5, 0.4, False, True)

    def pp5():
        return postprocess_batch(b["cls_preds"], b["bbox_preds"], b["anchors"], 0, b["im_szs"], 0.05, 0.5, 100)

    t5 = {}
    for name, fn in (("loss_fwd_bwd", lambda: loss5(True)), ("loss_fwd", lambda: loss5(False)), ("match", match5), ("postprocess", pp5)):
        for _ in range(2):
            fn()
        t5[name] = timed(fn, reps)
    bytes_fb = N * (2 * 4 * A * C + 2 * 16 * A) + 16 * A + 24 * gsum + 12 * N
    bytes_f = N * (4 * A * C + 16 * A) + 16 * A + 24 * gsum + 12 * N
    bytes_p = N * (4 * A * C + 16 * A + 100 * 28) + 16 * A
    pairs = float(A) * gsum
    out["config5"] = {"what": "BASELINE.json configs[4]: dense crowd 1024x1024, 500 GT boxes/img, 80 classes, batch 64",
                      "batch": N, "gt_boxes": gsum, "steps": reps,
                      "loss_fwd_bwd": {"ms": t5["loss_fwd_bwd"], "images_per_s": N / (t5["loss_fwd_bwd"] * 1e-3),
                                       "hbm_frac": bytes_fb / (t5["loss_fwd_bwd"] * 1e-3) / 1e9 / peak},
                      "loss_fwd": {"ms": t5["loss_fwd"], "hbm_frac": bytes_f / (t5["loss_fwd"] * 1e-3) / 1e9 / peak},
                      "match": {"ms": t5["match"], "iou_pairs": pairs, "iou_pairs_per_s": pairs / (t5["match"] * 1e-3),
                                "note": "ALU-bound: nominal anchor x GT pairs (A x sum G) per second, not a bandwidth"},
                      "postprocess": {"ms": t5["postprocess"], "images_per_s": N / (t5["postprocess"] * 1e-3),
                                      "hbm_frac": bytes_p / (t5["postprocess"] * 1e-3) / 1e9 / peak}}
    del b
    torch.cuda.empty_cache()
    return out


def _emit(line: dict) -> None:
    """The ONE JSON line, written to the process's original stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


# Libraries (NCCL's version banner, torchrun) print to fd 1: keep the real stdout for the JSON line only and send
# everything else to stderr.
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU and CUDA-eager reference legs")
    ap.add_argument("--no-levels", action="store_true", help="skip the row-N1 (per-level NCHW) timing leg")
    ap.add_argument("--no-extra", action="store_true", help="skip extra.configs (BASELINE configs 3-5 at their stated sizes)")
    ap.add_argument("--only-step", action="store_true", help="profiling aid: run only the warm-up and the timed graph steps "
                                                                "(what `value` measures) and print a reduced line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    args.steps = max(args.steps, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
