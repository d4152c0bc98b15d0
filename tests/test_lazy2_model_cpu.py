"""CPU model of ``lazy2_nms_kernel``'s ALGORITHM (pytorch_retinanet_b200/csrc/postprocess.cu), checked against the oracle.

The kernel replaces "sort everything, NMS class by class, merge, take the top max_det" (retinanet/models.py:193-240) by:
score-bin counting sort of a PREFIX of the global order -> stable regrouping by class -> per-class greedy NMS that
advances in waves of ranks and stops once max_det candidates are kept -> compaction by rank; plus a rule for when the
prefix was not enough (second attempt with the long prefix, then hand-over to the older kernel).  This file restates
exactly that control flow in numpy — same constants, same cut / wave / cap / "more" rules — and compares it with the
reference semantics (oracle.torch_oracle.nms_keep per class + stable global order) on random and adversarial inputs:
heavy ties, one dominant class, everything suppressed, fewer candidates than max_det, prefixes that are too short.
It is a specification test of the algorithm, not of the CUDA code (the -m gpu tests do that bit for bit)."""
import numpy as np
import pytest
import torch

from oracle import torch_oracle as O

CAP, FIRST, WAVE, BINS, BIN_MAX = 4096, 1024, 512, 2048, 256       # LZ2_* in postprocess.cu


def _bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.int64)


def lazy2_model(scores, classes, anchors, boxes, thr, nms_thr, max_det, A):
    """Returns ("ok", ranks of the detections in output order) or ("v1", None) when the kernel would hand the image over."""
    K = len(scores)
    if K == 0:
        return "ok", np.zeros((0,), dtype=np.int64)
    hi = (~_bits(scores)) & 0xFFFFFFFF
    key = (hi << 32) | (classes.astype(np.int64) * A + anchors.astype(np.int64))
    hi0 = (~_bits(1.0)) & 0xFFFFFFFF
    rng = int(_bits(1.0) - _bits(max(thr, 0.0)))
    rng = max(rng, 1)
    shift = 0
    while (rng >> shift) >= BINS:
        shift += 1
    bins = np.where(hi <= hi0, 0, np.minimum((hi - hi0) >> shift, BINS - 1))
    hist = np.bincount(bins, minlength=BINS)
    incl = np.cumsum(hist)
    for attempt in range(2):
        fits = np.nonzero(incl <= CAP)[0]
        cut_full = int(fits[-1]) if len(fits) else -1
        reach = np.nonzero(incl >= FIRST)[0]
        first_bin = int(reach[0]) if len(reach) else BINS
        cut = first_bin if (attempt == 0 and first_bin <= cut_full) else cut_full
        if cut < 0 or (hist[:cut + 1] > BIN_MAX).any():
            return "v1", None
        Pn = int(incl[cut])
        if Pn == 0:
            return "v1", None
        sel = np.nonzero(bins <= cut)[0]
        order = sel[np.argsort(key[sel], kind="stable")]                 # the prefix in global order (keys are unique)
        cls_p = classes[order]
        kept = np.zeros(Pn, dtype=bool)
        ok = ((boxes[order, 2] - boxes[order, 0]) >= np.float32(0.01)) & ((boxes[order, 3] - boxes[order, 1]) >= np.float32(0.01))
        members = {c: np.nonzero(cls_p == c)[0] for c in np.unique(cls_p)}      # ranks of each class, ascending
        cur = {c: 0 for c in members}
        kidx = {c: [] for c in members}
        total = 0
        wave_end = WAVE
        while True:
            for c, ranks in members.items():
                while cur[c] < len(ranks) and len(kidx[c]) < max_det and ranks[cur[c]] < wave_end:
                    r = ranks[cur[c]]
                    cur[c] += 1
                    if not ok[r]:
                        continue
                    b = boxes[order[r]]
                    sup = False
                    for kr in kidx[c]:
                        kb = boxes[order[kr]]
                        w = np.float32(min(kb[2], b[2])) - np.float32(max(kb[0], b[0]))
                        h = np.float32(min(kb[3], b[3])) - np.float32(max(kb[1], b[1]))
                        if w > 0 and h > 0:
                            inter = np.float32(w) * np.float32(h)
                            ak = np.float32(kb[2] - kb[0]) * np.float32(kb[3] - kb[1])
                            ab = np.float32(b[2] - b[0]) * np.float32(b[3] - b[1])
                            if np.float32(inter / np.float32(np.float32(ak + ab) - inter)) > np.float32(nms_thr):
                                sup = True
                                break
                    if not sup:
                        kidx[c].append(r)
                        kept[r] = True
                        total += 1
            if total >= max_det or wave_end >= Pn:
                break
            wave_end += WAVE
        more = total < max_det and Pn < K
        if not more:
            return "ok", order[np.nonzero(kept)[0][:max_det]]
        if attempt == 1 or Pn >= CAP:
            return "v1", None
    return "v1", None


def reference(scores, classes, anchors, boxes, nms_thr, max_det):
    """models.py:193-240 with the documented tie rule: per-class torchvision-style NMS, then score desc / class asc / anchor asc."""
    keep = []
    for c in np.unique(classes):
        idx = np.nonzero(classes == c)[0]
        idx = idx[np.lexsort((anchors[idx], -scores[idx].astype(np.float64)))]      # score desc, anchor asc (stable order)
        small = ((boxes[idx, 2] - boxes[idx, 0]) >= np.float32(0.01)) & ((boxes[idx, 3] - boxes[idx, 1]) >= np.float32(0.01))
        idx = idx[small]
        k = O.nms_keep(torch.from_numpy(boxes[idx]), torch.from_numpy(scores[idx]), nms_thr).numpy()
        keep.append(idx[k])
    keep = np.concatenate(keep) if keep else np.zeros((0,), dtype=np.int64)
    order = np.lexsort((anchors[keep], classes[keep], -scores[keep].astype(np.float64)))
    return keep[order][:max_det]


def _case(rng, K, C, n_clusters, tie_levels=None, big=False):
    A = 50000
    anchors = rng.choice(A, size=K, replace=False).astype(np.int64)
    centers = rng.uniform(20, 480, size=(n_clusters, 2)).astype(np.float32)
    which = rng.integers(0, n_clusters, size=K)
    cls_of_cluster = rng.integers(0, C, size=n_clusters)
    classes = np.where(rng.random(K) < 0.8, cls_of_cluster[which], rng.integers(0, C, size=K)).astype(np.int64)
    ctr = centers[which] + rng.normal(0, 6 if not big else 60, size=(K, 2)).astype(np.float32)
    wh = rng.uniform(10, 60, size=(K, 2)).astype(np.float32)
    boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], 1).astype(np.float32)
    scores = (0.05 + 0.95 * rng.random(K) ** 4).astype(np.float32)
    scores = np.maximum(scores, np.nextafter(np.float32(0.05), np.float32(1)))
    if tie_levels:
        scores = (np.round(scores * tie_levels) / tie_levels).astype(np.float32)
        scores = np.maximum(scores, np.float32(0.0625))
    return scores, classes, anchors, boxes, A


@pytest.mark.parametrize("seed", range(6))
def test_lazy2_model_matches_reference_semantics(seed):
    rng = np.random.default_rng(seed)
    cases = [
        _case(rng, 300, 7, 12),                      # fewer candidates than the short prefix
        _case(rng, 3000, 20, 40),                    # short prefix suffices
        _case(rng, 9000, 80, 60),                    # config-2-like
        _case(rng, 6000, 3, 5),                      # few classes, few clusters: nearly everything suppressed -> long prefix / v1
        _case(rng, 5000, 80, 400, big=True),         # little overlap: max_det reached within the first wave
        _case(rng, 5000, 10, 30, tie_levels=64),     # heavy ties -> bins above LZ2_BIN_MAX -> hand-over
        _case(rng, 2500, 5, 8, tie_levels=4096),     # moderate ties inside bins (ranking inside a bin)
    ]
    handed_over = 0
    for k, (scores, classes, anchors, boxes, A) in enumerate(cases):
        for max_det in (100, 17):
            status, got = lazy2_model(scores, classes, anchors, boxes, 0.05, 0.5, max_det, A)
            if status == "v1":
                handed_over += 1
                continue
            want = reference(scores, classes, anchors, boxes, 0.5, max_det)
            assert np.array_equal(got, want), (seed, k, max_det, len(got), len(want))
    assert handed_over < 2 * len(cases)              # the fast path must actually be exercised
