"""Shared test helpers: golden loading, synthetic inputs, comparison utilities."""
import hashlib
import os

import numpy as np
import torch

import synth_data as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def digest(*tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def config_image(cid, image_index=0):
    """(batch dict, golden npz) for one image of a config; asserts the regenerated inputs are the
    ones the golden file was made from."""
    g = golden(f"config{cid}_img{image_index}.npz")
    b = S.make_batch(S.CONFIGS[cid], image_index, 1)
    t = b["targets"][0]
    assert digest(b["cls_preds"], b["bbox_preds"], t["boxes"], t["labels"]) == str(g["input_sha256"]), \
        "synthetic inputs differ from the ones the golden vectors were generated from"
    assert digest(b["anchors"]) == str(g["anchors_sha256"])
    return b, g


def to_cuda_targets(targets, dev="cuda"):
    return [{k: v.to(dev) for k, v in t.items()} for t in targets]


def rel_close(a, b, rtol=1e-5, atol=0.0):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return bool(torch.all((a - b).abs() <= atol + rtol * b.abs()))


def max_rel(a, b, floor=1e-30):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float(((a - b).abs() / b.abs().clamp_min(floor)).max()) if a.numel() else 0.0
