"""Deterministic synthetic weights and inputs (shared by bench.py, tests/ and oracle/; no arithmetic of the hot path lives here).

The reference ships no checkpoints (all .pt are missing), so every parity check runs on a synthetic
state dict.  It is generated from the committed key/shape list `tests/golden/state_spec.json` (dumped from the live
reference model by oracle/gen_golden.py) with numpy's frozen legacy RandomState, keyed per parameter NAME, so that the
container that generated the golden fixtures and the GPU box (where the reference tree is absent) build bit-identical weights.
Fresh-init values would make many ops an identity (EDFFN.fft = 1, Fusion weights = 1, BN mean 0 / var 1, ...), so every
tensor is drawn from a non-trivial distribution.
"""
import json
import os
import zlib

import numpy as np

SPEC_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "state_spec.json")


def _rs(seed, key):
    return np.random.RandomState((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def _draw(key, shape, seed):
    rs = _rs(seed, key)
    n = int(np.prod(shape)) if len(shape) else 1
    leaf = key.rsplit(".", 1)[-1]
    parent = key.rsplit(".", 2)[-2] if key.count(".") >= 1 else ""

    def normal(mu, sd):
        return (mu + sd * rs.standard_normal(n)).reshape(shape)

    def uniform(lo, hi):
        return rs.uniform(lo, hi, n).reshape(shape)

    if leaf == "num_batches_tracked":
        return np.zeros(shape, np.int64)
    if leaf == "running_mean":
        return normal(0.0, 0.1)
    if leaf == "running_var":
        return uniform(0.5, 1.5)
    if key.endswith("dfl.conv.weight"):  # frozen arange projection (nn/modules/block.py:70-74)
        return np.arange(n, dtype=np.float64).reshape(shape)
    if leaf == "fft":
        return normal(1.0, 0.1)
    if leaf == "fusion_weight":
        return uniform(0.5, 1.5)
    if leaf == "stage_attention":
        return uniform(0.2, 0.5)
    if leaf in ("residual_weight1", "residual_weight2"):
        return uniform(0.3, 0.6)
    if leaf == "alphas":
        return uniform(0.3, 1.5)
    if leaf == "scale_weights":
        return uniform(0.5, 1.5)
    if leaf == "temps":
        return uniform(0.5, 1.5)
    if leaf == "scale":  # head Scale modules
        return uniform(0.8, 1.2)
    is_norm = parent in ("bn", "norm", "gn", "bn1") or key.endswith("conv1x1.1.weight") or key.endswith("conv1x1.1.bias") \
        or ".dyt" in key and leaf in ("weight", "bias") and "importance_gate" not in key
    if is_norm:
        return uniform(0.5, 1.5) if leaf == "weight" else normal(0.0, 0.1)
    if leaf in ("bias", "in_proj_bias"):
        return normal(0.0, 0.1)
    if len(shape) >= 2:  # conv / linear / conv1d weights
        fan_in = int(np.prod(shape[1:]))
        if "attention.conv" in key:  # MLCA 1-D convs over (1,1,k)
            return normal(0.0, 0.5)
        return normal(0.0, 1.0 / np.sqrt(fan_in))
    return normal(0.0, 0.1)


def load_spec(path=SPEC_PATH):
    with open(path) as f:
        return json.load(f)


def make_state_dict_np(seed=1, spec=None):
    """-> dict name -> numpy array (float32 / int64) following the reference's state-dict key names."""
    spec = load_spec() if spec is None else spec
    out = {}
    for key, shape, dtype in spec:
        a = _draw(key, tuple(shape), seed)
        out[key] = a.astype(np.int64 if "int" in dtype else np.float32)
    return out


def make_state_dict(seed=1, spec=None):
    import torch
    return {k: torch.from_numpy(v.copy()) for k, v in make_state_dict_np(seed, spec).items()}


def state_checksum(sd_np):
    return float(sum(np.asarray(v, np.float64).sum() for k, v in sorted(sd_np.items())))


def make_images(batch, h, w, seed=2):
    """(B,3,H,W) float32 in [0,1)."""
    rs = np.random.RandomState(seed)
    return rs.random_sample((batch, 3, h, w)).astype(np.float32)


# ----- post-processing / training inputs -------------------------------------------------------------------------------------
def make_head_logits(batch, n_anchors, nc=80, reg_max=16, seed=0, cls_mu=-6.0, cls_sd=1.5):
    """raw head output (B, 4*reg_max+nc, N): box logits 2*randn, class logits mu+sd*randn."""
    rs = np.random.RandomState(seed)
    raw = rs.standard_normal((batch, 4 * reg_max + nc, n_anchors)).astype(np.float32)
    raw[:, :4 * reg_max] *= np.float32(2.0)
    raw[:, 4 * reg_max:] = np.float32(cls_mu) + np.float32(cls_sd) * raw[:, 4 * reg_max:]
    return raw


def make_predictions(batch, n_anchors, nc=80, seed=0, img=640.0, p_hot=0.04, cluster=True):
    """Decoded predictions (B, 4+nc, N) xywh px + scores in (0,1) for NMS tests: boxes cluster around a few centres so that
    suppression actually happens; a fraction p_hot of anchors carries one high class score."""
    rs = np.random.RandomState(seed)
    pred = np.empty((batch, 4 + nc, n_anchors), np.float32)
    for b in range(batch):
        if cluster:
            k = 12
            cen = rs.uniform(60, img - 60, (k, 2))
            wh0 = rs.uniform(30, 160, (k, 2))
            a = rs.randint(0, k, n_anchors)
            xy = cen[a] + rs.standard_normal((n_anchors, 2)) * 8.0
            wh = wh0[a] * np.exp(rs.standard_normal((n_anchors, 2)) * 0.15)
        else:
            xy = rs.uniform(0, img, (n_anchors, 2))
            wh = rs.uniform(8, 200, (n_anchors, 2))
        pred[b, 0:2] = xy.T
        pred[b, 2:4] = wh.T
        sc = rs.uniform(0.0, 0.05, (nc, n_anchors))
        hot = rs.random_sample(n_anchors) < p_hot
        cls = rs.randint(0, nc, n_anchors)
        val = rs.uniform(0.1, 0.99, n_anchors)
        idx = np.nonzero(hot)[0]
        sc[cls[idx], idx] = val[idx]
        pred[b, 4:] = sc
    return pred


def make_targets(batch, seed=3, max_per_img=32, nc=80, empty_images=(0,)):
    """batch dict pieces: batch_idx (M,), cls (M,1), bboxes (M,4) normalised xywh."""
    rs = np.random.RandomState(seed)
    bi, cl, bb = [], [], []
    for b in range(batch):
        n = 0 if b in empty_images else int(rs.randint(1, max_per_img + 1))
        cxcy = rs.uniform(0.2, 0.8, (n, 2))
        wh = rs.uniform(0.02, 0.32, (n, 2))
        bi.append(np.full((n,), b, np.float32))
        cl.append(rs.randint(0, nc, (n, 1)).astype(np.float32))
        bb.append(np.concatenate([cxcy, wh], 1).astype(np.float32))
    return np.concatenate(bi), np.concatenate(cl), np.concatenate(bb)
