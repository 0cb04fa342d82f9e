"""helpers shared by the GPU parity tests"""
import numpy as np
import torch

from yolo_ad_refine_b200 import ops
from yolo_ad_refine_b200.ops import Act

DEV = "cuda"


def to_act(x, dtype=torch.float32):
    """(n,c,h,w) CPU/GPU tensor -> Act"""
    return Act.from_nchw(x.to(DEV), dtype)


def from_act(a, c=None):
    t = a.nchw().float().cpu()
    return t if c is None else t[:, :c]


def rel_err(got, ref):
    """max over elements of |got - ref| / (|ref| + mean|ref|): element-wise relative error with the mean magnitude as the floor"""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return float((np.abs(got - ref) / (np.abs(ref) + np.abs(ref).mean() + 1e-12)).max())


def tol(dtype):
    # fp32 kernels: 1e-3 relative (north_star); bf16 storage: 8 mantissa bits -> 2^-9 relative per rounding, a few roundings per op
    return 1e-3 if dtype == torch.float32 else 2e-2


DTYPES = [torch.float32, torch.bfloat16]
