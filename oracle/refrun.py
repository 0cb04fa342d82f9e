"""TEST / BASELINE INFRASTRUCTURE -- runs the UNMODIFIED reference (its own DetectionModel, _predict_once, ops.non_max_suppression) on the CPU from
/root/reference or, where that is absent (GPU box), from the copy oracle/build_ref.py made under oracle/_ref.  Used by `bench.py --impl reference`
and by the plugin tests; never by product code."""
import os
import time

from . import ref_shims, synth


def available():
    return ref_shims.available()


def build_model(fuse=True):
    """the reference's DetectionModel for z-yaml/yolo11-701-YOLO-AD-Refine.yaml with the deterministic synthetic weights (oracle/synth.py)"""
    ref_shims.install()
    import torch
    from ultralytics.nn.tasks import DetectionModel
    yaml = os.path.join(ref_shims.REFERENCE_ROOT, "z-yaml", "yolo11-701-YOLO-AD-Refine.yaml")
    m = DetectionModel(yaml, ch=3, nc=80, verbose=False)
    m.load_state_dict(synth.make_state_dict(seed=1), strict=True)
    m.eval()
    if fuse:
        m.fuse(verbose=False)
    torch.set_num_threads(os.cpu_count() or 1)
    return m


def predict(m, img, conf_thres=0.25, iou_thres=0.7, max_det=300):
    """forward + decode + NMS through the reference's own code (engine/predictor.py:214-299 minus pre/post-processing)"""
    import torch
    from ultralytics.utils import ops as ref_ops
    m.model[-1].shape = None  # anchors are cached by shape (nn/modules/head.py:1184)
    with torch.inference_mode():
        y, _ = m(img)
        return ref_ops.non_max_suppression(y, conf_thres, iou_thres, max_det=max_det, max_time_img=1e9)


def time_steps(batch, imgsz, steps, warmup):
    """`steps` timed passes (after `warmup` untimed ones) of forward + decode + NMS on a batch of `batch` synthetic images; returns
    (seconds per step, threads used)"""
    import torch
    m = build_model()
    img = torch.from_numpy(synth.make_images(batch, imgsz, imgsz, seed=2))
    for _ in range(warmup):
        predict(m, img)
    t0 = time.perf_counter()
    for _ in range(steps):
        predict(m, img)
    return (time.perf_counter() - t0) / max(1, steps), torch.get_num_threads()
