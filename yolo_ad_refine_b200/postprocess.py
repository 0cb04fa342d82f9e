"""Drop-in for ultralytics.utils.ops.non_max_suppression (utils/ops.py:163-312) backed by the batched CUDA NMS in libyad.so."""
import torch

from . import ops

_WS = {}


def _workspace(nbytes, device):
    """Scratch for one yad_nms call.  While a CUDA graph is being captured every call gets its OWN tensor from the capture's private pool (it lives
    and dies with the graph, so a replay can never scribble on memory the caching allocator has handed to somebody else); eager calls share one
    buffer per (device, stream) -- kernels of one stream are ordered, and a buffer that is outgrown is released in stream order."""
    if torch.cuda.is_current_stream_capturing():
        return torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
    key = (device.index if device.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(device).cuda_stream)
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


def nms_raw(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False, max_det=300, nc=0,
            max_nms=30000, max_wh=7680, ws=None):
    """Device-side NMS without the host read-back: returns (out (B,max_det,6) fp32, out_idx (B,max_det,2) int32, count (B,) int32).
    ws: optional caller-owned uint8 workspace of at least ops.nms_workspace_bytes(...) bytes."""
    assert prediction.is_cuda, "the YOLO-AD-Refine NMS runs on the GPU only (no CPU fallback)"
    if prediction.dtype != torch.float32 or not prediction.is_contiguous():
        prediction = prediction.float().contiguous()
    bs, ch, n = prediction.shape
    nc = nc or (ch - 4)
    assert ch == 4 + nc, "mask coefficients (nm > 0) are not part of the YOLO-AD-Refine detection path"
    dev = prediction.device
    cm = None
    if classes is not None:
        cm = torch.zeros(nc, dtype=torch.uint8)
        cm[torch.as_tensor(classes, dtype=torch.long)] = 1
        cm = cm.to(dev)
    out = torch.empty((bs, max_det, 6), dtype=torch.float32, device=dev)
    out_idx = torch.empty((bs, max_det, 2), dtype=torch.int32, device=dev)
    count = torch.zeros((bs,), dtype=torch.int32, device=dev)
    need = ops.nms_workspace_bytes(bs, n, nc, multi_label and nc > 1, max_nms)
    if ws is None:
        ws = _workspace(need, dev)
    assert ws.is_cuda and ws.dtype == torch.uint8 and ws.numel() >= need, "nms_raw: workspace too small"
    ops.nms(prediction, float(conf_thres), float(iou_thres), cm, agnostic, multi_label, max_det, max_nms, float(max_wh), out, out_idx, count, ws)
    return out, out_idx, count


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False, labels=(),
                        max_det=300, nc=0, max_time_img=0.05, max_nms=30000, max_wh=7680, in_place=True, rotated=False,
                        return_idx=False):
    """Same signature and return value as the reference: list (length B) of (k, 6) tensors [x1, y1, x2, y2, conf, cls].
    Differences, all documented in DESIGN.md: GPU only; `labels` (autolabelling), `rotated` and mask coefficients are not on this
    path and raise; the reference's wall-clock early exit (`max_time_img`) does not exist; `prediction` is never modified."""
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    if isinstance(prediction, (list, tuple)):  # (inference_out, loss_out), ops.py:215-216
        prediction = prediction[0]
    if labels or rotated:
        raise NotImplementedError("labels= / rotated= are outside the YOLO-AD-Refine detection path")
    if prediction.shape[-1] == 6:  # end-to-end model output (ops.py:220-224)
        output = [pred[pred[:, 4] > conf_thres][:max_det] for pred in prediction]
        if classes is not None:
            cl = torch.tensor(classes, device=prediction.device)
            output = [pred[(pred[:, 5:6] == cl).any(1)] for pred in output]
        return output
    out, out_idx, count = nms_raw(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, max_det, nc, max_nms, max_wh)
    counts = count.tolist()  # the one host read-back the list-of-tensors API requires
    res = [out[i, :k] for i, k in enumerate(counts)]
    if return_idx:
        return res, [out_idx[i, :k] for i, k in enumerate(counts)]
    return res
