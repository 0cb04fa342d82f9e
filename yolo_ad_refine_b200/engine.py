"""RefineEngine: the whole inference hot path (forward -> fused decode -> batched NMS) for one GPU, captured in a CUDA graph.

One process per GPU; inference shards by batch with no collective, so multi-GPU inference is N independent engines.
"""
import torch

from . import functional as Fn
from . import ops
from .postprocess import nms_raw
from .weights import prepare


class RefineEngine:
    def __init__(self, state_dict, batch, imgsz=640, dtype=torch.bfloat16, device="cuda", nc=80, reg_max=16, strides=(8, 16, 32),
                 use_graph=True, conv_impl=0, nms_args=None, input_u8=False):
        if not torch.cuda.is_available():
            raise RuntimeError("RefineEngine needs a CUDA device: the YOLO-AD-Refine hot path has no CPU fallback")
        ops.lib()  # fail loudly now if libyad.so is missing
        self.device = torch.device(device)
        self.h, self.w = (imgsz, imgsz) if isinstance(imgsz, int) else imgsz
        self.batch, self.nc, self.reg_max, self.strides = batch, nc, reg_max, strides
        self.ctx = Fn.Ctx(prepare(state_dict, dtype, self.device), conv_impl)
        self.nms_args = dict(conf_thres=0.25, iou_thres=0.7, max_det=300)
        self.nms_args.update(nms_args or {})
        self.img = torch.zeros((batch, 3, self.h, self.w), dtype=torch.uint8 if input_u8 else torch.float32, device=self.device)
        self.graph = None
        self.launches_per_step = None
        self._out = None
        self.use_graph = use_graph

    # -- one eager pass of the hot path on the static input buffer
    def _run(self):
        y, feats = Fn.forward_model(self.ctx, self.img)
        det, det_idx, count = nms_raw(y, **self.nms_args)
        return y, feats, det, det_idx, count

    def _capture(self):
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):  # warm-up: populates the weight cache and the allocator
                before = ops.LAUNCHES
                self._out = self._run()
                self.launches_per_step = ops.LAUNCHES - before
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if self.use_graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._out = self._run()
            self.graph = g

    def step(self):
        """forward + decode + NMS on the current contents of self.img; returns device tensors (static across calls when graphed)."""
        if self._out is None:
            self._capture()
        if self.graph is not None:
            self.graph.replay()
        else:
            self._out = self._run()
        return self._out

    def forward(self, img):
        """img: (B, 3, H, W) float tensor in [0,1] (host or device).  Returns (y (B,4+nc,N) fp32, [raw (B,144,H,W) views])."""
        self.img.copy_(img, non_blocking=True)
        y, feats, *_ = self.step()
        return y, [f.nchw() for f in feats]

    def detect(self, img):
        """Full hot path.  Returns the reference's NMS output: list (B) of (k, 6) tensors [x1,y1,x2,y2,conf,cls]."""
        self.img.copy_(img, non_blocking=True)
        _, _, det, _, count = self.step()
        counts = count.tolist()
        return [det[i, :k] for i, k in enumerate(counts)]
