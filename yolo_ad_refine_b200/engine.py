"""RefineEngine: the whole inference hot path (forward -> fused decode -> batched NMS) for one GPU, captured in a CUDA graph.

One process per GPU; inference shards by batch with no collective, so multi-GPU inference is N independent engines.
"""
import torch

from . import functional as Fn
from . import ops
from .postprocess import nms_raw
from .weights import prepare


class RefineEngine:
    """use_graph: the hot path is captured in CUDA graphs.  pipeline_nms: forward + decode and NMS are two graphs on two streams with
    double-buffered predictions, so the NMS of batch i (one CTA per image: a low-occupancy tail) runs under the forward pass of batch i + 1.
    step() then only ENQUEUES a batch; join() (called by forward / detect / detect_many) orders the current stream after its results.
    overlap_batches (with pipeline_nms): the two buffer sets also own their input buffer and their forward stream, so TWO batches are in flight:
    the latency-bound stretches of one batch (the 20 x 20 maps of layer 10, 52 launches at the launch floor) run beside the throughput-bound
    convolutions of the other.  Every batch is still computed exactly as before (same graphs, same arithmetic); only the order in which the GPU
    interleaves the kernels of consecutive batches changes."""

    def __init__(self, state_dict, batch, imgsz=640, dtype=torch.bfloat16, device="cuda", nc=80, reg_max=16, strides=(8, 16, 32),
                 use_graph=True, conv_impl=0, nms_args=None, input_u8=False, pipeline_nms=True, overlap_batches=True):
        if not torch.cuda.is_available():
            raise RuntimeError("RefineEngine needs a CUDA device: the YOLO-AD-Refine hot path has no CPU fallback")
        ops.lib()  # fail loudly now if libyad.so is missing
        self.device = torch.device(device)
        self.h, self.w = (imgsz, imgsz) if isinstance(imgsz, int) else imgsz
        self.batch, self.nc, self.reg_max, self.strides = batch, nc, reg_max, strides
        self.ctx = Fn.Ctx(prepare(state_dict, dtype, self.device), conv_impl)
        self.nms_args = dict(conf_thres=0.25, iou_thres=0.7, max_det=300)
        self.nms_args.update(nms_args or {})
        self.use_graph = use_graph
        self.pipeline_nms = pipeline_nms and use_graph
        self.overlap_batches = overlap_batches and self.pipeline_nms
        # buffer sets: 2 with one batch in flight (forward i + 1 beside NMS i); overlap_batches = True -> 2 in flight, an integer n -> n in flight
        self._nsets = (2 if overlap_batches is True else max(2, int(overlap_batches))) if self.overlap_batches else 2
        self._imgs = [torch.zeros((batch, 3, self.h, self.w), dtype=torch.uint8 if input_u8 else torch.float32, device=self.device)
                      for _ in range(self._nsets if self.overlap_batches else 1)]
        self.graph = None
        self.launches_per_step = None
        self._out = None
        self._i = 0

    @property
    def img(self):
        """the static input buffer the NEXT step() reads (with overlap_batches the two buffer sets alternate)"""
        return self._imgs[self._i % self._nsets] if self.overlap_batches else self._imgs[0]

    def fill_inputs(self, x):
        """the same batch into every input buffer (device-resident benchmarking: step() then needs no copy at all)"""
        for t in self._imgs:
            t.copy_(x, non_blocking=True)

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
        if self.pipeline_nms:
            self._nms_stream = torch.cuda.Stream(device=self.device)
            self._gf, self._gn, self._bufs = [], [], []
            ns = self._nsets
            self._ev_f = [torch.cuda.Event() for _ in range(ns)]
            self._ev_n = [torch.cuda.Event() for _ in range(ns)]
            self._ev_in = [torch.cuda.Event() for _ in range(ns)]
            self._fstreams = [torch.cuda.Stream(device=self.device) for _ in range(ns)] if self.overlap_batches else None
            for b in range(ns):  # buffer sets: predictions of batch i are read by its NMS while batch i + 1 is being computed
                gf = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gf):
                    y, feats = Fn.forward_model(self.ctx, self._imgs[b if self.overlap_batches else 0])
                gn = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gn):
                    det, det_idx, count = nms_raw(y, **self.nms_args)
                self._gf.append(gf)
                self._gn.append(gn)
                self._bufs.append((y, feats, det, det_idx, count))
                self._ev_n[b].record(torch.cuda.current_stream())
                self._ev_f[b].record(torch.cuda.current_stream())
            self.graph = self._gf[0]
            self._out = self._bufs[0]
        elif self.use_graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._out = self._run()
            self.graph = g

    def step(self):
        """forward + decode + NMS on the current contents of self.img; returns device tensors (static across calls when graphed).
        With pipeline_nms the batch is only enqueued: call join() before reading the results on the current stream."""
        if self._out is None:
            self._capture()
        if self.pipeline_nms:
            b = self._i % self._nsets
            self._i += 1
            main = torch.cuda.current_stream()
            if self.overlap_batches:
                # this set's input was written on the current stream; its forward runs on the set's own stream, beside the other set's batch
                fs = self._fstreams[b]
                self._ev_in[b].record(main)
                fs.wait_event(self._ev_in[b])
                fs.wait_event(self._ev_n[b])   # the NMS that last read this buffer set (two batches ago) is done
                with torch.cuda.stream(fs):
                    self._gf[b].replay()
                    self._ev_f[b].record(fs)
                # the caller may refill the OTHER set's input next; that set's previous forward (one batch ago) must have consumed it
                main.wait_event(self._ev_f[(b + 1) % self._nsets])
            else:
                main.wait_event(self._ev_n[b])       # the NMS that last read this buffer set (two batches ago) is done
                self._gf[b].replay()
                self._ev_f[b].record(main)
            with torch.cuda.stream(self._nms_stream):
                self._nms_stream.wait_event(self._ev_f[b])
                self._gn[b].replay()
                self._ev_n[b].record(self._nms_stream)
            self._last = b
            self._out = self._bufs[b]
        elif self.graph is not None:
            self.graph.replay()
        else:
            self._out = self._run()
        return self._out

    def last_prediction(self):
        """y (B, 4 + nc, N) of the most recent step (valid after join()); the detections of that step are the NMS of exactly this tensor"""
        return self._out[0]

    def join(self):
        """order the current stream after every batch enqueued so far (no-op without pipeline_nms)"""
        if self.pipeline_nms and self._i:
            main = torch.cuda.current_stream()
            for ev in self._ev_n:
                main.wait_event(ev)

    def forward(self, img):
        """img: (B, 3, H, W) float tensor in [0,1] (host or device).  Returns (y (B,4+nc,N) fp32, [raw (B,144,H,W) views])."""
        self.img.copy_(img, non_blocking=True)
        y, feats, *_ = self.step()
        self.join()
        return y, [f.nchw() for f in feats]

    def detect(self, img):
        """Full hot path.  Returns the reference's NMS output: list (B) of (k, 6) tensors [x1,y1,x2,y2,conf,cls]."""
        self.img.copy_(img, non_blocking=True)
        _, _, det, _, count = self.step()
        self.join()
        counts = count.tolist()
        return [det[i, :k] for i, k in enumerate(counts)]

    def detect_images(self, images):
        """`predict()` for a list of exactly `batch` HWC uint8 BGR images (what the reference's loaders hand to BasePredictor.preprocess):
        device LetterBox + BGR->RGB + CHW (yad_letterbox) straight into the engine's static uint8 input -> forward -> decode -> NMS -> scale_boxes
        back to every source image's own coordinates (yad_scale_boxes; models/yolo/detect/predict.py:36-41).  Returns the list of (k, 6) tensors
        [x1, y1, x2, y2, conf, cls] the reference wraps into `Results`.  Needs an engine built with input_u8=True.  The list must fill the engine's
        batch: MLCA's global branch pools over the batch axis (block.py:1575-1579), so a partly stale static input would change every image's result."""
        if self.img.dtype != torch.uint8:
            raise RuntimeError("detect_images needs an engine built with input_u8=True")
        if len(images) != self.batch:
            raise ValueError(f"expected {self.batch} images (the engine's batch), got {len(images)}")
        if not hasattr(self, "_pre"):
            from .preprocess import DevicePreprocessor
            self._pre = DevicePreprocessor((self.h, self.w), stride=max(self.strides), auto=False, device=self.device)
        pb = self._pre(images, out=self.img)
        _, _, det, _, count = self.step()
        self.join()
        ops.scale_boxes(det, count, pb.desc)
        counts = count.tolist()
        return [det[i, :k] for i, k in enumerate(counts)]

    def detect_many(self, host_batches, det_host=None, cnt_host=None):
        """Streaming end-to-end path: yields, per host batch (pinned (B,3,H,W) uint8 or float tensor matching the engine input), the pinned host
        tensors (detections (B, max_det, 6), counts (B,)).  The host->device copy of batch i+1 runs on a copy stream while batch i computes
        (double-buffered staging), so the steady-state step time is max(copy, compute) instead of their sum."""
        if self._out is None:
            self._capture()
        main = torch.cuda.current_stream()
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._staging = [torch.empty_like(self.img) for _ in range(2)]
            self._ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._free = [torch.cuda.Event(), torch.cuda.Event()]
            for e in self._free:
                e.record(main)
        max_det = self.nms_args["max_det"]
        det_host = det_host if det_host is not None else torch.empty((self.batch, max_det, 6), dtype=torch.float32).pin_memory()
        cnt_host = cnt_host if cnt_host is not None else torch.empty((self.batch,), dtype=torch.int32).pin_memory()

        def enqueue(i, hb):
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(self._free[i % 2])
                self._staging[i % 2].copy_(hb, non_blocking=True)
                self._ready[i % 2].record(self._copy_stream)

        it = iter(host_batches)
        nxt = next(it, None)
        i = 0
        pending = None
        lag = []
        if nxt is not None:
            enqueue(0, nxt)
        while nxt is not None:
            nxt = next(it, None)
            if nxt is not None:
                enqueue(i + 1, nxt)
            main.wait_event(self._ready[i % 2])
            self.img.copy_(self._staging[i % 2], non_blocking=True)  # device-to-device into the graph's static input
            self._free[i % 2].record(main)
            _, _, det, _, count = self.step()
            if self.overlap_batches:
                # two batches in flight: the device -> host copy of batch i is enqueued right behind its NMS (pinned ring of three, so the host
                # may still be reading batch i - 2 while batch i - 1 lands), and the host only ever WAITS for batch i - 2 -- it keeps a whole
                # batch of enqueued work ahead of the GPU.  Results arrive in order, two batches late; the tail is flushed after the loop.
                nr = self._nsets + 1
                if not hasattr(self, "_ring"):
                    self._ring = [(torch.empty_like(det_host).pin_memory(), torch.empty_like(cnt_host).pin_memory(), torch.cuda.Event()) for _ in range(nr)]
                dh, ch, ev = self._ring[i % nr]
                with torch.cuda.stream(self._nms_stream):
                    dh.copy_(det, non_blocking=True)
                    ch.copy_(count, non_blocking=True)
                    ev.record(self._nms_stream)
                lag.append(i % nr)
                if len(lag) > self._nsets:
                    yield self._take(lag.pop(0), det_host, cnt_host)
            elif self.pipeline_nms:
                # software pipeline: batch i is only enqueued here; the detections handed to the caller are those of batch i - 1, read back on
                # the NMS stream while batch i computes (results arrive in order, one batch late; the last one is flushed after the loop)
                if pending is not None:
                    yield self._read_back(pending, det_host, cnt_host)
                pending = self._last
            else:
                det_host.copy_(det, non_blocking=True)
                cnt_host.copy_(count, non_blocking=True)
                main.synchronize()  # the caller reads this batch's detections on the host
                yield det_host, cnt_host
            i += 1
        while lag:
            yield self._take(lag.pop(0), det_host, cnt_host)
        if pending is not None:
            yield self._read_back(pending, det_host, cnt_host)

    def _take(self, slot, det_host, cnt_host):
        """wait for ring slot `slot` (its NMS and device -> host copy) and hand its rows to the caller's pinned tensors"""
        dh, ch, ev = self._ring[slot]
        ev.synchronize()
        det_host.copy_(dh)
        cnt_host.copy_(ch)
        return det_host, cnt_host

    def _read_back(self, b, det_host, cnt_host):
        """D2H copy of buffer set b's detections on the NMS stream (ordered after its NMS), then a host wait on that stream only"""
        _, _, det, _, count = self._bufs[b]
        with torch.cuda.stream(self._nms_stream):
            self._nms_stream.wait_event(self._ev_n[b])
            det_host.copy_(det, non_blocking=True)
            cnt_host.copy_(count, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._nms_stream)
        done.synchronize()
        return det_host, cnt_host
