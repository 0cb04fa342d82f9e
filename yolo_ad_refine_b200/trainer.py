"""TrainEngine: one training step of YOLO-AD-Refine on one GPU -- train()-mode forward, v8DetectionLoss (TaskAlignedAssigner + CIoU/NWD + DFL +
SlideLoss-BCE), the full backward pass, gradient exchange, clip + SGD(nesterov) + EMA -- all as libyad.so kernels.

Mirrors engine/trainer.py:382-397 (forward, loss, backward) and :580-588 (optimizer_step: unscale, clip_grad_norm_(10), step, zero_grad, EMA).
Multi-GPU: one process per GPU; the reference wraps the model in DistributedDataParallel (trainer.py:273), which
broadcasts rank 0's BatchNorm buffers before each forward and averages gradients, and multiplies the loss by world_size (trainer.py:387), so
the applied gradient is the SUM over ranks.  Here the flat fp32 gradient arena is all-reduced (SUM) in ONE NCCL call.
"""
import torch

from . import ops
from . import training as T
from .loss import detection_loss_raw, pack_targets_static, preprocess_targets
from .tal import make_anchors
from .train_params import TrainParams


class TrainEngine:
    def __init__(self, state_dict, dtype=torch.bfloat16, device="cuda", conv_impl=0, nc=80, reg_max=16, strides=(8, 16, 32),
                 gains=(7.5, 0.5, 1.5), topk=10, lr=0.01, momentum=0.937, weight_decay=5e-4, process_group=None, world_size=1, optimizer="SGD"):
        if not torch.cuda.is_available():
            raise RuntimeError("TrainEngine needs a CUDA device: the YOLO-AD-Refine training path has no CPU fallback")
        ops.lib()
        self.device = torch.device(device)
        self.tp = TrainParams(state_dict, dtype, self.device)
        self.conv_impl, self.nc, self.reg_max, self.strides, self.gains, self.topk = conv_impl, nc, reg_max, strides, gains, topk
        self.lr, self.momentum, self.weight_decay, self.optimizer = lr, momentum, weight_decay, optimizer
        self.pg, self.world_size = process_group, world_size
        self._anchors = {}
        self.last = None

    def _anchor(self, shapes):
        if shapes not in self._anchors:
            a, s = make_anchors(list(shapes), list(self.strides))
            self._anchors[shapes] = (a.to(self.device), s.to(self.device))
        return self._anchors[shapes]

    def forward_backward(self, img, batch_idx, cls, bboxes, update_bn=True, keep=False, zero_grad=True, assign=None, static_n_max=None):
        """img fp32 (B, 3, H, W) in [0, 1] or uint8 in [0, 255], on the device; targets as in the reference's batch dict (utils/loss.py:443-446).
        Leaves the gradients in tp.grad (reference state-dict layout) and returns out4 = [box, cls, dfl, total * B] (device, fp32).
        zero_grad=False accumulates onto the gradients of the previous call (the reference's `accumulate = max(round(nbs / batch), 1)`,
        engine/trainer.py:305,373: several forward / backward passes per optimizer step)."""
        tp = self.tp
        if zero_grad:
            tp.zero_grad()
        else:  # the packed weight-gradient scratch of the previous call has already been folded into tp.grad
            tp.arena_g[:max(tp._top_t, 8)].zero_()
            tp.arena_gf[:max(tp._top_f, 8)].zero_()
        tp.pack()
        g = T.Graph(tp, self.conv_impl, update_bn)
        outs, layers = T.forward_model(g, img)
        B = img.shape[0]
        N = sum(o.h * o.w for o in outs)
        reg_ch = 4 * self.reg_max
        distri = torch.empty((B, N, reg_ch), dtype=torch.float32, device=self.device)
        logits = torch.empty((B, N, self.nc), dtype=torch.float32, device=self.device)
        a0 = 0
        for o in outs:
            ops.head_pack(o, a0, N, reg_ch, self.nc, distri, logits)
            a0 += o.h * o.w
        anchors, stride_t = self._anchor(tuple((o.h, o.w) for o in outs))
        if static_n_max is not None:  # CUDA-graph step: fixed-shape device tensors, no host synchronisation
            gt_labels, gt_bboxes, mask_gt = pack_targets_static(batch_idx, cls, bboxes, B, (img.shape[2], img.shape[3]), static_n_max)
        else:
            gt_labels, gt_bboxes, mask_gt = preprocess_targets(batch_idx, cls, bboxes, B, (img.shape[2], img.shape[3]), self.device)
        out4, gd, gs, aux = detection_loss_raw(distri, logits, anchors, stride_t, gt_labels, gt_bboxes, mask_gt, self.gains, self.topk, self.reg_max,
                                                  assign=assign)
        a0 = 0
        for o in outs:
            g.mark(o)
            ops.head_unpack(gd, gs, 1.0, a0, N, reg_ch, self.nc, g.grad(o))
            a0 += o.h * o.w
        g.backward()
        tp.unpack_grads()
        if update_bn and not torch.cuda.is_current_stream_capturing():
            tp.bump_batches_tracked()
        if keep:
            self.last = dict(outs=outs, layers=layers, aux=aux, graph=g)
        return out4

    def exchange(self):
        """DDP gradient exchange (sum over ranks) + rank-0 BatchNorm buffers"""
        if self.world_size > 1:
            from .parallel import exchange_gradients
            exchange_gradients(self.tp, self.pg)

    def step(self, img, batch_idx, cls, bboxes, lr=None):
        out4 = self.forward_backward(img, batch_idx, cls, bboxes)
        self.exchange()
        self.tp.optimizer_step(lr=self.lr if lr is None else lr, momentum=self.momentum, weight_decay=self.weight_decay, optimizer=self.optimizer)
        return out4

    # ---- CUDA-graph step ------------------------------------------------------------------------------------------------------------
    def capture(self, batch, imgsz, m_cap=None, n_max=32, img_dtype=torch.uint8):
        """Capture the whole step for a fixed batch geometry: static inputs (image batch, m_cap target rows padded with batch_idx = -1, at most
        n_max targets per image), forward + loss + backward [+ the gradient all-reduce] + clip / optimizer / EMA.  With one rank everything is
        ONE graph; with several the NCCL all-reduce runs between two graphs (forward + backward | optimizer), issued eagerly on the same
        stream -- a replay costs two launches and one collective call instead of ~1,450 kernel launches through ctypes.  Per-step scalars
        (learning rate, Adam bias correction, EMA decay) live in a device buffer (TrainParams.set_hyper)."""
        h, w = (imgsz, imgsz) if isinstance(imgsz, int) else imgsz
        m_cap = m_cap or batch * n_max
        dev = self.device
        self._s_img = torch.zeros((batch, 3, h, w), dtype=img_dtype, device=dev)
        self._s_bi = torch.full((m_cap,), -1.0, dtype=torch.float32, device=dev)
        self._s_cls = torch.zeros((m_cap, 1), dtype=torch.float32, device=dev)
        self._s_box = torch.zeros((m_cap, 4), dtype=torch.float32, device=dev)
        self._s_nmax, self._s_mcap = n_max, m_cap
        if self.optimizer == "AdamW":
            self.tp.ensure_adamw_state()
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):  # warm-up: weight layouts, allocator; BatchNorm buffers untouched, gradients discarded
                self.forward_backward(self._s_img, self._s_bi, self._s_cls, self._s_box, update_bn=False, static_n_max=n_max)
            self.tp.zero_grad()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._gA = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._gA):
            self._s_out4 = self.forward_backward(self._s_img, self._s_bi, self._s_cls, self._s_box, static_n_max=n_max)
            if self.world_size == 1:
                self.tp.optimizer_step_dev(self.optimizer)
        self._gB = None
        if self.world_size > 1:
            self._gB = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._gB, pool=self._gA.pool()):
                self.tp.optimizer_step_dev(self.optimizer)
        return self

    def load_static(self, img, batch_idx, cls, bboxes):
        """copy one batch into the captured step's static inputs (asynchronous on the current stream; host tensors should be pinned)"""
        m = batch_idx.numel()
        assert m <= self._s_mcap, f"{m} targets exceed the captured capacity {self._s_mcap}"
        self._s_img.copy_(img, non_blocking=True)
        self._s_bi.fill_(-1.0)
        if m:
            self._s_bi[:m].copy_(batch_idx.reshape(-1), non_blocking=True)
            self._s_cls[:m].copy_(cls.reshape(-1, 1), non_blocking=True)
            self._s_box[:m].copy_(bboxes.reshape(-1, 4), non_blocking=True)

    def step_graphed(self, img=None, batch_idx=None, cls=None, bboxes=None, lr=None):
        """one replayed step; with img=None the static inputs are used as they are.  Returns the static out4 tensor [box, cls, dfl, total * B]."""
        if img is not None:
            self.load_static(img, batch_idx, cls, bboxes)
        self.tp.set_hyper(lr=self.lr if lr is None else lr, momentum=self.momentum, weight_decay=self.weight_decay, optimizer=self.optimizer)
        self._gA.replay()
        if self._gB is not None:
            self.exchange()
            self._gB.replay()
        self.tp.bump_batches_tracked()
        return self._s_out4
