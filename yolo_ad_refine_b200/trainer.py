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
from .loss import detection_loss_raw, preprocess_targets
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

    def forward_backward(self, img, batch_idx, cls, bboxes, update_bn=True, keep=False, zero_grad=True, assign=None):
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
