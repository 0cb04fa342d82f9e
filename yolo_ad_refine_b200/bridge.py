"""Training through the REFERENCE's own API: `trainer.model(batch)` -> `.backward()` -> `optimizer_step()` of
engine/trainer.py:382-397, 580-588 with the whole step -- train()-mode forward, TaskAlignedAssigner, v8DetectionLoss and the complete backward --
executed by libyad.so.

The reference trains by autograd over its modules.  Here `DetectionModel.loss(batch)` (nn/tasks.py:290-302), which `BaseModel.forward` calls for
a dict input, is routed to a `TrainBridge`: the module parameters stay the master copy (the reference's optimizer, gradient clipping, EMA, DDP
hooks and checkpoints keep working on them unchanged); before each step they are copied into the TrainEngine's flat fp32 arena, the engine runs
forward + loss + backward, and the returned loss tensor carries an autograd node whose backward hands each parameter its gradient (a view of the
engine's flat gradient arena, scaled by the incoming gradient -- the trainer multiplies the loss by world_size, trainer.py:387).  BatchNorm running
statistics written by the kernels are copied back into the module buffers, `num_batches_tracked` is incremented as nn.BatchNorm2d does.

`plugin.install()` binds this in; nothing here is a CPU path: without a CUDA device the bridge raises.
"""
import math

import torch

from .trainer import TrainEngine


class _StepFn(torch.autograd.Function):
    """loss = f(parameters): forward has already been run by the engine; backward distributes the engine's gradients"""

    @staticmethod
    def forward(ctx, bridge, out4, *params):
        ctx.bridge = bridge
        return out4[3].clone()

    @staticmethod
    def backward(ctx, g):
        br = ctx.bridge
        flat = br.eng.tp.grad * g.to(br.eng.tp.grad.dtype)  # one launch; the views below are what autograd accumulates into .grad
        tp = br.eng.tp
        grads = []
        for k, p in zip(br.keys, br.params):
            if not p.requires_grad:
                grads.append(None)
                continue
            n = math.prod(tp.shape[k])
            grads.append(flat[tp.off[k]:tp.off[k] + n].view(tp.shape[k]).to(p.dtype))
        return (None, None, *grads)


class TrainBridge:
    def __init__(self, model, dtype=torch.bfloat16, conv_impl=0):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("yolo_ad_refine_b200: training runs on a CUDA device only (no CPU fallback); move the model to the GPU first")
        head = model.model[-1]
        args = getattr(model, "args", None)

        def hyp(k, d):
            if args is None:
                return d
            return float(args[k] if isinstance(args, dict) else getattr(args, k, d))

        sd = model.state_dict()
        self.eng = TrainEngine(sd, dtype=dtype, device=dev, conv_impl=conv_impl, nc=head.nc, reg_max=head.reg_max,
                               strides=tuple(float(s) for s in head.stride), gains=(hyp("box", 7.5), hyp("cls", 0.5), hyp("dfl", 1.5)))
        named_p, named_b = dict(model.named_parameters()), dict(model.named_buffers())
        tp = self.eng.tp
        self.keys = [k for k in tp.keys if k in named_p]
        missing = [k for k in tp.keys if k not in named_p]
        assert not missing, f"parameters of the libyad training graph that the model does not hold: {missing[:5]}"
        self.params = [named_p[k] for k in self.keys]
        self._pviews = [tp.p(k) for k in self.keys]
        self.buf_keys = [k for k in tp.buf_keys if k in named_b]
        self._bufs = [named_b[k] for k in self.buf_keys]
        self._bviews = [tp.buf(k).view(tp.shape[k]) for k in self.buf_keys]
        self._tracked = [b for k, b in named_b.items() if k.endswith("num_batches_tracked")]
        self.last_items = None

    def loss(self, batch):
        """the reference's `model.loss(batch)`: returns (loss * batch_size with an autograd node, detached loss items [box, cls, dfl])"""
        with torch.no_grad():
            torch._foreach_copy_(self._pviews, [p.detach().float() if p.dtype != torch.float32 else p.detach() for p in self.params])
            if self._bufs:
                torch._foreach_copy_(self._bviews, [b.float() for b in self._bufs])
        img = batch["img"]
        if not img.is_cuda:
            img = img.to(self.eng.device, non_blocking=True)
        if img.dtype not in (torch.uint8, torch.float32):
            img = img.float()
        out4 = self.eng.forward_backward(img, batch["batch_idx"], batch["cls"], batch["bboxes"])
        with torch.no_grad():
            if self._bufs:
                torch._foreach_copy_(self._bufs, [v.to(b.dtype) for v, b in zip(self._bviews, self._bufs)])
            for t in self._tracked:
                t += 1
        loss = _StepFn.apply(self, out4, *self.params)
        self.last_items = out4[:3].detach()
        return loss, self.last_items


def model_loss(orig_loss):
    """wrapper for `DetectionModel.loss`: training-mode calls without precomputed predictions go through the bridge; everything else (the
    validator's `model.loss(batch, preds)`, eval mode) keeps the reference's own route to the (libyad-backed) criterion"""

    def loss(self, batch, preds=None):
        if preds is not None or not self.training or not next(self.parameters()).is_cuda:
            return orig_loss(self, batch, preds)
        br = self.__dict__.get("_yad_bridge")
        if br is None:
            br = TrainBridge(self, dtype=self.__dict__.get("_yad_train_dtype", torch.bfloat16), conv_impl=self.__dict__.get("_yad_conv_impl", 0))
            self.__dict__["_yad_bridge"] = br  # not a submodule / parameter: keep it out of state_dict and .to()
        return br.loss(batch)

    loss._yad_wrapped = True
    return loss
