"""Parameter store of the training path: fp32 master copy of every trainable tensor in ONE flat arena (reference state-dict layout, so a
state dict is a set of views), a matching flat gradient arena (what DDP all-reduces), momentum / EMA arenas, and the
kernel-layout copies (packed conv weights and their dgrad twins, padded biases, depthwise taps) produced once per step by a table-driven
permute kernel (yad_permute_pack) -- with the reverse accumulation of the packed weight gradients (yad_permute_unpack).

Optimizer semantics follow engine/trainer.py:580-588 (clip_grad_norm_(10) then SGD step) and :784-808 (three parameter groups), EMA
utils/torch_utils.py:530-541.  PyTorch is used for allocation and views only.
"""
import ctypes as C
import math
import re

import torch

from . import ops
from ._lib import YadPermuteEntry

_NORM_WEIGHT = re.compile(r"\.(bn|gn|norm|bn1)\.weight$|\.conv1x1\.1\.weight$")
ALIGN = 8  # elements


def _pad8(n):
    return (n + 7) // 8 * 8


def optimizer_group(key):
    """engine/trainer.py:796-803: 2 = bias (no decay), 1 = weight of a normalisation layer (no decay), 0 = everything else (decay)."""
    if "bias" in key:
        return 2
    return 1 if _NORM_WEIGHT.search(key) else 0


def is_frozen(key):
    """parameters the reference never gives a gradient: DFL's fixed projection (requires_grad False, block.py:72) and AdaptiveDynamicTanh's
    unused scale_weights (block.py:2511); SGD skips them (grad is None)."""
    return key.endswith("dfl.conv.weight") or key.endswith(".scale_weights")


class PackedW:
    """kernel-layout weight: .w (activation dtype, [n0p][taps][n2p]), .gw fp32 gradient in the same layout (or None)"""
    __slots__ = ("w", "gw", "kh", "kw", "cin", "cout")

    def __init__(self, w, gw, kh, kw, cout, cin):
        self.w, self.gw, self.kh, self.kw, self.cout, self.cin = w, gw, kh, kw, cout, cin


class TrainParams:
    def __init__(self, sd, dtype=torch.bfloat16, device="cuda"):
        self.dtype, self.device = dtype, torch.device(device)
        self.keys, self.shape, self.off = [], {}, {}
        self.buf_keys, self.buf_off = [], {}
        total = btotal = 0
        self.other = {}
        for k, v in sd.items():
            if not isinstance(v, torch.Tensor):
                continue
            if k.endswith(("running_mean", "running_var")):
                self.buf_keys.append(k)
                self.buf_off[k] = btotal
                self.shape[k] = tuple(v.shape)
                btotal += _pad8(v.numel())
            elif v.dtype.is_floating_point:
                self.keys.append(k)
                self.off[k] = total
                self.shape[k] = tuple(v.shape)
                total += _pad8(v.numel())
            else:
                self.other[k] = v.clone()  # num_batches_tracked
        self.total, self.btotal = total, btotal
        dev = self.device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.mom = torch.zeros(total, dtype=torch.float32, device=dev)
        self.ema = torch.zeros(total, dtype=torch.float32, device=dev)
        self.bufs = torch.zeros(max(btotal, 8), dtype=torch.float32, device=dev)
        self.ema_bufs = torch.zeros(max(btotal, 8), dtype=torch.float32, device=dev)
        group = torch.full((total,), 255, dtype=torch.uint8)
        host = torch.zeros(total, dtype=torch.float32)
        for k in self.keys:
            n = math.prod(self.shape[k])
            host[self.off[k]:self.off[k] + n] = sd[k].detach().float().reshape(-1)
            if not is_frozen(k):
                group[self.off[k]:self.off[k] + n] = optimizer_group(k)
        bhost = torch.zeros(max(btotal, 8), dtype=torch.float32)
        for k in self.buf_keys:
            n = math.prod(self.shape[k])
            bhost[self.buf_off[k]:self.buf_off[k] + n] = sd[k].detach().float().reshape(-1)
        self.flat.copy_(host)
        self.ema.copy_(host)
        self.bufs.copy_(bhost)
        self.ema_bufs.copy_(bhost)
        self.group = group.to(dev)
        # kernel-layout arenas (bump-allocated on first use; generous fixed capacity: every conv weight twice + padding)
        cap = 3 * total + (1 << 16)
        self.arena_t = torch.zeros(cap, dtype=dtype, device=dev)
        self.arena_g = torch.zeros(cap, dtype=torch.float32, device=dev)   # packed weight gradients (same offsets as arena_t)
        self.arena_f = torch.zeros(total // 4 + (1 << 14), dtype=torch.float32, device=dev)
        self.arena_gf = torch.zeros_like(self.arena_f)
        self._top_t = self._top_f = 0
        self._entries, self._packed = [], {}
        self._tables = None
        self.norm_sq = torch.zeros(1, dtype=torch.float64, device=dev)
        # per-step scalars of the graph-captured optimizer (set_hyper / optimizer_step_dev).  Allocated HERE, never inside a capture: a tensor
        # created while capturing lives in the graph's pool and its zero-fill would be replayed with every step
        self._hyper = torch.zeros(16, dtype=torch.float32, device=dev)
        # a small ring of pinned staging rows: the asynchronous upload of step k may still be pending when the host prepares step k + 1
        self._hyper_host = torch.zeros(8, 16, dtype=torch.float32).pin_memory() if dev.type == "cuda" else torch.zeros(8, 16)
        self.steps = 0
        self.ema_updates = 0

    # ---- views ------------------------------------------------------------------------------------------------------
    def _view(self, arena, key):
        n = math.prod(self.shape[key])
        return arena[self.off[key]:self.off[key] + n].view(self.shape[key])

    def p(self, key):
        return self._view(self.flat, key)

    def g(self, key):
        return self._view(self.grad, key)

    def buf(self, key):
        n = math.prod(self.shape[key])
        return self.bufs[self.buf_off[key]:self.buf_off[key] + n]

    def has(self, key):
        return key in self.off

    def state_dict(self, ema=False):
        """reference-layout state dict (copies)"""
        out = {}
        src, bsrc = (self.ema, self.ema_bufs) if ema else (self.flat, self.bufs)
        for k in self.keys:
            out[k] = self._view(src, k).clone()
        for k in self.buf_keys:
            n = math.prod(self.shape[k])
            out[k] = bsrc[self.buf_off[k]:self.buf_off[k] + n].clone().view(self.shape[k])
        out.update({k: v.clone() for k, v in self.other.items()})
        return out

    def checkpoint(self):
        """everything a resumed run needs (the reference's `last.pt` keeps model, EMA, optimizer state and the EMA update count,
        engine/trainer.py:507-540): flat fp32 arenas as CPU tensors + counters; the key / offset table is implied by the state-dict order"""
        ck = dict(flat=self.flat.cpu(), mom=self.mom.cpu(), ema=self.ema.cpu(), bufs=self.bufs.cpu(), ema_bufs=self.ema_bufs.cpu(),
                  steps=self.steps, ema_updates=self.ema_updates, keys=list(self.keys), total=self.total,
                  other={k: v.cpu() for k, v in self.other.items()}, optimizer=getattr(self, "optimizer_name", None))
        if hasattr(self, "mom2"):  # AdamW's second moment: without it a resumed run divides by sqrt(~0) on its first steps
            ck["mom2"] = self.mom2.cpu()
        return ck

    def load_checkpoint(self, ck):
        assert ck["keys"] == self.keys and ck["total"] == self.total, "checkpoint was written for a different parameter layout"
        for name in ("flat", "mom", "ema", "bufs", "ema_bufs"):
            getattr(self, name).copy_(ck[name])
        if ck.get("mom2") is not None:
            self.mom2 = ck["mom2"].to(self.device)
        elif hasattr(self, "mom2"):
            del self.mom2
        if ck.get("optimizer") is not None:
            cur = getattr(self, "optimizer_name", None)
            assert cur is None or cur == ck["optimizer"], f"checkpoint was written by {ck['optimizer']}, this run uses {cur}"
            self.optimizer_name = ck["optimizer"]
        for k, v in (ck.get("other") or {}).items():
            self.other[k] = v.clone()
        self.steps, self.ema_updates = int(ck["steps"]), int(ck["ema_updates"])

    # ---- kernel layouts ---------------------------------------------------------------------------------------------
    def _wshape(self, key):
        s = self.shape[key]
        a, b = s[0], s[1]
        kh = s[2] if len(s) > 2 else 1
        kw = s[3] if len(s) > 3 else 1
        return a, b, kh, kw

    def conv(self, key, kind="fwd"):
        """Packed copy of a dense conv / linear / Conv1d / ConvTranspose2d weight.
        kind: fwd          conv2d weight (co,ci,kh,kw) -> [co][t][ci]                       (gradient-bearing)
              dgrad        the same weight for the input gradient of a stride-1 conv: [ci][flip t][co]
              dgrad_t      ... of a stride-2 conv (runs in yad_conv2d's TRANSPOSED mode): [ci][t][co]
              convT_fwd    ConvTranspose2d weight (ci,co,kh,kw) -> [co][t][ci]
              convT_dgrad  ... as the stride-2 NORMAL conv that computes its input gradient: [ci][t][co]; ALSO the layout in which
                           yad_conv_wgrad (operands swapped) produces the ConvTranspose2d weight gradient (gradient-bearing)"""
        ck = (key, kind)
        if ck in self._packed:
            return self._packed[ck]
        a, b, kh, kw = self._wshape(key)
        taps = kh * kw
        if kind == "fwd":
            n0, n2, s0, s2, flip, grad = a, b, b * taps, taps, 0, True
        elif kind in ("dgrad", "dgrad_t"):
            n0, n2, s0, s2, flip, grad = b, a, taps, b * taps, int(kind == "dgrad"), False
        elif kind == "convT_fwd":
            n0, n2, s0, s2, flip, grad = b, a, taps, b * taps, 0, False
        elif kind == "convT_dgrad":
            n0, n2, s0, s2, flip, grad = a, b, b * taps, taps, 0, True
        elif kind == "col_dgrad":
            # input gradient of a convolution applied to a tap-major column tensor (deformable conv, training.py): [t*ci + i][co], one table
            # entry per tap
            p0, p2 = _pad8(b), _pad8(a)
            assert p0 == b, "col_dgrad: input channels must be a multiple of 8"
            size = taps * p0 * p2
            off = self._top_t
            self._top_t += _pad8(size)
            assert self._top_t <= self.arena_t.numel(), "packed weight arena exhausted"
            for t in range(taps):
                self._entries.append((self.off[key] + t, off + t * p0 * p2, b, 1, a, p0, p2, taps, 0, b * taps, 0, 0, False))
            self._packed[ck] = PackedW(self.arena_t[off:off + size].view(taps * p0, 1, p2), None, 1, 1, taps * p0, p2)
            self._registered(taps)
            return self._packed[ck]
        else:
            raise ValueError(kind)
        p0, p2 = _pad8(n0), _pad8(n2)
        size = p0 * taps * p2
        off = self._top_t
        self._top_t += _pad8(size)
        assert self._top_t <= self.arena_t.numel(), "packed weight arena exhausted"
        self._entries.append((self.off[key], off, n0, taps, n2, p0, p2, s0, 1, s2, flip, 0, grad))
        w = self.arena_t[off:off + size].view(p0, taps, p2)
        gw = self.arena_g[off:off + size].view(p0, taps, p2) if grad else None
        self._packed[ck] = PackedW(w, gw, kh, kw, p0, p2)
        self._registered(1)
        return self._packed[ck]

    def f32(self, key, kind="pad"):
        """fp32 kernel-layout copy: 'pad' = vector zero-padded to a multiple of 8 (biases of 27- / 1-channel convs);
        'dw' = depthwise weight (c,1,k,k) -> [k*k][c].  Returns (tensor, gradient tensor)."""
        ck = (key, kind)
        if ck in self._packed:
            return self._packed[ck]
        s = self.shape[key]
        if kind == "pad":
            n = math.prod(s)
            n0, n1, n2, p0, p2, s0, s1, s2 = 1, 1, n, 1, _pad8(n), 0, 0, 1
            shape = (p2,)
        elif kind in ("dw", "dw_flip"):  # dw_flip: taps reversed = the depthwise conv that computes the input gradient
            c, taps = s[0], s[2] * s[3]
            n0, n1, n2, p0, p2, s0, s1, s2 = 1, taps, c, 1, c, 0, 1, taps
            shape = (taps, c)
        else:
            raise ValueError(kind)
        size = p0 * n1 * p2
        off = self._top_f
        self._top_f += _pad8(size)
        assert self._top_f <= self.arena_f.numel(), "fp32 packed arena exhausted"
        self._entries.append((self.off[key], off, n0, n1, n2, p0, p2, s0, s1, s2, int(kind == "dw_flip"), 1, kind != "dw_flip"))
        self._packed[ck] = (self.arena_f[off:off + size].view(shape), self.arena_gf[off:off + size].view(shape))
        self._registered(1)
        return self._packed[ck]

    def _registered(self, count):
        """a layout registered while a step is being traced is filled right away (its table entries alone), so the first forward already
        sees packed weights; from then on pack() refreshes every layout in one launch"""
        self._tables = None
        t, n, mx = self._table(self._entries[-count:])
        ops._call("yad_permute_pack", C.c_void_p(t.data_ptr()), n, mx, ops._fp(self.flat), ops._fp(self.arena_t), ops._fp(self.arena_f),
                  ops.dt(self.dtype), ops.stream_ptr())
        self._keep = getattr(self, "_keep", []) + [t]  # the table must outlive the asynchronous launch

    def _table(self, rows):
        arr = (YadPermuteEntry * max(len(rows), 1))()
        mx = 1
        for i, r in enumerate(rows):
            arr[i] = YadPermuteEntry(*r[:12])
            mx = max(mx, r[5] * r[3] * r[6])
        t = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)
        return t, len(rows), mx

    def _build_tables(self):
        def table(rows):
            arr = (YadPermuteEntry * max(len(rows), 1))()
            mx = 1
            for i, r in enumerate(rows):
                arr[i] = YadPermuteEntry(*r[:12])
                mx = max(mx, r[5] * r[3] * r[6])
            raw = bytes(arr)
            t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device)
            return t, len(rows), mx

        ent = self._entries
        self._tables = dict(pack=table(ent), unpack_t=table([e for e in ent if e[12] and not e[11]]),
                            unpack_f=table([e for e in ent if e[12] and e[11]]))

    def pack(self):
        """master fp32 parameters -> every registered kernel layout (one launch)"""
        if self._tables is None:
            self._build_tables()
        t, n, mx = self._tables["pack"]
        if n:
            ops._call("yad_permute_pack", C.c_void_p(t.data_ptr()), n, mx, ops._fp(self.flat), ops._fp(self.arena_t), ops._fp(self.arena_f),
                      ops.dt(self.dtype), ops.stream_ptr())

    def zero_grad(self):
        self.grad.zero_()
        self.arena_g[:max(self._top_t, 8)].zero_()
        self.arena_gf[:max(self._top_f, 8)].zero_()

    def unpack_grads(self):
        """packed weight gradients -> += flat gradient arena (reference layout)"""
        if self._tables is None:
            self._build_tables()
        for name, arena in (("unpack_t", self.arena_g), ("unpack_f", self.arena_gf)):
            t, n, mx = self._tables[name]
            if n:
                ops._call("yad_permute_unpack", C.c_void_p(t.data_ptr()), n, mx, ops._fp(arena), ops._fp(self.grad), ops.stream_ptr())

    def bump_batches_tracked(self):
        """nn.BatchNorm2d increments num_batches_tracked once per training forward (torch/nn/modules/batchnorm.py); kept so that state_dict()
        equals the reference's after the same number of steps"""
        for k in self.other:
            if k.endswith("num_batches_tracked"):
                self.other[k] += 1

    # ---- optimizer with device-resident hyper-parameters (CUDA-graph capturable) -----------------------------------------------------
    def set_hyper(self, lr=0.01, bias_lr=None, momentum=0.937, weight_decay=5e-4, max_norm=10.0, ema_decay=0.9999, ema_tau=2000.0, optimizer="SGD",
                  beta2=0.999, eps=1e-8):
        """Advance the step counters and upload this step's scalars (13 floats, pinned -> device, asynchronous): call right before replaying a
        graph that contains optimizer_step_dev()."""
        self.steps += 1
        self.ema_updates += 1
        t = float(self.steps)
        d = ema_decay * (1 - math.exp(-self.ema_updates / ema_tau)) if ema_decay > 0 else 1.0
        vals = [lr, lr, lr if bias_lr is None else bias_lr, weight_decay, 0.0, 0.0, momentum, beta2, eps, 1.0 - momentum ** t,
                math.sqrt(1.0 - beta2 ** t), max_norm, d]
        row = self._hyper_host[self.steps % self._hyper_host.shape[0]]
        row[:13] = torch.tensor(vals, dtype=torch.float32)
        self._hyper.copy_(row, non_blocking=True)
        self.optimizer_name = optimizer

    def ensure_adamw_state(self):
        if not hasattr(self, "mom2"):
            self.mom2 = torch.zeros_like(self.mom)  # second-moment arena (the first moment reuses the momentum arena)

    def optimizer_step_dev(self, optimizer="SGD"):
        """clip + optimizer + EMA with all scalars read from the device buffer set_hyper() fills: capturable, replayable"""
        self.norm_sq.zero_()
        st = ops.stream_ptr()
        hp = ops._fp(self._hyper)
        ops._call("yad_sqnorm", ops._fp(self.grad), self.total, ops._fp(self.norm_sq), st)
        if optimizer == "SGD":
            ops._call("yad_sgd_step_dev", ops._fp(self.flat), ops._fp(self.grad), ops._fp(self.mom), C.c_void_p(self.group.data_ptr()), self.total, hp,
                      ops._fp(self.norm_sq), st)
        elif optimizer == "AdamW":
            assert hasattr(self, "mom2"), "AdamW: call ensure_adamw_state() before capturing (state must not be allocated inside a capture)"
            ops._call("yad_adamw_step_dev", ops._fp(self.flat), ops._fp(self.grad), ops._fp(self.mom), ops._fp(self.mom2),
                      C.c_void_p(self.group.data_ptr()), self.total, hp, ops._fp(self.norm_sq), st)
        else:
            raise NotImplementedError(f"optimizer {optimizer!r}: SGD and AdamW are built")
        dptr = C.c_void_p(self._hyper.data_ptr() + 12 * 4)
        ops._call("yad_ema_update_dev", ops._fp(self.ema), ops._fp(self.flat), self.total, dptr, st)
        if self.btotal:
            ops._call("yad_ema_update_dev", ops._fp(self.ema_bufs), ops._fp(self.bufs), self.btotal, dptr, st)

    # ---- optimizer (engine/trainer.py:580-588) --------------------------------------------------------------------------
    def optimizer_step(self, lr=0.01, bias_lr=None, momentum=0.937, weight_decay=5e-4, max_norm=10.0, ema_decay=0.9999, ema_tau=2000.0,
                       optimizer="SGD", beta2=0.999, eps=1e-8):
        """clip_grad_norm_(max_norm) + the optimizer step over the flat arenas, then ModelEMA.update (parameters and BatchNorm buffers).
        optimizer: "SGD" (nesterov, momentum) or "AdamW" (betas = (momentum, beta2)): the two BaseTrainer.build_optimizer creates for this model
        (engine/trainer.py:773-808)."""
        self.optimizer_name = optimizer
        self.norm_sq.zero_()
        st = ops.stream_ptr()
        ops._call("yad_sqnorm", ops._fp(self.grad), self.total, ops._fp(self.norm_sq), st)
        lr3 = (C.c_float * 3)(lr, lr, lr if bias_lr is None else bias_lr)
        wd3 = (C.c_float * 3)(weight_decay, 0.0, 0.0)
        if optimizer == "SGD":
            ops._call("yad_sgd_step", ops._fp(self.flat), ops._fp(self.grad), ops._fp(self.mom), C.c_void_p(self.group.data_ptr()), self.total, lr3,
                      wd3, momentum, max_norm, ops._fp(self.norm_sq), int(self.steps == 0), st)
        elif optimizer == "AdamW":
            if not hasattr(self, "mom2"):
                self.mom2 = torch.zeros_like(self.mom)  # second-moment arena (the first moment reuses the momentum arena)
            ops._call("yad_adamw_step", ops._fp(self.flat), ops._fp(self.grad), ops._fp(self.mom), ops._fp(self.mom2), C.c_void_p(self.group.data_ptr()),
                      self.total, lr3, wd3, momentum, beta2, eps, self.steps + 1, max_norm, ops._fp(self.norm_sq), st)
        else:
            raise NotImplementedError(f"optimizer {optimizer!r}: SGD and AdamW are built")
        self.steps += 1
        if ema_decay > 0:
            self.ema_updates += 1
            d = ema_decay * (1 - math.exp(-self.ema_updates / ema_tau))
            ops._call("yad_ema_update", ops._fp(self.ema), ops._fp(self.flat), self.total, d, st)
            if self.btotal:
                ops._call("yad_ema_update", ops._fp(self.ema_bufs), ops._fp(self.bufs), self.btotal, d, st)
