"""CPU ORACLE (test infrastructure, NOT product code) -- the reference's optimizer step restated over a state dict:
torch.nn.utils.clip_grad_norm_(max_norm=10) + torch.optim.SGD(nesterov=True) with the three parameter groups of
BaseTrainer.build_optimizer (engine/trainer.py:580-588, 784-808) and ModelEMA.update (utils/torch_utils.py:511-541).

Parity status: PINNED -- tests/golden/opt_step.npz holds two steps of the live reference (oracle/gen_golden.py opt_step); group membership is
pinned by tests/golden/optimizer_groups.json.  Only tests/ may import this file."""
import math

import torch


def clip_coef(grads, max_norm=10.0):
    """torch.nn.utils.clip_grad_norm_: total L2 norm over every parameter that has a gradient; coefficient clamped to 1"""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values() if g is not None)).float()
    return float(torch.clamp(max_norm / (total + 1e-6), max=1.0)), float(total)


def sgd_step(params, grads, momentum_bufs, groups, lr=0.01, momentum=0.937, weight_decay=5e-4, max_norm=10.0):
    """params / grads: {key: tensor}; groups: {key: 0 decay | 1 norm weight | 2 bias}; momentum_bufs: {} on the first step (torch.optim.SGD
    initialises the buffer with the first gradient).  Updates params and momentum_bufs in place; returns the gradient norm."""
    coef, total = clip_coef(grads, max_norm)
    for k, p in params.items():
        g = grads.get(k)
        if g is None:
            continue  # no gradient: SGD skips the parameter (DFL's frozen projection, AdaptiveDynamicTanh.scale_weights)
        d = g * coef + (weight_decay if groups[k] == 0 else 0.0) * p
        if k not in momentum_bufs:
            momentum_bufs[k] = d.clone()
        else:
            momentum_bufs[k].mul_(momentum).add_(d)
        d = d + momentum * momentum_bufs[k]  # nesterov
        p.sub_(lr * d)
    return total


def ema_update(ema, model_sd, updates, decay=0.9999, tau=2000.0):
    """ModelEMA.update over every floating-point state-dict entry (parameters AND BatchNorm buffers)"""
    d = decay * (1 - math.exp(-updates / tau))
    for k, v in ema.items():
        if v.dtype.is_floating_point:
            v.mul_(d).add_((1 - d) * model_sd[k])
    return d


def adamw_step(params, grads, state, groups, step, lr=1e-3, betas=(0.937, 0.999), eps=1e-8, weight_decay=5e-4, max_norm=10.0):
    """torch.optim.AdamW as BaseTrainer.build_optimizer configures it (engine/trainer.py:805: betas = (momentum, 0.999); decoupled weight decay on
    group 0 only).  state: {key: (exp_avg, exp_avg_sq)}; step counts from 1.  Pinned by tests/golden/opt_step_adamw.npz."""
    coef, total = clip_coef(grads, max_norm)
    b1, b2 = betas
    for k, p in params.items():
        g = grads.get(k)
        if g is None:
            continue
        g = g * coef
        if k not in state:
            state[k] = (torch.zeros_like(p), torch.zeros_like(p))
        m, v = state[k]
        p.mul_(1 - lr * (weight_decay if groups[k] == 0 else 0.0))
        m.mul_(b1).add_((1 - b1) * g)
        v.mul_(b2).add_((1 - b2) * g * g)
        denom = v.sqrt() / math.sqrt(1 - b2 ** step) + eps
        p.sub_((lr / (1 - b1 ** step)) * m / denom)
    return total
