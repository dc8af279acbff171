"""fp32 oracle of the ``UNetAttn`` EXTENSION.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED BY THE REFERENCE: the reference has no such network -- its UNet (/root/reference/src/models/unet.py:37-65)
is five full-resolution blocks with no attention, down/up-sampling, skip connections or class input.  This module is
OUR definition of the variant BASELINE configs[2] names ("256x256 with attention at 32^2/16^2", "class-conditional
UNets"); every number measured against it is an extension result, never reference parity.  What IS the reference's: the
residual block (unet.py:18-34) and the time MLP (unet.py:11-16, :40-45), reused unchanged from superdiff_oracle.

    enc0 RB(1,64)@R  -pool-> enc1 RB(64,128)@R/2 -pool-> enc2 RB(128,128)@R/4 -pool-> enc3 RB(128,128)+Attn@R/8 -pool->
    enc4 RB(128,128)+Attn@R/16 -> mid RB(128,128)+Attn@R/16 -up,+enc3-> dec0 RB(128,128)+Attn@R/8 -up,+enc2->
    dec1 RB(128,128)@R/4 -up,+enc1-> dec2 RB(128,64)@R/2 -up,+enc0-> out RB(64,1)@R

pool = F.avg_pool2d(., 2); up = nearest-neighbour x2; skips are ADDED; Attn = superdiff_oracle.attention_block (pre-norm
GroupNorm(4,128), 2 heads of 64); every block's time embedding is time_mlp(t) + class_emb[y].  State-dict keys follow the
product module (super_diff_disease_b200.UNetAttn): time_mlp.{1,3}, class_emb, enc.{0..4}, mid, dec.{0..3}, attn.{0..3}.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from oracle import superdiff_oracle as O

BLOCKS = ("enc.0", "enc.1", "enc.2", "enc.3", "enc.4", "mid", "dec.0", "dec.1", "dec.2", "dec.3")
CHANS = ((1, 64), (64, 128), (128, 128), (128, 128), (128, 128), (128, 128), (128, 128), (128, 128), (128, 64), (64, 1))


def init_params(seed: int, num_classes: int = 2) -> O.Params:
    """Random parameters with the product module's keys and shapes (deterministic CPU generator)."""
    g = torch.Generator().manual_seed(1000 + seed)

    def u(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * b

    p: O.Params = {}
    p["time_mlp.1.weight"] = u((1024, 256), 256)
    p["time_mlp.1.bias"] = u((1024,), 256)
    p["time_mlp.3.weight"] = u((256, 1024), 1024)
    p["time_mlp.3.bias"] = u((256,), 1024)
    p["class_emb.weight"] = torch.randn((num_classes, 256), generator=g) * 0.5
    for name, (ci, co) in zip(BLOCKS, CHANS):
        p[f"{name}.block.0.weight"] = 1.0 + 0.1 * (torch.rand((ci,), generator=g) * 2 - 1)
        p[f"{name}.block.0.bias"] = 0.1 * (torch.rand((ci,), generator=g) * 2 - 1)
        p[f"{name}.block.2.weight"] = u((co, ci, 3, 3), ci * 9)
        p[f"{name}.block.2.bias"] = u((co,), ci * 9)
        p[f"{name}.block.3.weight"] = 1.0 + 0.1 * (torch.rand((co,), generator=g) * 2 - 1)
        p[f"{name}.block.3.bias"] = 0.1 * (torch.rand((co,), generator=g) * 2 - 1)
        p[f"{name}.block.5.weight"] = u((co, co, 3, 3), co * 9)
        p[f"{name}.block.5.bias"] = u((co,), co * 9)
        p[f"{name}.time_emb.weight"] = u((co, 256), 256)
        p[f"{name}.time_emb.bias"] = u((co,), 256)
    for i in range(4):
        p[f"attn.{i}.norm.weight"] = 1.0 + 0.1 * (torch.rand((128,), generator=g) * 2 - 1)
        p[f"attn.{i}.norm.bias"] = 0.1 * (torch.rand((128,), generator=g) * 2 - 1)
        p[f"attn.{i}.qkv.weight"] = u((384, 128), 128)
        p[f"attn.{i}.qkv.bias"] = u((384,), 128)
        p[f"attn.{i}.proj.weight"] = u((128, 128), 128)
        p[f"attn.{i}.proj.bias"] = u((128,), 128)
    return p


def _attn(p: O.Params, i: int, h: torch.Tensor) -> torch.Tensor:
    """NCHW in / out around superdiff_oracle.attention_block (NHWC)."""
    y = O.attention_block(h.permute(0, 2, 3, 1), p[f"attn.{i}.norm.weight"], p[f"attn.{i}.norm.bias"],
                          p[f"attn.{i}.qkv.weight"], p[f"attn.{i}.qkv.bias"], p[f"attn.{i}.proj.weight"],
                          p[f"attn.{i}.proj.bias"], heads=2)
    return y.permute(0, 3, 1, 2)


def unet_attn_forward(p: O.Params, x: torch.Tensor, t: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """x [B,1,R,R], t int64 [B], y int64 [B] class labels -> eps-hat [B,1,R,R]."""
    emb = O.time_mlp(p, t) + p["class_emb.weight"][y]
    rb = lambda name, h: O.residual_block(p, name, h, emb)  # noqa: E731  (unet.py:18-34, unchanged)
    up = lambda h: F.interpolate(h, scale_factor=2, mode="nearest")  # noqa: E731
    e0 = rb("enc.0", x)
    e1 = rb("enc.1", F.avg_pool2d(e0, 2))
    e2 = rb("enc.2", F.avg_pool2d(e1, 2))
    e3 = _attn(p, 0, rb("enc.3", F.avg_pool2d(e2, 2)))
    e4 = _attn(p, 1, rb("enc.4", F.avg_pool2d(e3, 2)))
    m = _attn(p, 2, rb("mid", e4))
    d0 = _attn(p, 3, rb("dec.0", up(m) + e3))
    d1 = rb("dec.1", up(d0) + e2)
    d2 = rb("dec.2", up(d1) + e1)
    return rb("dec.3", up(d2) + e0)


def superposed_sample(params_list, labels, sched: O.Schedule, noise_stack: torch.Tensor, **kw):
    """superdiff_oracle.superposed_sample with the variant's forward (model i conditions on class labels[i])."""
    dev = noise_stack.device
    B = noise_stack.shape[1]
    T, M = sched.T, len(params_list)
    x = noise_stack[0].clone()
    logq = torch.zeros(B, M, device=dev)
    kappas = torch.zeros(T, B, M, device=dev)
    logqs = torch.zeros(T + 1, B, M, device=dev)
    k = 1
    with torch.no_grad():
        for it, t in enumerate(reversed(range(T))):
            tt = torch.full((B,), t, dtype=torch.long, device=dev)
            noise = noise_stack[k] if t > 0 else torch.zeros_like(x)
            k += 1 if t > 0 else 0
            eps = [unet_attn_forward(p, x, tt, torch.full((B,), int(c), dtype=torch.long, device=dev))
                   for p, c in zip(params_list, labels)]
            x, logq, kappa = O.superpose_step(x, eps, noise, logq, sched.alphas[t], sched.alpha_bars[t], sched.betas[t], **kw)
            kappas[it] = kappa
            logqs[it + 1] = logq
    return x, kappas, logqs
