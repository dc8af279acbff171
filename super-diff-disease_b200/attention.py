"""Self-attention core on the B200.

A north-star-only extension: the reference UNet (src/models/unet.py:37-65) has NO attention, so there is no reference
parity to claim here -- the checker is this repo's own ``oracle.attention_core``.  ``attention_core`` wraps
``sdd_attention_fwd``: the fused flash-style tcgen05 / TMEM / TMA kernel in csrc/attention.cuh, sized for the 32^2 and
16^2 feature maps the north star names (1024 / 256 tokens, head_dim 64).  No fallback: CUDA tensors only.
"""
import math

import torch

from super_diff_disease_b200 import _lib


@torch.no_grad()
def attention_core(q, k, v, scale=None, *, v_is_transposed=False):
    """softmax(scale * q k^T) v per (batch, head).

    q, k: fp16 CUDA [B, heads, S, 64];  v: fp16 [B, heads, S, 64], or [B, heads, 64, S] with v_is_transposed=True (the
    layout the kernel consumes; a producer that writes V^T directly saves the transpose).  S % 128 == 0.
    Returns fp16 [B, heads, S, 64].
    """
    for name, t in (("q", q), ("k", k), ("v", v)):
        _lib.require_cuda(t, name)
        if t.dtype != torch.float16 or t.dim() != 4:
            raise _lib.SddError(f"{name} must be a 4-d float16 tensor")
    B, Hh, S, D = q.shape
    if D != 64 or S % 128 != 0:
        raise _lib.SddError(f"head_dim must be 64 and S a multiple of 128, got S={S}, head_dim={D}")
    if tuple(k.shape) != (B, Hh, S, D):
        raise _lib.SddError("k must have q's shape")
    vt = v if v_is_transposed else v.transpose(2, 3)
    if tuple(vt.shape) != (B, Hh, D, S):
        raise _lib.SddError("v must be [B, heads, S, 64] (or [B, heads, 64, S] with v_is_transposed=True)")
    qc, kc, vtc = q.contiguous(), k.contiguous(), vt.contiguous()
    out = torch.empty_like(qc)
    sc = float(scale) if scale is not None else 1.0 / math.sqrt(D)
    with torch.cuda.device(q.device):
        _lib.check(_lib.lib().sdd_attention_fwd(qc.data_ptr(), kc.data_ptr(), vtc.data_ptr(), out.data_ptr(), B * Hh, S,
                                                D, sc, _lib.stream_ptr(q.device)))
    return out


@torch.no_grad()
def attention_block(x, gn_weight, gn_bias, w_qkv, b_qkv, w_out, b_out, heads=2):
    """out = x + proj(attention(q, k, v)) with [q|k|v] = GroupNorm(4, 128)(x) W_qkv^T + b_qkv (2 heads of 64).

    x: fp16 CUDA NHWC [B, H, W, 128] (H*W % 128 == 0: 32^2, 16^2 ...); gn_weight / gn_bias fp32 [128];
    w_qkv fp32 [384, 128], b_qkv [384]; w_out fp32 [128, 128], b_out [128].  Returns fp16 [B, H, W, 128].
    Wraps sdd_attention_block_nhwc (GroupNorm statistics -> mma.sync projection with the GroupNorm affine fused on load
    and a head-split / V-transposed epilogue -> fused tcgen05 attention -> projection + bias + residual)."""
    _lib.require_cuda(x, "x")
    if x.dtype != torch.float16 or x.dim() != 4 or x.shape[-1] != 128:
        raise _lib.SddError("x must be float16 NHWC [B, H, W, 128]")
    B, H, W, C = x.shape
    dev = x.device
    f32 = lambda t: torch.as_tensor(t, dtype=torch.float32, device=dev).contiguous()  # noqa: E731
    gw, gb, wq, bq, wo, bo = (f32(t) for t in (gn_weight, gn_bias, w_qkv, b_qkv, w_out, b_out))
    if tuple(wq.shape) != (3 * C, C) or tuple(wo.shape) != (C, C) or bq.numel() != 3 * C or bo.numel() != C:
        raise _lib.SddError("w_qkv must be [384, 128], b_qkv [384], w_out [128, 128], b_out [128]")
    xc = x.contiguous()
    out = torch.empty_like(xc)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().sdd_attention_block_nhwc(xc.data_ptr(), gw.data_ptr(), gb.data_ptr(), wq.data_ptr(),
                                                       bq.data_ptr(), wo.data_ptr(), bo.data_ptr(), out.data_ptr(), B,
                                                       H * W, C, int(heads), _lib.stream_ptr(dev)))
    return out
