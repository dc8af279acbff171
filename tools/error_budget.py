"""Error budget of the UNet forward: emulate, in fp32 torch on the CPU, each rounding point of the
CUDA pipeline separately and measure the rel-L2 of eps-hat against the unrounded fp32 forward (== the reference).

Rounding points of the kernels (csrc/conv_tc4.cuh, unet_kernels.cuh):
  W  conv weights of the 64/128-channel layers rounded to the MMA operand type
  S  raw conv outputs (inter-layer activations) stored in HBM in a 16-bit type (GroupNorm statistics come from the
     fp32 accumulators, i.e. from the UNROUNDED values)
  A  GroupNorm+SiLU output rounded to the MMA operand type
  T  SiLU through tanh.approx.f32 (max relative error 2^-11 per the PTX ISA): modelled as a uniform random relative
     perturbation of tanh of that size
  I  the 1->64 input conv on TF32 operands (10-bit mantissa)

    python tools/error_budget.py            # prints a table; CPU only
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import superdiff_oracle as O  # noqa: E402


def rnd(x, kind):
    if kind == "f32":
        return x
    if kind == "bf16":
        return x.to(torch.bfloat16).float()
    if kind == "f16":
        return x.to(torch.float16).float()
    if kind == "tf32":  # round-to-nearest-away on 13 dropped bits (cvt.rna.tf32.f32)
        i = x.view(torch.int32)
        return ((i + 0x1000) & ~0x1FFF).view(torch.float32)
    raise ValueError(kind)


def silu(v, tanh_err, gen):
    if not tanh_err:
        return F.silu(v)
    h = 0.5 * v
    t = torch.tanh(h)
    t = t * (1 + (torch.rand(t.shape, generator=gen) * 2 - 1) * 2.0 ** -11)
    return h + h * t


def forward(p, x, t, W="f32", S="f32", A="f32", T=False, I="f32", seed=0):
    gen = torch.Generator().manual_seed(seed)
    temb = O.time_mlp(p, t)
    h = x
    for name in O.BLOCKS:
        cin = p[f"{name}.block.0.weight"].shape[0]
        cout = p[f"{name}.block.3.weight"].shape[0]
        for gi, ci, (gn, cv) in ((0, cin, (0, 2)), (1, cout, (3, 5))):
            co = cout
            tc = ci >= 64 and co >= 64          # tcgen05 layers
            o1 = ci == 64 and co == 1           # conv_out1 (mma.sync, 16-bit operands)
            cin1 = ci == 1 and co == 64         # conv_in (TF32)
            # statistics from the unrounded producer output, normalisation applied to the stored (rounded) values
            G = min(4, ci)
            B = h.shape[0]
            hv = h.reshape(B, G, -1)
            mean, var = hv.mean(2, keepdim=True), hv.var(2, unbiased=False, keepdim=True)
            stored = rnd(h, S) if ci >= 64 else h
            hn = ((stored.reshape(B, G, -1) - mean) / torch.sqrt(var + 1e-5)).reshape(h.shape)
            hn = hn * p[f"{name}.block.{gn}.weight"][None, :, None, None] + p[f"{name}.block.{gn}.bias"][None, :, None, None]
            a = silu(hn, T and (tc or o1), gen)
            w = p[f"{name}.block.{cv}.weight"]
            if tc or o1:
                a, w = rnd(a, A), rnd(w, W)
            elif cin1:
                a, w = rnd(a, I), rnd(w, I)
            h = F.conv2d(a, w, p[f"{name}.block.{cv}.bias"], padding=1)
        te = F.linear(temb, p[f"{name}.time_emb.weight"], p[f"{name}.time_emb.bias"])
        h = h + te[:, :, None, None]
    return h


def main():
    torch.set_num_threads(os.cpu_count())
    rows = [
        ("all fp32 (sanity)", dict()),
        ("W bf16 only", dict(W="bf16")),
        ("S bf16 only", dict(S="bf16")),
        ("A bf16 only", dict(A="bf16")),
        ("T tanh.approx only", dict(T=True)),
        ("I tf32 only", dict(I="tf32")),
        ("round-1 pipeline: W,S,A bf16 + T + I", dict(W="bf16", S="bf16", A="bf16", T=True, I="tf32")),
        ("W f16 only", dict(W="f16")),
        ("S f16 only", dict(S="f16")),
        ("A f16 only", dict(A="f16")),
        ("W,A f16, S bf16 + T + I", dict(W="f16", S="bf16", A="f16", T=True, I="tf32")),
        ("W,S,A f16 + T + I", dict(W="f16", S="f16", A="f16", T=True, I="tf32")),
        ("W,S,A f16 + I (exact SiLU)", dict(W="f16", S="f16", A="f16", I="tf32")),
    ]
    cases = [(0, 64, 999), (1, 64, 1), (0, 128, 250)]
    print(f"{'variant':44s}" + "".join(f"  w{w} R{R} t{t:<4d}" for w, R, t in cases))
    with torch.no_grad():
        refs = {}
        for w, R, t in cases:
            p = O.init_unet_params(w)
            x = torch.randn((2, 1, R, R), generator=torch.Generator().manual_seed(100 + R))
            tt = torch.full((2,), t, dtype=torch.long)
            refs[(w, R, t)] = (p, x, tt, O.unet_forward(p, x, tt))
        for label, kw in rows:
            out = []
            for c in cases:
                p, x, tt, ref = refs[c]
                y = forward(p, x, tt, **kw)
                out.append(((y - ref).norm() / ref.norm()).item())
            print(f"{label:44s}" + "".join(f"  {v:12.3e}" for v in out))


if __name__ == "__main__":
    main()
