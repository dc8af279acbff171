"""GPU (-m gpu): the UNetAttn EXTENSION against ITS OWN oracle.

No reference parity is claimed anywhere in this file: the reference has no attention / multi-resolution /
class-conditional UNet (/root/reference/src/models/unet.py:37-65), so the checker is oracle/unet_attn_oracle.py (our
definition; its residual blocks and time MLP are the reference's, reused from superdiff_oracle).  Tolerances: forward
rel-L2 <= 5e-3 (fp16 operands / storage, fp16 q, k, v, P inside the attention blocks); short superposed runs: x rel-L2
<= 2e-3, kappa <= 2e-2, log q rel-to-max <= 5e-3.
"""
import copy
import json
import os

import pytest
import torch

from oracle import superdiff_oracle as O
from oracle import unet_attn_oracle as A

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def S():
    import __graft_entry__ as G
    G.build()
    import super_diff_disease_b200 as S
    assert torch.cuda.is_available()
    return S


def _report(**kw):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.jsonl"), "a") as f:
        f.write(json.dumps(kw) + "\n")
    print("PARITY", kw)


def _model(S, seed, dev, label=0):
    p = A.init_params(seed)
    m = S.UNetAttn()
    m.load_state_dict(p, strict=True)
    return p, m.to(dev).eval().set_label(label)


def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


@pytest.mark.parametrize("seed,B,R,ts,ys", [(0, 2, 256, [0, 249], [0, 1]), (1, 1, 256, [125], [1]), (0, 1, 512, [999], [0])])
def test_unet_attn_forward_matches_its_oracle(S, seed, B, R, ts, ys):
    dev = torch.device("cuda:0")
    p, m = _model(S, seed, dev)
    x = torch.randn((B, 1, R, R), generator=torch.Generator().manual_seed(7 + R))
    t, y = torch.tensor(ts), torch.tensor(ys)
    with torch.no_grad():
        ref = A.unet_attn_forward(p, x, t, y)
    out = m(x.to(dev), t.to(dev), y.to(dev)).cpu()
    rel, mx = _rel(out, ref), ((out - ref).abs().max() / ref.abs().max()).item()
    _report(test="unet_attn_forward", R=R, B=B, rel_l2=rel, relmax=mx)
    assert rel <= 5e-3 and mx <= 2e-2
    # the class embedding is live: another label gives another eps-hat; y=None uses the handle's label
    other = m(x.to(dev), t.to(dev), (1 - y).to(dev)).cpu()
    assert _rel(other, ref) > 1e-2
    m.set_label(int(ys[0]))
    same = m(x[:1].to(dev), t[:1].to(dev)).cpu()
    assert torch.equal(same, out[:1])


def test_unet_attn_chunking_and_copies(S):
    dev = torch.device("cuda:0")
    _, m = _model(S, 0, dev, label=1)
    x = torch.randn((3, 1, 256, 256), generator=torch.Generator().manual_seed(3)).to(dev)
    t = torch.tensor([5, 100, 200], device=dev)
    y = m(x, t)
    m2 = copy.deepcopy(m)
    assert m2._handle is None and m2.label == 1
    m2.set_max_chunk(2)  # passes of 2 + 1 samples
    assert torch.equal(m2(x, t), y)
    assert len(m.state_dict()) == 129
    with pytest.raises(S.SddError):
        m(torch.zeros(1, 1, 128, 128, device=dev), t[:1])  # five levels need H % 256 == 0


@pytest.mark.parametrize("mode", ["or", "and"])
def test_superposed_sampling_with_class_conditional_attention_unets(S, mode):
    """Two class-conditional attention UNets (label 0 = TB, label 1 = Pneumonia) through the same captured-graph sampler
    and fused update kernel as the reference path, vs the extension's oracle fed the same noise stack."""
    dev = torch.device("cuda:0")
    T, shape = 6, (2, 1, 256, 256)
    p0, m0 = _model(S, 0, dev, label=0)
    p1, m1 = _model(S, 1, dev, label=1)
    stack = torch.randn((T,) + shape, generator=torch.Generator().manual_seed(11))
    xr, kr, lr = A.superposed_sample([p0, p1], [0, 1], O.Schedule(T), stack, mode=mode)
    x, kap, lq = S.superposed_sample([m0, m1], S.DDPM(T), shape, dev, noise=stack.to(dev), return_trajectory=True, mode=mode)
    xe, kape, lqe = S.superposed_sample([m0, m1], S.DDPM(T), shape, dev, noise=stack.to(dev), return_trajectory=True,
                                        mode=mode, use_graph=False)
    assert torch.equal(x, xe) and torch.equal(kap, kape) and torch.equal(lq, lqe)  # graph replay == eager
    rel = _rel(x.cpu(), xr)
    ek = (kap.cpu() - kr).abs().max().item()
    el = ((lq.cpu() - lr).abs().max() / lr.abs().max()).item()
    _report(test="unet_attn_superposed", mode=mode, T=T, shape=list(shape), x_rel_l2=rel, kappa_abs=ek, logq_rel=el)
    assert rel <= 2e-3 and el <= 5e-3 and (ek <= 2e-2 or mode == "and")
    # the label is part of the model: swapping it changes the samples, Philox sharding still changes no bit
    a = S.superposed_sample([m0, m1], S.DDPM(T), shape, dev, seed=5)
    lo = S.superposed_sample([m0, m1], S.DDPM(T), (1, 1, 256, 256), dev, seed=5, sample_offset=0)
    hi = S.superposed_sample([m0, m1], S.DDPM(T), (1, 1, 256, 256), dev, seed=5, sample_offset=1)
    assert torch.equal(torch.cat([lo, hi]), a)
    m1.set_label(0)
    b = S.superposed_sample([m0, m1], S.DDPM(T), shape, dev, seed=5)
    assert not torch.equal(a, b)
