"""CPU: pin the oracle against golden vectors produced by the reference's own modules."""
import os

import numpy as np
import pytest
import torch

from oracle import superdiff_oracle as O


def _seeded_input(seed, shape):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g)


def _draw_noise_stack(seed, shape, T):
    torch.manual_seed(seed)
    st = [torch.randn(shape)]
    for _ in range(T - 1):
        st.append(torch.randn_like(st[0]))
    return torch.stack(st, 0)


@pytest.fixture(scope="module")
def fwd_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "unet_forward.npz"))


@pytest.fixture(scope="module")
def sample_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "ddpm_sample.npz"))


def test_param_checksum(fwd_golden):
    for w in (0, 1):
        p = O.init_unet_params(w)
        cs = float(sum(v.double().abs().sum().item() for v in p.values()))
        assert cs == float(fwd_golden[f"param_checksum_w{w}"])
        assert len(p) == 54  # 54 state-dict tensors


def test_k1_unet_forward_matches_reference(fwd_golden):
    keys = [k for k in fwd_golden.files if k.startswith("fwd_")]
    assert len(keys) == 14
    for k in keys:
        _, w, xs, B, R, t = k.split("_")
        w, xs, B, R, t = int(w[1:]), int(xs[1:]), int(B[1:]), int(R[1:]), int(t[1:])
        p = O.init_unet_params(w)
        x = _seeded_input(xs, (B, 1, R, R))
        with torch.no_grad():
            y = O.unet_forward(p, x, torch.full((B,), t, dtype=torch.long))
        ref = torch.from_numpy(fwd_golden[k])
        # same ATen kernels, same thread count may differ -> allow reduction-order noise only
        assert torch.allclose(y, ref, rtol=0, atol=2e-5), (k, (y - ref).abs().max())


def test_schedule(sample_golden):
    s = O.Schedule(1000)
    assert np.array_equal(s.betas.numpy(), sample_golden["sched1000_betas"])
    assert np.array_equal(s.alpha_bars.numpy(), sample_golden["sched1000_alpha_bars"])


@pytest.mark.parametrize("name", ["small", "c1"])
def test_k2_replay_matches_reference_sample(sample_golden, name):
    wseed, nseed, T, *shape = [int(v) for v in sample_golden[f"sample_{name}_meta"]]
    p = O.init_unet_params(wseed)
    st = _draw_noise_stack(nseed, tuple(shape), T)
    y = O.ddpm_sample_replay(p, O.Schedule(T), st)
    ref = torch.from_numpy(sample_golden[f"sample_{name}"])
    assert torch.allclose(y, ref, rtol=0, atol=5e-4), (y - ref).abs().max()


def test_k3_self_superposition_is_ddpm_sample():
    """superposed_sample([m, m]) == DDPM.sample(m) bit-for-bit; kappa == 1/2; logq equal."""
    p = O.init_unet_params(0)
    T, shape = 12, (2, 1, 16, 16)
    st = _draw_noise_stack(7, shape, T)
    s = O.Schedule(T)
    y1 = O.ddpm_sample_replay(p, s, st)
    y2, kap, lq = O.superposed_sample([p, p], s, st)
    assert torch.equal(y1, y2)
    assert torch.equal(kap, torch.full_like(kap, 0.5))
    assert torch.equal(lq[..., 0], lq[..., 1])
    assert torch.equal(lq[0], torch.zeros_like(lq[0]))


def test_k4_kappa_properties():
    torch.manual_seed(0)
    B, D, M = 3, 64, 2
    x = torch.randn(B, 1, 8, 8)
    eps = [torch.randn(B, 1, 8, 8) for _ in range(M)]
    z = torch.randn(B, 1, 8, 8)
    s = O.Schedule(10)
    logq = torch.randn(B, M) * 5
    x1, lq1, k1 = O.superpose_step(x, eps, z, logq, s.alphas[5], s.alpha_bars[5], s.betas[5])
    assert torch.allclose(k1.sum(1), torch.ones(B), atol=1e-6)
    x2, lq2, k2 = O.superpose_step(x, eps, z, logq + 3.25, s.alphas[5], s.alpha_bars[5], s.betas[5])
    assert torch.allclose(k1, k2, atol=1e-6)
    assert torch.allclose(x1, x2, atol=1e-5)
    # two different models: increments differ
    assert not torch.allclose(lq1[:, 0] - logq[:, 0], lq1[:, 1] - logq[:, 1])


def test_and_mode_equalises_log_density_increments():
    """with the AND weights every model's Ito log-density increment of the step is the same
    (the defining property), kappa sums to 1, and identical models give the uniform weights."""
    g = torch.Generator().manual_seed(11)
    B, shape = 3, (1, 32, 32)
    x = torch.randn((B,) + shape, generator=g)
    z = torch.randn((B,) + shape, generator=g)
    s = O.Schedule(100)
    t = 40
    for M in (2, 3, 4):
        eps = [torch.randn((B,) + shape, generator=g) + 0.1 * m * x for m in range(M)]
        logq = torch.randn(B, M, generator=g)
        _, lq, kap = O.superpose_step(x, eps, z, logq, s.alphas[t], s.alpha_bars[t], s.betas[t], mode="and")
        inc = lq - logq
        assert torch.allclose(kap.sum(1), torch.ones(B), atol=1e-5)
        assert (inc - inc[:, :1]).abs().max().item() <= 1e-3
    _, lq, kap = O.superpose_step(x, [eps[0], eps[0]], z, torch.zeros(B, 2), s.alphas[t], s.alpha_bars[t], s.betas[t],
                                  mode="and")
    assert torch.equal(kap, torch.full((B, 2), 0.5))


def test_and_mode_sampler_keeps_densities_equal():
    params = [O.init_unet_params(0), O.init_unet_params(1)]
    T, shape = 6, (2, 1, 16, 16)
    stack = torch.randn((T,) + shape, generator=torch.Generator().manual_seed(3))
    x, kap, lq = O.superposed_sample(params, O.Schedule(T), stack, mode="and")
    assert torch.isfinite(x).all()
    assert ((lq[..., 0] - lq[..., 1]).abs().max() / lq.abs().max()).item() <= 1e-5


def _train_eval_case(g, name):
    meta = g[f"meta_{name}"]
    wseed, nseed, T = (int(v) for v in meta[:3])
    shape = tuple(int(v) for v in meta[3:7])
    t = torch.tensor([int(v) for v in meta[7:]], dtype=torch.long)
    x0 = torch.tanh(_seeded_input(nseed, shape))
    torch.manual_seed(nseed)
    noise = torch.randn_like(x0)
    return wseed, T, x0, t, noise


@pytest.mark.parametrize("name", ["a", "b"])
def test_n4_q_sample_and_p_losses_match_reference(golden_dir, name):
    """q_sample bit-exact, p_losses to fp32 reduction order, vs the reference's own ddpm.py:13-24."""
    g = np.load(os.path.join(golden_dir, "train_eval.npz"))
    wseed, T, x0, t, noise = _train_eval_case(g, name)
    s = O.Schedule(T)
    assert np.array_equal(O.q_sample(s, x0, t, noise).numpy(), g[f"q_{name}"])
    loss = O.p_losses(O.init_unet_params(wseed), s, x0, t, noise).item()
    assert abs(loss - float(g[f"loss_{name}"])) <= 1e-6 * abs(float(g[f"loss_{name}"]))


def test_attention_core_oracle_properties():
    g = torch.Generator().manual_seed(2)
    q = torch.randn(1, 2, 32, 8, generator=g); k = torch.randn(1, 2, 32, 8, generator=g)
    v = torch.randn(1, 2, 32, 8, generator=g)
    out = O.attention_core(q, k, v)
    want = torch.nn.functional.scaled_dot_product_attention(q, k, v)
    assert torch.allclose(out, want, atol=1e-5)
    assert torch.allclose(O.attention_core(q, k[:, :, :1].expand_as(k), v), v.mean(2, keepdim=True).expand_as(v), atol=1e-5)


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, exp in kat:
        r = O.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert tuple(int(v) for v in r[0]) == exp


def test_philox_normal_moments_and_sharding():
    a = O.philox_normal(1234, np.arange(4), 3, 4096)
    assert abs(a.mean()) < 0.02 and abs(a.std() - 1) < 0.02
    b = O.philox_normal(1234, np.array([2, 3]), 3, 4096)
    assert np.array_equal(a[2:], b)  # keyed by global sample id -> shard invariant
    c = O.philox_normal(1234, np.arange(4), 4, 4096)
    assert not np.array_equal(a, c)


def test_unet_attn_oracle_extension_properties():
    """EXTENSION oracle (oracle/unet_attn_oracle.py; no reference code): shape, class-conditioning, and the reduction to
    an attention-free network when the attention projections are zero (out = x + 0)."""
    from oracle import unet_attn_oracle as A
    p = A.init_params(0)
    assert len(p) == 129
    x = torch.randn(1, 1, 256, 256, generator=torch.Generator().manual_seed(0))
    t = torch.tensor([40])
    with torch.no_grad():
        e0 = A.unet_attn_forward(p, x, t, torch.tensor([0]))
        e1 = A.unet_attn_forward(p, x, t, torch.tensor([1]))
        assert e0.shape == x.shape and torch.isfinite(e0).all()
        assert (e0 - e1).abs().max() > 1e-3
        q = dict(p)
        for i in range(4):
            q[f"attn.{i}.proj.weight"] = torch.zeros(128, 128)
            q[f"attn.{i}.proj.bias"] = torch.zeros(128)
        a = A.unet_attn_forward(q, x, t, torch.tensor([0]))
        q2 = dict(q)
        q2["attn.0.qkv.weight"] = torch.randn(384, 128)  # irrelevant once proj is zero
        assert torch.equal(a, A.unet_attn_forward(q2, x, t, torch.tensor([0])))
