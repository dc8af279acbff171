"""Generate golden vectors by running the UNMODIFIED reference modules on CPU.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/unet_forward.npz, unet_forward_r256.npz, ddpm_sample.npz and train_eval.npz.

Weights come from oracle.init_unet_params(seed) (deterministic CPU generator) and are
loaded into the reference ``UNet`` with ``load_state_dict(strict=True)``, so the fixtures
carry outputs only; ``param_checksum`` guards against generator drift.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from models.unet import UNet  # noqa: E402  (reference, unmodified)
from models.ddpm import DDPM  # noqa: E402

from oracle.superdiff_oracle import Schedule, ddpm_sample_replay, init_unet_params  # noqa: E402


def param_checksum(p):
    return float(sum(v.double().abs().sum().item() for v in p.values()))


def ref_model(seed):
    p = init_unet_params(seed)
    m = UNet()
    m.load_state_dict(p, strict=True)
    m.eval()
    return m, p


def seeded_input(seed, shape):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g)


def draw_noise_stack(seed, shape, T):
    """Same draws, same order as ddpm.py:33,36 under torch.manual_seed(seed) on CPU."""
    torch.manual_seed(seed)
    st = [torch.randn(shape)]
    for _ in range(T - 1):
        st.append(torch.randn_like(st[0]))
    return torch.stack(st, 0)


def main():
    torch.set_num_threads(os.cpu_count())
    out = {}
    # K1: UNet.forward
    cases = [(101, 2, 16, [0, 1, 49]), (102, 2, 64, [0, 1, 999]), (103, 1, 128, [250])]
    for wseed in (0, 1):
        m, p = ref_model(wseed)
        out[f"param_checksum_w{wseed}"] = np.float64(param_checksum(p))
        for xseed, B, R, ts in cases:
            x = seeded_input(xseed, (B, 1, R, R))
            for t in ts:
                with torch.no_grad():
                    y = m(x, torch.full((B,), t, dtype=torch.long))
                out[f"fwd_w{wseed}_x{xseed}_B{B}_R{R}_t{t}"] = y.numpy()
    np.savez_compressed(os.path.join(HERE, "unet_forward.npz"), **out)
    # K1 at the headline resolution (BASELINE configs[2]): R = 256, both weight seeds
    out = {}
    for wseed in (0, 1):
        m, _ = ref_model(wseed)
        x = seeded_input(104, (1, 1, 256, 256))
        for t in (0, 125, 249):
            with torch.no_grad():
                y = m(x, torch.full((1,), t, dtype=torch.long))
            out[f"fwd_w{wseed}_x104_B1_R256_t{t}"] = y.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "unet_forward_r256.npz"), **out)

    # K2: DDPM.sample (reference RNG) -- replay must reproduce it from the drawn stack
    out = {}
    for name, wseed, nseed, shape, T in [("small", 0, 7, (2, 1, 16, 16), 12),
                                         ("c1", 1, 1234, (4, 1, 64, 64), 50)]:
        m, _ = ref_model(wseed)
        torch.manual_seed(nseed)
        with torch.no_grad():
            y = DDPM(num_timesteps=T).sample(m, shape, "cpu")
        out[f"sample_{name}"] = y.numpy()
        out[f"sample_{name}_meta"] = np.array([wseed, nseed, T, *shape], dtype=np.int64)
        # K2: replaying the drawn stack through the oracle's loop gives the reference's bits
        st = draw_noise_stack(nseed, shape, T)
        _, p = ref_model(wseed)
        assert torch.equal(ddpm_sample_replay(p, Schedule(T), st), y), name
    d = DDPM(num_timesteps=1000)
    out["sched1000_betas"] = d.betas.numpy()
    out["sched1000_alpha_bars"] = d.alpha_bars.numpy()
    np.savez_compressed(os.path.join(HERE, "ddpm_sample.npz"), **out)
    # N4: q_sample and p_losses (ddpm.py:13-24) -- the loss draws its noise first thing, so seeding and drawing
    # randn_like(x0) ourselves reproduces the reference's draw
    out = {}
    for name, wseed, nseed, shape, T, ts in [("a", 0, 21, (3, 1, 32, 32), 100, [0, 57, 99]),
                                             ("b", 1, 22, (2, 1, 64, 64), 1000, [3, 998])]:
        m, _ = ref_model(wseed)
        d = DDPM(num_timesteps=T)
        x0 = torch.tanh(seeded_input(nseed, shape))
        t = torch.tensor(ts, dtype=torch.long)
        torch.manual_seed(nseed)
        noise = torch.randn_like(x0)
        out[f"q_{name}"] = d.q_sample(x0, t, noise).numpy()
        torch.manual_seed(nseed)
        with torch.no_grad():
            out[f"loss_{name}"] = np.float32(d.p_losses(m, x0, t).item())
        out[f"meta_{name}"] = np.array([wseed, nseed, T, *shape, *ts], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "train_eval.npz"), **out)
    print("golden written")


if __name__ == "__main__":
    main()
