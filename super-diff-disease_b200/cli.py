"""Command-line superposed sampler (checkpoint I/O + entry point).

Loads the per-disease checkpoints the reference's trainer writes --
``<checkpoints>/<experiment>/<run>/{TB,PNEUMONIA}/ema_epoch{N}.pt`` (plain ``state_dict`` files:
/root/reference/src/train/training_logic.py:47-48; directory layout: /root/reference/src/utils/env.py:18-37) --
runs ``superposed_sample`` on the B200 and writes the samples plus the kappa / log-density traces:

    python -m super_diff_disease_b200.cli --tb .../TB/ema_epoch100.pt --pneumonia .../PNEUMONIA/ema_epoch100.pt \
        --batch 64 --resolution 256 --steps 250 --seed 1234 --out samples.npz

``--checkpoint-root/--experiment/--run/--epoch`` resolve the two paths with the reference's layout instead.
No CPU fallback: without a B200 and the built library this exits with the library's error.
"""
import argparse
import os
import sys

import numpy as np
import torch


def checkpoint_path(root, experiment, run, task, epoch, ema=True):
    """<root>/<experiment>/<run>/<task>/{ema,ddpm}_epoch{epoch}.pt  (env.py:26, training_logic.py:47-48)."""
    return os.path.join(root, experiment, run, task, f"{'ema' if ema else 'ddpm'}_epoch{epoch}.pt")


def load_unet(path, device):
    """A reference checkpoint is ``torch.save(model.state_dict(), path)`` of the default ``UNet()``; an ``ema_pytorch``
    wrapper's own state_dict (keys prefixed ``ema_model.``) is accepted too."""
    from super_diff_disease_b200 import UNet
    sd = torch.load(path, map_location="cpu", weights_only=True)
    if any(k.startswith("ema_model.") for k in sd):
        sd = {k[len("ema_model."):]: v for k, v in sd.items() if k.startswith("ema_model.")}
    m = UNet()
    m.load_state_dict(sd, strict=True)
    return m.to(device)


def save_grid_pgm(x, path, cols=8):
    """Samples [B,1,H,W] -> one 8-bit PGM contact sheet (min/max normalised per image, like the reference's
    matplotlib ``imshow`` in utils/visualization.py:6-28; no plotting dependency needed)."""
    x = x.detach().float().cpu().numpy()[:, 0]
    B, H, W = x.shape
    cols = min(cols, B)
    rows = (B + cols - 1) // cols
    sheet = np.zeros((rows * H, cols * W), dtype=np.uint8)
    for i in range(B):
        lo, hi = float(x[i].min()), float(x[i].max())
        img = (x[i] - lo) / (hi - lo + 1e-12)
        r, c = divmod(i, cols)
        sheet[r * H:(r + 1) * H, c * W:(c + 1) * W] = np.round(img * 255).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(f"P5\n{sheet.shape[1]} {sheet.shape[0]}\n255\n".encode())
        f.write(sheet.tobytes())


def main(argv=None):
    ap = argparse.ArgumentParser(description="SuperDiff TB+Pneumonia superposed sampling on B200")
    ap.add_argument("--tb", help="Tuberculosis UNet checkpoint (.pt state_dict)")
    ap.add_argument("--pneumonia", help="Pneumonia UNet checkpoint (.pt state_dict)")
    ap.add_argument("--checkpoint-root", help="reference layout: <root>/<experiment>/<run>/<task>/ema_epoch<N>.pt")
    ap.add_argument("--experiment")
    ap.add_argument("--run")
    ap.add_argument("--epoch", type=int)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--resolution", type=int, default=256)
    ap.add_argument("--steps", type=int, default=250, help="DDPM num_timesteps (ddpm.py:7)")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--temperature", type=float, default=1.0)
    ap.add_argument("--bias", type=float, nargs="*", default=None, help="per-model logit bias l_i")
    ap.add_argument("--mode", choices=["or", "and"], default="or", help="SuperDiff OR (softmax of log q) or AND (equal densities)")
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--out", default="superposed_samples.npz")
    ap.add_argument("--grid", default=None, help="also write a PGM contact sheet here")
    args = ap.parse_args(argv)

    from super_diff_disease_b200 import DDPM, superposed_sample
    paths = [args.tb, args.pneumonia]
    if args.checkpoint_root:
        if not (args.experiment and args.run and args.epoch is not None):
            ap.error("--checkpoint-root needs --experiment, --run and --epoch")
        paths = [checkpoint_path(args.checkpoint_root, args.experiment, args.run, t, args.epoch) for t in ("TB", "PNEUMONIA")]
    if not all(paths):
        ap.error("give --tb and --pneumonia, or --checkpoint-root/--experiment/--run/--epoch")
    for p in paths:
        if not os.path.exists(p):
            raise FileNotFoundError(f"checkpoint not found: {p}")  # same failure the reference raises (train.py:66)
    dev = torch.device(args.device)
    models = [load_unet(p, dev) for p in paths]
    ddpm = DDPM(num_timesteps=args.steps)
    shape = (args.batch, 1, args.resolution, args.resolution)
    bias = torch.tensor(args.bias, dtype=torch.float32) if args.bias else None
    x, kappa, logq = superposed_sample(models, ddpm, shape, dev, seed=args.seed, temperature=args.temperature,
                                       bias=bias, return_trajectory=True, mode=args.mode)
    torch.cuda.synchronize(dev)
    np.savez_compressed(args.out, samples=x.cpu().numpy(), kappa=kappa.cpu().numpy(), logq=logq.cpu().numpy(),
                        checkpoints=np.array(paths), seed=args.seed, steps=args.steps)
    if args.grid:
        save_grid_pgm(x, args.grid)
    k = kappa.float().mean(dim=1)[-1].tolist()
    print(f"wrote {args.out}: samples {tuple(x.shape)}, final mean kappa (TB, PNEUMONIA) = ({k[0]:.3f}, {k[1]:.3f})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
