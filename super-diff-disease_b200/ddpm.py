"""Drop-in ``DDPM`` for /root/reference/src/models/ddpm.py:6-45 (schedule + ancestral sampler).

``sample(model, image_shape, device)`` keeps the reference's positional contract and its RNG
consumption (x_T from the CPU generator, ddpm.py:33; one device ``randn_like`` per step with t > 0,
ddpm.py:36) so ``torch.manual_seed`` reproduces the reference's noise.  The loop itself -- UNet
forward, x update -- runs in the C library (sdd_sampler_run with one model).
"""
import torch

from super_diff_disease_b200 import _lib, sampling


class DDPM:
    def __init__(self, num_timesteps=1000, beta_start=1e-4, beta_end=0.02):
        self.T = num_timesteps
        # CPU fp32 tensors computed by the same torch ops as ddpm.py:9-11 => bit-identical tables
        self.betas = torch.linspace(beta_start, beta_end, self.T)
        self.alphas = 1.0 - self.betas
        self.alpha_bars = torch.cumprod(self.alphas, dim=0)

    def draw_noise_stack(self, image_shape, device):
        """[T, *image_shape] noise with the reference's draw order and generators (ddpm.py:33,36)."""
        x = torch.randn(image_shape).to(device)
        stack = torch.empty((self.T,) + tuple(image_shape), dtype=torch.float32, device=device)
        stack[0] = x
        for k in range(1, self.T):
            stack[k] = torch.randn_like(x)
        return stack

    @torch.no_grad()
    def sample(self, model, image_shape, device, *, noise=None, seed=None, use_graph=True):
        """Reverse diffusion (ddpm.py:31-45).  Extra keyword-only arguments:
        noise: explicit fp32 stack [T, *image_shape] (stack[0] = x_T, stack[k] = k-th in-loop draw);
        seed:  use the in-kernel Philox generator instead of torch's RNG (throughput mode)."""
        model.eval()
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.SddError("DDPM.sample runs on a CUDA device only (no CPU fallback); "
                                "the CPU reference lives in oracle/ for tests")
        if noise is None and seed is None:
            noise = self.draw_noise_stack(image_shape, device)
        return sampling.superposed_sample([model], self, image_shape, device, noise=noise, seed=seed,
                                          use_graph=use_graph)
