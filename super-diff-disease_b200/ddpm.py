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

    # ---- training-side forward pieces: evaluation only, no autograd graph is built
    @torch.no_grad()
    def q_sample(self, x_start, t, noise=None):
        """Forward noising (ddpm.py:13-17) on the GPU; bit-identical to the reference for the same inputs."""
        _lib.require_cuda(x_start, "x_start")
        if noise is None:
            noise = torch.randn_like(x_start)  # same draw as ddpm.py:15
        dev = x_start.device
        # coefficients by the reference's own ops on the CPU fp32 table (ddpm.py:16-17), then one fused pass
        ab = self.alpha_bars[t.detach().cpu().long()]
        ca = torch.sqrt(ab).to(dev).contiguous()
        cb = torch.sqrt(1 - ab).to(dev).contiguous()
        x0 = x_start.to(torch.float32).contiguous()
        nz = noise.to(torch.float32).contiguous()
        out = torch.empty_like(x0)
        B, D = x0.shape[0], x0[0].numel()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().sdd_q_sample(x0.data_ptr(), nz.data_ptr(), ca.data_ptr(), cb.data_ptr(),
                                               out.data_ptr(), B, D, _lib.stream_ptr(dev)))
        return out

    def p_losses(self, denoise_model, x_start, t, noise=None):
        """eps-MSE of ddpm.py:20-24, forward only (validation loss): q_sample -> UNet forward on the sampler's
        kernels -> deterministic MSE reduction.  Returns a 0-dim CUDA tensor WITHOUT a grad_fn; called from a training
        loop (gradients enabled on the model's parameters) it raises.  ``noise=`` pins the draw for parity."""
        if hasattr(denoise_model, "_refuse_training"):
            denoise_model._refuse_training()
        with torch.no_grad():
            return self._p_losses(denoise_model, x_start, t, noise)

    def _p_losses(self, denoise_model, x_start, t, noise=None):
        _lib.require_cuda(x_start, "x_start")
        if noise is None:
            noise = torch.randn_like(x_start)
        noise = noise.to(torch.float32).contiguous()
        x_noisy = self.q_sample(x_start, t, noise)
        pred = denoise_model(x_noisy, t).contiguous()
        L = _lib.lib()
        dev = x_start.device
        ws = torch.empty(L.sdd_mse_workspace(), dtype=torch.uint8, device=dev)
        out = torch.empty((), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.sdd_mse(pred.data_ptr(), noise.data_ptr(), pred.numel(), out.data_ptr(), ws.data_ptr(),
                                 ws.numel(), _lib.stream_ptr(dev)))
        return out

    def training_step(self, model, x):
        """ddpm.py:26-29 (same t draw), forward only: the value of the training loss, not a differentiable graph
        (raises when called with gradients enabled on the model's parameters, i.e. from training_logic.py:32)."""
        if hasattr(model, "_refuse_training"):
            model._refuse_training()
        bsz = x.size(0)
        t = torch.randint(0, self.T, (bsz,), device=x.device).long()
        return self.p_losses(model, x, t)

    def draw_noise_stack(self, image_shape, device):
        """[T, *image_shape] noise with the reference's draw order and generators (ddpm.py:33,36).  O(T * B * H * W)
        memory: tests and small shapes only -- ``sample`` itself draws per step."""
        x = torch.randn(image_shape).to(device)
        stack = torch.empty((self.T,) + tuple(image_shape), dtype=torch.float32, device=device)
        stack[0] = x
        for k in range(1, self.T):
            stack[k] = torch.randn_like(x)
        return stack

    def _sample_torch_rng(self, model, image_shape, device):
        """The default call: torch's generators, consumed exactly like ddpm.py:33,36 (x_T on the CPU generator, then one
        device ``randn_like`` per step with t > 0, drawn before the model call), with O(B * H * W) memory like the
        reference: the loop runs step by step through the operator entry points (UNet forward + fused update)."""
        x = torch.randn(image_shape).to(device)
        B = x.shape[0]
        logq = torch.zeros(B, 1, dtype=torch.float32, device=device)
        ws = xstats = None
        for t in reversed(range(self.T)):
            noise = torch.randn_like(x) if t > 0 else None
            t_tensor = torch.full((B,), t, dtype=torch.long, device=device)
            eps = model._forward(x, t_tensor, xstats) if hasattr(model, "_forward") else model(x, t_tensor)
            if ws is None:
                nbytes = _lib.lib().sdd_superpose_update_workspace(B, x[0].numel(), 1)
                ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
            x, logq, _, xstats = sampling.superpose_update(x, eps.unsqueeze(0), logq, self.alphas[t].item(),
                                                      self.alpha_bars[t].item(), self.betas[t].item(), noise=noise,
                                                      workspace=ws)
        return x

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
            return self._sample_torch_rng(model, image_shape, device)
        return sampling.superposed_sample([model], self, image_shape, device, noise=noise, seed=seed,
                                          use_graph=use_graph)
