"""Superposed reverse-diffusion sampling -- the module the reference leaves empty (src/sampling.py is
0 bytes; README.md:5,9 describe it).  Spec: SURVEY.md section 8(a) row A7 (SuperDiff "OR" with the Ito
density estimator in the reference's DDPM variables; x update byte-identical to ddpm.py:42-44).

    x, kappas, logq = superposed_sample([unet_tb, unet_pn], ddpm, (B,1,H,W), "cuda",
                                        seed=1234, return_trajectory=True)

Python only allocates tensors and passes pointers; the loop (2 UNet forwards + one fused update
kernel per step, replayed as a CUDA graph) lives in libsdd_b200.so (sdd_sampler_run).
"""
import ctypes
import weakref

import torch

from super_diff_disease_b200 import _lib

_SAMPLERS = {}


class _Sampler:
    """Owns one sdd_sampler_t for (models, schedule, B, H, W) on one device."""

    def __init__(self, models, ddpm, B, H, W, device):
        L = _lib.lib()
        self.handles = [m.handle() for m in models]
        self.T, self.B, self.H, self.W, self.M = ddpm.T, B, H, W, len(models)
        arr = (ctypes.c_void_p * self.M)(*[h.value for h in self.handles])
        a = ddpm.alphas.to(torch.float32).contiguous().cpu()
        ab = ddpm.alpha_bars.to(torch.float32).contiguous().cpu()
        b = ddpm.betas.to(torch.float32).contiguous().cpu()
        self.ptr = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(L.sdd_sampler_create(ctypes.byref(self.ptr), arr, self.M, a.data_ptr(), ab.data_ptr(),
                                            b.data_ptr(), self.T, B, H, W, _lib.stream_ptr(device)))
        self.pending = []  # (event, tensors) of calls whose asynchronous work may still be reading the tensors

    def launches(self):
        return int(_lib.lib().sdd_sampler_launches_per_run(self.ptr))

    def graph_instantiations(self):
        return int(_lib.lib().sdd_sampler_graph_instantiations(self.ptr))

    def close(self):
        if self.ptr:
            _lib.lib().sdd_sampler_destroy(self.ptr)
            self.ptr = None


def _get_sampler(models, ddpm, B, H, W, device):
    key = (tuple(id(m) for m in models), tuple(m._param_key() for m in models), ddpm.T,
           ddpm.betas.numpy().tobytes(), B, H, W, tuple(getattr(m, "label", None) for m in models), str(device))
    s = _SAMPLERS.get(key)
    if s is None:
        stale = [k for k in _SAMPLERS if k[0] == key[0] and k[-1] == key[-1]]
        for k in stale:  # same models, different shape/weights: free the old workspaces first
            _SAMPLERS.pop(k).close()
        s = _Sampler(models, ddpm, B, H, W, device)
        _SAMPLERS[key] = s
        for m in models:
            weakref.finalize(m, _drop_samplers_for, id(m))
    return s


def _drop_samplers_for(model_id):
    for k in [k for k in _SAMPLERS if model_id in k[0]]:
        try:
            _SAMPLERS.pop(k).close()
        except Exception:  # pragma: no cover
            pass


def clear_cache():
    for k in list(_SAMPLERS):
        _SAMPLERS.pop(k).close()


_MODES = {"or": 0, "and": 1}


@torch.no_grad()
def superposed_sample(models, ddpm, image_shape, device, *, temperature=1.0, bias=None, noise=None, seed=None,
                      return_trajectory=False, use_graph=True, sample_offset=0, return_launches=False, mode="or",
                      return_x_trajectory=False, noise_chunk_steps=0):
    """Sample from the superposition of ``models`` (one model == DDPM.sample).

    models: sequence of super_diff_disease_b200.UNet (or the UNetAttn extension, each conditioning on the class set
    with ``set_label``) on ``device``; ddpm: DDPM (schedule);
    image_shape: (B, 1, H, W), H % 16 == 0, W % 8 == 0.
    noise: fp32 [T, B, 1, H, W] stack (parity mode), or seed: int for in-kernel Philox keyed by
    (seed, sample_offset + b, draw, element) -- shard-invariant.  Exactly one of the two.
    A noise stack on the HOST (CPU tensor; pinned, or it is pinned here) is STREAMED: the library copies it to the
    device in chunks of ``noise_chunk_steps`` slices (0 = ~64 MB) on its own copy stream while earlier steps compute,
    through a two-chunk device ring -- the [T,B,1,H,W] stack never exists on the device and the host->device copy
    overlaps the loop.  Results are bit-identical to passing the same stack as a CUDA tensor.
    mode: "or" (kappa = softmax(temperature * log q + bias)) or "and" (kappa solved per sample and
    step so that every model's log-density increment is equal; temperature / bias unused).
    Returns x [B,1,H,W]; with return_trajectory also kappas [T,B,M] and logq [T+1,B,M]
    (row k <-> loop iteration k, t = T-1-k); with return_x_trajectory additionally xs [T+1,B,1,H,W] (xs[0] = x_T,
    xs[k+1] = the state after iteration k; the strip utils/visualization.py:6-28 plots).
    """
    if mode not in _MODES:
        raise _lib.SddError(f"mode must be 'or' or 'and', got {mode!r}")
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.SddError("superposed_sample runs on a CUDA device only (no CPU fallback)")
    if (noise is None) == (seed is None):
        raise _lib.SddError("pass exactly one of noise= (explicit stack) or seed= (in-kernel Philox)")
    B, C, H, W = image_shape
    if C != 1:
        raise _lib.SddError("the reference UNet is single-channel (in_channels=1)")
    models = list(models)
    M, T = len(models), ddpm.T
    for m in models:
        m.eval()
    s = _get_sampler(models, ddpm, B, H, W, device)
    args = _lib.SampleArgs()
    keep = []
    if noise is not None:
        if not isinstance(noise, torch.Tensor):
            raise _lib.SddError("noise must be a torch.Tensor")
        if tuple(noise.shape) != (T, B, 1, H, W):
            raise _lib.SddError(f"noise must be [T,B,1,H,W] = {(T, B, 1, H, W)}, got {tuple(noise.shape)}")
        noise = noise.to(torch.float32).contiguous()
        if noise.is_cuda:
            args.noise_stack = noise.data_ptr()
        else:  # host stack: streamed by the library (pinned memory is a requirement of asynchronous copies)
            if not noise.is_pinned():
                noise = noise.pin_memory()
            args.noise_stack = None
            args.noise_host = noise.data_ptr()
            args.noise_host_chunk = int(noise_chunk_steps)
        keep.append(noise)
    else:
        args.noise_stack = None
        args.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    args.sample_offset = int(sample_offset)
    args.temperature = float(temperature)
    if bias is not None:
        bias = torch.as_tensor(bias, dtype=torch.float32, device=device).contiguous()
        if bias.numel() != M:
            raise _lib.SddError("bias must have one entry per model")
        keep.append(bias)
        args.bias = bias.data_ptr()
    x = torch.empty((B, 1, H, W), dtype=torch.float32, device=device)
    args.x_out = x.data_ptr()
    kap = lq = None
    if return_trajectory:
        kap = torch.empty((T, B, M), dtype=torch.float32, device=device)
        lq = torch.empty((T + 1, B, M), dtype=torch.float32, device=device)
        args.kappa_traj, args.logq_traj = kap.data_ptr(), lq.data_ptr()
        keep += [kap, lq]
    xs = None
    if return_x_trajectory:
        xs = torch.empty((T + 1, B, 1, H, W), dtype=torch.float32, device=device)
        args.x_traj = xs.data_ptr()
        keep.append(xs)
    args.use_graph = 1 if use_graph else 0
    args.mode = _MODES[mode]
    with torch.cuda.device(device):
        _lib.check(_lib.lib().sdd_sampler_run(s.ptr, ctypes.byref(args), _lib.stream_ptr(device)))
    # the streams may still be reading these (the noise stack, possibly in pinned host memory that the library copies from
    # asynchronously): they stay referenced until an event recorded behind the call has completed
    done = torch.cuda.Event()
    done.record(torch.cuda.current_stream(device))
    s.pending = [(e, k) for (e, k) in s.pending if not e.query()] + [(done, keep)]
    out = (x, kap, lq) if return_trajectory else x
    if return_x_trajectory:
        out = (out + (xs,)) if isinstance(out, tuple) else (out, xs)
    if return_launches:
        return out, s.launches()
    return out


@torch.no_grad()
def superpose_update(x, eps, logq, alpha, alpha_bar, beta, *, noise=None, seed=None, sample_offset=0, draw_index=0,
                     temperature=1.0, bias=None, out=None, workspace=None, mode="or"):
    """One fused superposition update (operator form of A7; wraps sdd_superpose_update / sdd_superpose_update_and).

    x [B,...] fp32, eps [M,B,...] fp32, logq [B,M] fp32.  noise [B,...] or seed (Philox) or neither
    (z = 0, the t == 0 step).  Returns (x_new, logq_new, kappa[B,M], xstats[B,2]).
    """
    _lib.require_cuda(x, "x")
    L = _lib.lib()
    B = x.shape[0]
    D = x[0].numel()
    M = eps.shape[0]
    xc = x.to(torch.float32).contiguous()
    ec = eps.to(torch.float32).contiguous()
    lq = logq.to(torch.float32).contiguous()
    x_new = out if out is not None else torch.empty_like(xc)
    lq_new = torch.empty_like(lq)
    kap = torch.empty_like(lq)
    xst = torch.empty((B, 2), dtype=torch.float32, device=x.device)
    nbytes = L.sdd_superpose_update_workspace(B, D, M)
    ws = workspace if workspace is not None else torch.zeros(nbytes, dtype=torch.uint8, device=x.device)
    nptr = None
    if noise is not None:
        noise = noise.to(torch.float32).contiguous()
        nptr = noise.data_ptr()
    di = draw_index if (seed is not None) else -1
    bptr = None
    if bias is not None:
        bias = torch.as_tensor(bias, dtype=torch.float32, device=x.device).contiguous()
        bptr = bias.data_ptr()
    if mode not in _MODES:
        raise _lib.SddError(f"mode must be 'or' or 'and', got {mode!r}")
    if mode == "and":
        with torch.cuda.device(x.device):
            _lib.check(L.sdd_superpose_update_and(xc.data_ptr(), x_new.data_ptr(), ec.data_ptr(), nptr, lq.data_ptr(),
                                                  lq_new.data_ptr(), kap.data_ptr(), xst.data_ptr(), B, D, M,
                                                  float(alpha), float(alpha_bar), float(beta),
                                                  int(seed or 0) & 0xFFFFFFFFFFFFFFFF, int(sample_offset), int(di),
                                                  ws.data_ptr(), ws.numel(), _lib.stream_ptr(x.device)))
        return x_new, lq_new, kap, xst
    with torch.cuda.device(x.device):
        _lib.check(L.sdd_superpose_update(xc.data_ptr(), x_new.data_ptr(), ec.data_ptr(), nptr, lq.data_ptr(),
                                          lq_new.data_ptr(), kap.data_ptr(), xst.data_ptr(), B, D, M, float(alpha),
                                          float(alpha_bar), float(beta), float(temperature), bptr,
                                          int(seed or 0) & 0xFFFFFFFFFFFFFFFF, int(sample_offset), int(di),
                                          ws.data_ptr(), ws.numel(), _lib.stream_ptr(x.device)))
    return x_new, lq_new, kap, xst
