"""Drop-in ``UNet`` for /root/reference/src/models/unet.py:37-65.

Same constructor, same module tree (hence the same 54 ``state_dict`` keys, so reference checkpoints
``load_state_dict(strict=True)`` unchanged), same ``forward(x, t)`` contract.  The modules below
only HOLD parameters; ``forward`` hands raw device pointers to the C ABI (sdd_unet_forward), which
runs the hand-written sm_100a kernels.  Inference only (sampling runs under ``torch.no_grad()``,
src/train/training_logic.py:54): called in ``train()`` mode with autograd enabled on parameters
that require grad -- i.e. from the training loop, training_logic.py:28-36 -- ``forward`` raises
instead of returning a loss that cannot be back-propagated.
"""
import ctypes

import torch
import torch.nn as nn

from super_diff_disease_b200 import _lib


class SinusoidalPosEmb(nn.Module):
    """Parameter-free placeholder at index 0 of ``time_mlp`` (unet.py:6-16); evaluated on device by
    sinusoid_kernel with the reference's (half-1) divisor and sin-then-cos order."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, t):  # pragma: no cover - never used on the product path
        raise _lib.SddError("SinusoidalPosEmb is evaluated inside sdd_unet_forward; call UNet.forward")


class ResidualBlock(nn.Module):
    """Parameter container for unet.py:18-34: block = [GN, SiLU, Conv3x3, GN, SiLU, Conv3x3], time_emb."""

    def __init__(self, in_ch, out_ch, time_emb_dim):
        super().__init__()
        layers = [nn.GroupNorm(min(4, in_ch), in_ch), nn.SiLU(), nn.Conv2d(in_ch, out_ch, 3, padding=1),
                  nn.GroupNorm(min(4, out_ch), out_ch), nn.SiLU(), nn.Conv2d(out_ch, out_ch, 3, padding=1)]
        self.block = nn.Sequential(*layers)
        self.time_emb = nn.Linear(time_emb_dim, out_ch)

    def forward(self, x, t):  # pragma: no cover
        raise _lib.SddError("ResidualBlock runs fused inside sdd_unet_forward; call UNet.forward")


class _DeviceNet(nn.Module):
    """Parameter container whose forward runs behind the C ABI: owns one sdd_unet_t* for the current parameters."""

    def _init_handle_state(self):
        self._handle = None
        self._handle_key = None
        self._max_chunk = 0

    def _create_handle(self, L, arr, n, stream):  # -> ctypes.c_void_p
        raise NotImplementedError

    # ---- copies (ema_pytorch deep-copies the model, training_logic.py:16; torch.save(model) pickles it) ----------
    # The sdd_unet_t* is a process-local device resource: it is never copied or pickled.  A copy starts without a
    # handle and builds its own lazily, so two modules can never free the same handle.
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_handle"] = None
        state["_handle_key"] = None
        return state

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k in ("_handle", "_handle_key") else copy.deepcopy(v, memo)
        return new

    def set_max_chunk(self, max_samples):
        """Cap the samples one pass of the forward processes at a time (0 = automatic).  Changes no result bit."""
        self._max_chunk = int(max_samples)
        if self._handle is not None:
            _lib.check(_lib.lib().sdd_unet_set_max_chunk(self._handle, self._max_chunk))

    # ---- C-ABI handle management -------------------------------------------------------------
    def _param_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.state_dict(keep_vars=True).values())

    def handle(self):
        """The sdd_unet_t* for the current parameters (rebuilt if they were reloaded or moved)."""
        sd = self.state_dict(keep_vars=True)
        tensors = list(sd.values())
        for k, v in sd.items():
            _lib.require_cuda(v, f"UNet parameter {k}")
        key = self._param_key()
        if self._handle is not None and key == self._handle_key:
            return self._handle
        self._free()
        L = _lib.lib()
        dev = tensors[0].device
        with torch.cuda.device(dev):
            flat = [v.detach().to(torch.float32).contiguous() for v in tensors]
            arr = (ctypes.c_void_p * len(flat))(*[v.data_ptr() for v in flat])
            h = self._create_handle(L, arr, len(flat), _lib.stream_ptr(dev))
            if self._max_chunk:
                _lib.check(L.sdd_unet_set_max_chunk(h, self._max_chunk))
        self._handle, self._handle_key = h, key
        return h

    def _free(self):
        if getattr(self, "_handle", None) is not None:
            try:
                _lib.lib().sdd_unet_destroy(self._handle)
            except Exception:  # pragma: no cover - interpreter shutdown
                pass
            self._handle = None
            self._handle_key = None

    def __del__(self):
        try:
            self._free()
        except Exception:  # interpreter shutdown: torch internals may already be torn down
            pass

    def _refuse_training(self):
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise _lib.SddError(
                "the B200 UNet is forward-only (no autograd graph): it was called in train() mode with gradients enabled "
                "on parameters that require grad, so loss.backward() (training_logic.py:33) could not work.  Train with "
                "the reference modules and build this UNet from the trained / EMA state_dict for sampling; for "
                "evaluation call .eval() or wrap the call in torch.no_grad().")

    def _run_forward(self, x, t, y=None, xstats=None):
        _lib.require_cuda(x, "x")
        if x.dim() != 4 or x.shape[1] != 1:
            raise _lib.SddError(f"x must be [B,1,H,W], got {tuple(x.shape)}")
        B, _, H, W = x.shape
        xc = x.detach().to(torch.float32).contiguous()
        tc = t.to(device=x.device, dtype=torch.int64).contiguous()
        if tc.numel() != B:
            raise _lib.SddError("t must have one entry per sample")
        yc = None
        if y is not None:
            yc = torch.as_tensor(y, device=x.device).to(torch.int64).contiguous()
            if yc.numel() != B:
                raise _lib.SddError("y must have one class label per sample")
        out = torch.empty_like(xc)
        h = self.handle()
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().sdd_unet_forward_labeled(h, xc.data_ptr(), None if xstats is None else xstats.data_ptr(),
                                                           tc.data_ptr(), None if yc is None else yc.data_ptr(),
                                                           out.data_ptr(), B, H, W, _lib.stream_ptr(x.device)))
        return out


class UNet(_DeviceNet):
    def __init__(self, in_channels=1, out_channels=1, time_emb_dim=256, base_channels=64):
        super().__init__()
        if (in_channels, out_channels, time_emb_dim, base_channels) != (1, 1, 256, 64):
            # train.py:88 always default-constructs the UNet; the kernels are specialised to that.
            raise _lib.SddError("the B200 kernels implement the reference's default UNet() only "
                                "(in=1, out=1, time_emb_dim=256, base_channels=64)")
        self.time_mlp = nn.Sequential(SinusoidalPosEmb(time_emb_dim), nn.Linear(time_emb_dim, time_emb_dim * 4),
                                      nn.SiLU(), nn.Linear(time_emb_dim * 4, time_emb_dim))
        c = base_channels
        self.downs = nn.ModuleList([ResidualBlock(in_channels, c, time_emb_dim),
                                    ResidualBlock(c, 2 * c, time_emb_dim)])
        self.mid = ResidualBlock(2 * c, 2 * c, time_emb_dim)
        self.ups = nn.ModuleList([ResidualBlock(2 * c, c, time_emb_dim),
                                  ResidualBlock(c, out_channels, time_emb_dim)])
        self._init_handle_state()

    def _create_handle(self, L, arr, n, stream):
        h = ctypes.c_void_p()
        _lib.check(L.sdd_unet_create(ctypes.byref(h), arr, n, stream))
        return h

    def forward(self, x, t):
        """x fp32 [B,1,H,W] (CUDA), t int64 [B] -> predicted noise fp32 [B,1,H,W] (unet.py:57-65)."""
        self._refuse_training()
        with torch.no_grad():
            return self._run_forward(x, t)

    def _forward(self, x, t, xstats=None):
        """xstats (internal): fp32 [B,2] (mean, rstd) of each x sample from the update kernel that produced x."""
        return self._run_forward(x, t, None, xstats)


class AttnBlock(nn.Module):
    """Parameter container of one pre-norm self-attention block (extension N2): out = x + proj(attention(q, k, v)),
    [q | k | v] = qkv(GroupNorm(4, C)(x)), heads of 64 channels.  Runs fused inside sdd_unet_forward."""

    def __init__(self, channels=128):
        super().__init__()
        self.norm = nn.GroupNorm(4, channels)
        self.qkv = nn.Linear(channels, 3 * channels)
        self.proj = nn.Linear(channels, channels)

    def forward(self, x):  # pragma: no cover
        raise _lib.SddError("AttnBlock runs fused inside sdd_unet_forward; call UNetAttn.forward")


class UNetAttn(_DeviceNet):
    """EXTENSION -- the reference has no such module (its UNet, unet.py:37-65, is five full-resolution blocks with no
    attention, resampling, skip connections or class input); the oracle is this repo's oracle/unet_attn_oracle.py.

    Class-conditional multi-resolution UNet with self-attention at R/8 and R/16 (32^2 and 16^2 at R = 256: BASELINE
    configs[2] as worded), built from the reference's own ResidualBlock so that every conv runs on the reference path's
    kernels.  Architecture: include/sdd_b200.h (sdd_unet_attn_create).  forward(x, t, y=None): y int64 [B] class labels
    (None: the label set with set_label, default 0).  Resolutions: H % 256 == 0, W % 128 == 0."""

    def __init__(self, in_channels=1, out_channels=1, time_emb_dim=256, base_channels=64, num_classes=2):
        super().__init__()
        if (in_channels, out_channels, time_emb_dim, base_channels) != (1, 1, 256, 64):
            raise _lib.SddError("UNetAttn is specialised like the reference default (in=1, out=1, time_emb_dim=256, "
                                "base_channels=64)")
        c, d = base_channels, time_emb_dim
        self.num_classes = int(num_classes)
        self.time_mlp = nn.Sequential(SinusoidalPosEmb(d), nn.Linear(d, d * 4), nn.SiLU(), nn.Linear(d * 4, d))
        self.class_emb = nn.Embedding(self.num_classes, d)
        self.enc = nn.ModuleList([ResidualBlock(in_channels, c, d), ResidualBlock(c, 2 * c, d), ResidualBlock(2 * c, 2 * c, d),
                                  ResidualBlock(2 * c, 2 * c, d), ResidualBlock(2 * c, 2 * c, d)])
        self.mid = ResidualBlock(2 * c, 2 * c, d)
        self.dec = nn.ModuleList([ResidualBlock(2 * c, 2 * c, d), ResidualBlock(2 * c, 2 * c, d), ResidualBlock(2 * c, c, d),
                                  ResidualBlock(c, out_channels, d)])
        self.attn = nn.ModuleList([AttnBlock(2 * c) for _ in range(4)])  # after enc.3, enc.4, mid, dec.0
        self.label = 0
        self._init_handle_state()

    def _create_handle(self, L, arr, n, stream):
        h = ctypes.c_void_p()
        _lib.check(L.sdd_unet_attn_create(ctypes.byref(h), arr, n, self.num_classes, stream))
        _lib.check(L.sdd_unet_set_label(h, int(self.label)))
        return h

    def set_label(self, label):
        """The class a sampler built on this model (and a forward without y) conditions on."""
        label = int(label)
        if not 0 <= label < self.num_classes:
            raise _lib.SddError(f"label must be in [0, {self.num_classes})")
        self.label = label
        if self._handle is not None:
            _lib.check(_lib.lib().sdd_unet_set_label(self._handle, label))
        return self

    def forward(self, x, t, y=None):
        """x fp32 [B,1,H,W] (CUDA), t int64 [B], y int64 [B] or None -> predicted noise fp32 [B,1,H,W]."""
        self._refuse_training()
        with torch.no_grad():
            return self._run_forward(x, t, y)

    def _forward(self, x, t, xstats=None):
        return self._run_forward(x, t, None, xstats)
