"""ctypes binding of libsdd_b200.so (include/sdd_b200.h).  Fails loudly; never falls back."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class SddError(RuntimeError):
    pass


def lib_path():
    """The built library; SDD_LIB names an alternative build (same-box A/B of kernel variants, tools/README.md)."""
    return os.environ.get("SDD_LIB") or os.path.join(_HERE, "libsdd_b200.so")


class SampleArgs(ctypes.Structure):
    _fields_ = [("noise_stack", ctypes.c_void_p), ("seed", ctypes.c_uint64), ("sample_offset", ctypes.c_int64),
                ("temperature", ctypes.c_float), ("bias", ctypes.c_void_p), ("x_out", ctypes.c_void_p),
                ("kappa_traj", ctypes.c_void_p), ("logq_traj", ctypes.c_void_p), ("x_traj", ctypes.c_void_p),
                ("use_graph", ctypes.c_int), ("mode", ctypes.c_int),
                ("noise_host", ctypes.c_void_p), ("noise_host_chunk", ctypes.c_int)]


# name -> (restype, argtypes); must list every symbol include/sdd_b200.h declares.
_vp, _i, _i64, _u64, _f, _sz = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_float,
                                ctypes.c_size_t)
SYMBOLS = {
    "sdd_abi_version": (_i, []),
    "sdd_last_error": (ctypes.c_char_p, []),
    "sdd_device_check": (_i, []),
    "sdd_unet_create": (_i, [ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i, _vp]),
    "sdd_unet_destroy": (_i, [_vp]),
    "sdd_unet_set_max_chunk": (_i, [_vp, _i]),
    "sdd_unet_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "sdd_unet_forward_xstats": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "sdd_unet_attn_create": (_i, [ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i, _i, _vp]),
    "sdd_unet_set_label": (_i, [_vp, _i]),
    "sdd_unet_forward_labeled": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "sdd_superpose_update_workspace": (_sz, [_i, _i, _i]),
    "sdd_superpose_update": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _f, _f, _vp, _u64,
                                  _i64, _i, _vp, _sz, _vp]),
    "sdd_superpose_update_and": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _f, _u64, _i64, _i,
                                      _vp, _sz, _vp]),
    "sdd_attention_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    "sdd_attention_block_nhwc": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "sdd_attention_profile": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, ctypes.POINTER(_f), _vp]),
    "sdd_q_sample": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "sdd_mse_workspace": (_sz, []),
    "sdd_mse": (_i, [_vp, _vp, _sz, _vp, _vp, _sz, _vp]),
    "sdd_philox_normal": (_i, [_vp, _i, _i, _u64, _i64, _i, _vp]),
    "sdd_sampler_create": (_i, [ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "sdd_sampler_run": (_i, [_vp, ctypes.POINTER(SampleArgs), _vp]),
    "sdd_sampler_destroy": (_i, [_vp]),
    "sdd_sampler_launches_per_run": (_i64, [_vp]),
    "sdd_sampler_graph_instantiations": (_i64, [_vp]),
    "sdd_conv3x3_fused_nhwc": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "sdd_conv3x3_profile": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _sz, ctypes.POINTER(_f), _vp]),
    "sdd_superpose_update_profile_rotating": (_i, [_i, _i, _i, _i, _i, _sz, ctypes.POINTER(_f), _vp]),
    "sdd_superpose_update_profile": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, ctypes.POINTER(_f), _vp]),
}


def lib():
    """Load the CUDA library.  Raises SddError if it has not been built (no CPU fallback exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise SddError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(needs nvcc); this package has no CPU or PyTorch fallback")
    try:
        L = ctypes.CDLL(path)
    except OSError as e:  # pragma: no cover
        raise SddError(f"cannot load {path}: {e}") from e
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.sdd_abi_version() != 2:
        raise SddError("libsdd_b200.so ABI version mismatch")
    _LIB = L
    return L


def check(rc):
    if rc != 0:
        msg = lib().sdd_last_error()
        raise SddError(f"sdd error {rc}: {msg.decode() if msg else ''}")


def require_cuda(t, name):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise SddError(f"{name} must be a CUDA tensor: the B200 path has no CPU fallback")


def stream_ptr(device):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
