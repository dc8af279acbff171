"""B200-native SuperDiff sampling hot path (drop-in for mo-rsa24/super-diff-disease's sampler).

    from super_diff_disease_b200 import UNet, DDPM, superposed_sample

Python here is plumbing only (device memory, streams, torch.distributed); every arithmetic step of
the sampling loop runs in hand-written sm_100a CUDA behind the C ABI in include/sdd_b200.h.
There is no CPU fallback: without a B200 and the built library the calls raise.
"""
from super_diff_disease_b200._lib import SddError, lib, lib_path  # noqa: F401
from super_diff_disease_b200.unet import UNet, UNetAttn, AttnBlock, SinusoidalPosEmb, ResidualBlock  # noqa: F401
from super_diff_disease_b200.ddpm import DDPM  # noqa: F401
from super_diff_disease_b200.sampling import superposed_sample, superpose_update  # noqa: F401
from super_diff_disease_b200.dist import shard_range, sharded_sample  # noqa: F401
from super_diff_disease_b200.attention import attention_core, attention_block  # noqa: F401

__all__ = ["UNet", "UNetAttn", "DDPM", "superposed_sample", "superpose_update", "sharded_sample", "shard_range", "attention_core", "attention_block",
           "SddError", "lib", "lib_path"]
