"""Attention core alone at the north star's sizes (32^2 and 16^2 tokens), B = 64 x 2 heads: TFLOP/s = 4 S^2 d BH / t."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import super_diff_disease_b200 as S
dev = torch.device("cuda:0")
L = S.lib()
for BH, S_ in ((128, 1024), (128, 256), (512, 1024)):
    q = torch.randn(BH, S_, 64, device=dev).to(torch.float16)
    k = torch.randn(BH, S_, 64, device=dev).to(torch.float16)
    vt = torch.randn(BH, 64, S_, device=dev).to(torch.float16)
    out = torch.empty_like(q)
    ms = ctypes.c_float()
    rc = L.sdd_attention_profile(q.data_ptr(), k.data_ptr(), vt.data_ptr(), out.data_ptr(), BH, S_, 64, 0.125, 20,
                                 ctypes.byref(ms), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, L.sdd_last_error()
    fl = 4.0 * S_ * S_ * 64 * BH
    print(f"attention BH={BH} S={S_}: {ms.value*1000:.1f} us  {fl/ms.value/1e9:.1f} TFLOP/s ({fl/ms.value/1e9/1626.5:.3f} of bf16 burst)")
