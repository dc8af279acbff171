"""Experiment: does UMMA accept 128B-aligned (not 1024B-aligned) start addresses and SBO=1280 with SWIZZLE_128B?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import super_diff_disease_b200 as S
dev = torch.device("cuda:0")
for Cin, Cout in [(64, 64), (128, 128), (64, 128), (128, 64)]:
    for B, H, W in [(1, 16, 8), (2, 32, 24), (2, 64, 64)]:
        g = torch.Generator().manual_seed(1)
        act = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16).to(dev)
        w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).to(dev)
        bias = torch.randn(Cout, generator=g).to(dev)
        outs = []
        for impl in (0, 2):
            out = torch.zeros(B, H, W, Cout, dtype=torch.bfloat16, device=dev)
            rc = S.lib().sdd_conv3x3_nhwc(act.data_ptr(), w.data_ptr(), bias.data_ptr(), 0, out.data_ptr(), None,
                                          B, H, W, Cin, Cout, impl, None)
            assert rc == 0, S.lib().sdd_last_error()
            torch.cuda.synchronize()
            outs.append(out.float())
        d = (outs[0] - outs[1]).abs().max().item()
        print(f"Cin={Cin} Cout={Cout} B={B} H={H} W={W}: halo vs baseline max|diff| = {d:.4g}  equal={torch.equal(outs[0], outs[1])}")
