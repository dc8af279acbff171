import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import super_diff_disease_b200 as S
dev = torch.device("cuda:0")
for cin, cout in [(128, 128), (64, 64), (64, 128), (128, 64)]:
    for chunk in [int(c) for c in os.environ.get("CHUNKS", "2 3 4 8 16 32 64").split()]:
        res = []
        for impl in (1, 2):
            tf, ms = bench.conv_roofline(S, dev, 256, chunk, iters=5, cin=cin, cout=cout, impl=impl, flush_l2=True)
            res.append(f"{'plain' if impl == 1 else 'FUSED'} {ms*1000:8.1f} us {tf:6.0f} TF/s ({tf/1626.5:.3f})")
        print(f"{cin:3d}->{cout:3d} chunk={chunk:2d}: " + " | ".join(res))
