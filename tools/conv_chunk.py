"""Product conv vs chunk size, with the L2 flushed before every launch and with the working set left in L2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import super_diff_disease_b200 as S
dev = torch.device("cuda:0")
for cin, cout in [(128, 128), (64, 64), (64, 128), (128, 64)]:
    for chunk in [int(c) for c in os.environ.get("CHUNKS", "2 3 4 6 8 16 32").split()]:
        res = []
        for fl in (True, False):
            tf, ms = bench.conv_roofline(S, dev, 256, chunk, iters=6, cin=cin, cout=cout, impl=2, flush_l2=fl)
            res.append(f"{'flushed' if fl else 'L2-warm'} {ms*1000:7.1f} us {tf:6.0f} TF/s ({tf/1626.5:.3f}) {ms*1000/chunk:6.2f} us/sample")
        print(f"{cin:3d}->{cout:3d} chunk={chunk:2d}: " + " | ".join(res), flush=True)
