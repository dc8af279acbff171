import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import super_diff_disease_b200 as S
dev = torch.device("cuda:0")
names = {0: "full", 2: "no-store", 4: "no-MMA", 6: "no-MMA,no-store", 22: "no-MMA,no-store,no-TMA"}
for cin, cout in [(128, 128), (64, 128), (128, 64), (64, 64)]:
    for fl in (True, False):
        line = f"{cin:3d}->{cout:3d} flush={int(fl)}: "
        for dbg, nm in names.items():
            tf, ms = bench.conv_roofline(S, dev, 256, 3, iters=10, cin=cin, cout=cout, impl=1 + 16 * dbg, flush_l2=fl)
            line += f"{nm} {ms*1000:.1f} | "
        tf, ms = bench.conv_roofline(S, dev, 256, 3, iters=10, cin=cin, cout=cout, impl=2, flush_l2=fl)
        line += f"FUSED {ms*1000:.1f} us = {tf:.0f} TF/s ({tf/1626.5:.3f}) | "
        tf, ms = bench.conv_roofline(S, dev, 256, 3, iters=10, cin=cin, cout=cout, impl=0, flush_l2=fl)
        line += f"v1 {ms*1000:.1f}"
        print(line)
