import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import super_diff_disease_b200 as S
dev = torch.device("cuda:0")
for cin, cout in [(128, 128), (64, 64)]:
    line = f"{cin}->{cout} fused: "
    for dbg, nm in [(0, "full"), (64, "skip-transform-body"), (128, "LDS+STS copy only"), (144, "no math no smem")]:
        tf, ms = bench.conv_roofline(S, dev, 256, 3, iters=10, cin=cin, cout=cout, impl=2 + 16 * dbg, flush_l2=False)
        line += f"{nm} {ms*1000:.1f} | "
    print(line)
