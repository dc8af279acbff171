"""Kernel-only timing of the conv variants at the sampler's launch shape (CUDA events, L2 flushed)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import super_diff_disease_b200 as S
dev = torch.device("cuda:0")
R, chunk = int(os.environ.get("R", 256)), int(os.environ.get("CHUNK", 3))
for cin, cout in [(128, 128), (64, 128), (128, 64), (64, 64)]:
    for impl in (0, 1, 2):
        tf, ms = bench.conv_roofline(S, dev, R, chunk, iters=10, cin=cin, cout=cout, impl=impl)
        print(f"{cin:3d}->{cout:3d} impl={impl} {ms*1000:8.1f} us  {tf:7.1f} TFLOP/s  ({tf/1626.5:.3f} of burst)")
for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512):
    gb, ms = bench.update_roofline(S, dev, B, 65536, iters=10)
    gb2, ms2 = bench.update_roofline(S, dev, B, 65536, iters=10, noise=True)
    print(f"update B={B:4d}: philox {ms*1000:7.1f} us {gb:7.1f} GB/s ({gb/6545.9:.3f}) | noise tensor {ms2*1000:7.1f} us {gb2:7.1f} GB/s ({gb2/6545.9:.3f})")
