"""Fused update kernel at one batch size (for ncu): B from the environment."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import super_diff_disease_b200 as S
dev = torch.device("cuda:0")
B = int(os.environ.get("B", 64))
for noise in (False, True):
    gb, ms = bench.update_roofline(S, dev, B, 65536, iters=int(os.environ.get("ITERS", 10)), noise=noise)
    print(f"update B={B} {'noise tensor (20 B/el)' if noise else 'philox (16 B/el)'}: {ms*1000:.1f} us {gb:.0f} GB/s ({gb/6545.9:.3f})")
