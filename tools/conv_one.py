"""One configuration of the product conv (for ncu): CIN COUT CHUNK IMPL from the environment."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import super_diff_disease_b200 as S
dev = torch.device("cuda:0")
cin, cout = int(os.environ.get("CIN", 128)), int(os.environ.get("COUT", 128))
chunk, impl = int(os.environ.get("CHUNK", 32)), int(os.environ.get("IMPL", 2))
tf, ms = bench.conv_roofline(S, dev, 256, chunk, iters=int(os.environ.get("ITERS", 3)), cin=cin, cout=cout, impl=impl, flush_l2=True)
print(f"{cin}->{cout} chunk={chunk} impl={impl}: {ms*1000:.1f} us {tf:.0f} TF/s ({tf/1626.5:.3f})")
