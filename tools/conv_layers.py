"""The four tensor-core conv layers (+ the plain, un-fused launch) at the sampler's launch shape: us and fraction of the
measured bf16 burst.  CHUNK (default 64), RES (256); SDD_LIB selects an alternative build of the library (A/B)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import super_diff_disease_b200 as S
dev = torch.device("cuda:0")
chunk, res = int(os.environ.get("CHUNK", 64)), int(os.environ.get("RES", 256))
peak = bench.peaks()[1]
print("lib", S.lib_path(), "chunk", chunk, "res", res)
tot = 0.0
for cin, cout, n in [(128, 128, 3), (64, 128, 1), (128, 64, 1), (64, 64, 2)]:
    tf, ms = bench.conv_roofline(S, dev, res, chunk, iters=6, cin=cin, cout=cout, impl=2, flush_l2=True)
    tfp, msp = bench.conv_roofline(S, dev, res, chunk, iters=4, cin=cin, cout=cout, impl=1, flush_l2=True)
    tot += n * ms
    print(f"{cin:3d}->{cout:3d}: fused {ms*1000:7.1f} us ({tf/peak:.3f}) | plain {msp*1000:7.1f} us ({tfp/peak:.3f})", flush=True)
fl = 2.0 * 9 * (3 * 128 * 128 + 2 * 64 * 128 + 2 * 64 * 64) * chunk * res * res
print(f"seven tensor-core convs of one forward: {tot*1000:.0f} us = {fl/tot/1e9/peak:.3f} of burst")
