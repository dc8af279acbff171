"""Timing experiments on the product conv at large chunks: dbg variants + per-role clock64 traces (CTA pair 0).
impl codes of sdd_conv3x3_profile: 1/2 = product kernel plain/fused (SDD_CONV_V3=1 selects the single-group loader), 0 = bring-up kernel; dbg bits << 4."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import super_diff_disease_b200 as S
dev = torch.device("cuda:0")
os.makedirs("gpurun_out", exist_ok=True)
chunk = int(os.environ.get("CHUNK", 32))
print("prefetch =", os.environ.get("SDD_CONV_PREFETCH", "default"), "chunk =", chunk)
names = [(0, "full"), (2, "no-store"), (8, "no-stats"), (2 | 8, "no-store/stats"), (4, "no-MMA"), (64, "no-xform-math"), (2 | 8 | 4 | 64, "loads+epi-ld only")]
if os.environ.get("QUICK"):
    names = names[:1]
for cin, cout in [(64, 64), (128, 128), (64, 128), (128, 64)]:
    line = f"{cin:3d}->{cout:3d}: "
    for dbg, nm in names:
        tf, ms = bench.conv_roofline(S, dev, 256, chunk, iters=4, cin=cin, cout=cout, impl=2 + 16 * dbg, flush_l2=True)
        line += f"{nm} {ms*1000:.0f}" + (f" ({tf/1626.5:.3f})" if dbg == 0 else "") + " | "
    tf, ms = bench.conv_roofline(S, dev, 256, chunk, iters=4, cin=cin, cout=cout, impl=1, flush_l2=True)
    line += f"plain {ms*1000:.0f}"
    print(line, flush=True)
if os.environ.get("TRACE", "1") == "1":
    roles = {0: 'prod', 1: 'mma', 2: 'load', 3: 'epi', 4: 'pub'}
    for cin, cout, impl, tag in [(64, 64, 2, "f64"), (128, 128, 2, "f128")]:
        path = f"gpurun_out/trace_{tag}.txt"
        os.environ["SDD_CONV_TRACE"] = path
        tf, ms = bench.conv_roofline(S, dev, 256, chunk, iters=2, cin=cin, cout=cout, impl=impl, flush_l2=True)
        print(tag, f"{ms*1000:.1f} us")
        rows = [list(map(int, l.split())) for l in open(path)]
        for cta in (0, 1):
            d = collections.defaultdict(dict)
            for c, r, i, *ev in rows:
                if r < 5 and c == cta: d[r][i] = ev
            c0 = min(v for r in d for ev in d[r].values() for v in ev if v)
            print(" CTA", cta, "(clock64 is per SM: compare within a CTA only); w12-15 / w16-19 = each loader warp's last arrive of the tile")
            names = {0: 'w12-15', 4: 'w16-19', 2: 'load', 1: 'mma', 3: 'epi'}
            for i in range(5, 11):
                print("  it", i, " | ".join(names[r] + ":" + ",".join(str(v - c0) if v else "-" for v in d[r][i]) for r in (0, 4, 2, 1, 3) if i in d[r]))
    del os.environ["SDD_CONV_TRACE"]
