import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import super_diff_disease_b200 as S
dev = torch.device("cuda:0")
os.makedirs("gpurun_out", exist_ok=True)
roles = {0: 'prod', 1: 'mma', 2: 'xform', 3: 'epi', 4: 'pub'}
for cin, cout, impl, tag in [(128, 128, 1, "plain"), (128, 128, 2, "fused"), (128, 128, 1 + 16 * 2, "plain-nostore")]:
    path = f"gpurun_out/trace_{tag}.txt"
    os.environ["SDD_CONV_TRACE"] = path
    tf, ms = bench.conv_roofline(S, dev, 256, 3, iters=3, cin=cin, cout=cout, impl=impl, flush_l2=False)
    print(tag, f"{ms*1000:.1f} us")
    rows = [list(map(int, l.split())) for l in open(path)]
    g = [r for r in rows if r[1] == 5]
    t0 = min(r[3] for r in g)
    ex = sorted((r[6] - t0, r[5] - t0, r[4] - t0) for r in g)
    print("  per-CTA (exit, loopdone, setup) ns: min", ex[0], "median", ex[len(ex) // 2], "max", ex[-1])
    d = collections.defaultdict(dict)
    for c, r, i, *ev in rows:
        if r < 5 and c == 0: d[r][i] = ev
    c0 = min(v for r in d for ev in d[r].values() for v in ev if v)
    for i in (3, 4, 5, 6, 7):
        print("  it", i, " | ".join(roles[r] + ":" + ",".join(str(v - c0) if v else "-" for v in d[r][i]) for r in (0, 2, 1, 3, 4) if i in d[r]))
