"""Per-role cycle accounting of the tensor-core conv kernel (CTA 0) from a -DSDD_CONV_PROF build:
    python -c "import sys; sys.path.insert(0,'super-diff-disease_b200'); import build; build.build(force=True, out='tools/_lib_prof.so', flags=['-DSDD_CONV_PROF'])"
    SDD_LIB=tools/_lib_prof.so python tools/conv_prof.py
Prints, per layer and launch mode, the cycles per 64-channel item each role spends in each phase."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import super_diff_disease_b200 as S
dev = torch.device("cuda:0")
chunk, res = int(os.environ.get("CHUNK", 64)), int(os.environ.get("RES", 256))
lib = S.lib()
buf = (ctypes.c_ulonglong * 64)()
lib.sdd_conv_prof_read.argtypes = [ctypes.c_void_p]
ITERS = 3
MMA = ["wait_tempty", "wait_ready", "issue", "loop"]
EPI = ["wait_tfull", "tmem->regs", "stats+stores", "loop"]
LDR = ["raw_lds", "wait_empty", "load_stall", "xform+sts", "fence", "next+arrive", "-", "loop_top"]
for cin, cout in [(64, 64), (128, 64), (64, 128), (128, 128)]:
    for impl, name in ((2, "fused"), (1, "plain")):
        lib.sdd_conv_prof_read(buf)
        tf, ms = bench.conv_roofline(S, dev, res, chunk, iters=ITERS, cin=cin, cout=cout, impl=impl, flush_l2=True)
        lib.sdd_conv_prof_read(buf)
        v = list(buf)
        tiles = chunk * (res // 16) * (res // 8)
        pairs_cta = (tiles // 2) / 74.0
        items = pairs_cta * (cin // 64) * ITERS
        f = lambda names, base, n_items: "  ".join(f"{nm} {v[base + i] / n_items:7.0f}" for i, nm in enumerate(names) if nm != "-")
        print(f"{cin}->{cout} {name}: {ms * 1000:.1f} us; items/CTA {items / ITERS:.0f}; cycles per item:")
        print("   mma      ", f(MMA, 0, items), " sum", f"{sum(v[0:4]) / items:.0f}")
        print("   epilogue ", f(EPI, 8, items), " sum", f"{sum(v[8:12]) / items:.0f}", "(per item; x%d per tile)" % (cin // 64))
        for g in range(2):
            print(f"   loader g{g}", f(LDR, 16 + 8 * g, items / 2), " sum", f"{sum(v[16 + 8 * g:24 + 8 * g]) / (items / 2):.0f}", "(per item of the group)")
