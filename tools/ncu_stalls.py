"""Stall-reason breakdown of one launch in an .ncu-rep source page (SASS), per kernel ROLE region.
Regions are split at USETMAXREG instructions (the role branches of the warp-specialised conv kernel) in address order.
usage: python tools/ncu_stalls.py rep [ntop]"""
import csv, subprocess, sys
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
col = {n: i for i, n in enumerate(h)}
data = [r for r in rows[hi + 1:] if len(r) == len(h)]
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
def g(r, n):
    try: return int(float(r[col[n]] or 0))
    except ValueError: return 0
tot = sum(g(r, '# Samples') for r in data)
print("total samples", tot, "instructions", len(data))
# regions
bounds = [0] + [i for i, r in enumerate(data) if "SETMAXREG" in r[col['Source']]] + [len(data)]
for a, b in zip(bounds[:-1], bounds[1:]):
    seg = data[a:b]
    s = sum(g(r, '# Samples') for r in seg)
    ex = sum(g(r, 'Instructions Executed') for r in seg)
    if s == 0: continue
    br = sorted(((sum(g(r, n) for r in seg), n) for n in stalls), reverse=True)[:7]
    print(f"region [{a},{b}) first='{seg[0][col['Source']].strip()[:40]}' samples {s} ({100*s/tot:.1f}%) warp-instr {ex}")
    print("    " + "  ".join(f"{n[6:]} {100*v/s:.0f}%" for v, n in br if v))
top = sorted(range(len(data)), key=lambda i: -g(data[i], '# Samples'))[:ntop]
for i in sorted(top):
    r = data[i]
    br = sorted(((g(r, n), n) for n in stalls), reverse=True)[:2]
    print(f"{i:5d} {g(r,'# Samples'):6d} {100*g(r,'# Samples')/max(tot,1):5.1f}% exec={g(r,'Instructions Executed'):8d} {r[col['Source']].strip()[:70]:70s} " + " ".join(f"{n[6:]}={v}" for v, n in br if v))
