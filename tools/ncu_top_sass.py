"""Top stall-sampled SASS instructions of one launch in an .ncu-rep (source page exported with --print-source sass)."""
import csv, subprocess, sys
rep, skip = sys.argv[1], sys.argv[2]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
col = {n: i for i, n in enumerate(h)}
data = [r for r in rows[hi + 1:] if len(r) == len(h) and r[0] != "Address"]
def g(r, n):
    try: return int(float(r[col[n]] or 0))
    except ValueError: return 0
tot = sum(g(r, '# Samples') for r in data)
print([r for r in rows[:hi] if r][:3])
print("total samples", tot, "instructions", len(data))
top = sorted(range(len(data)), key=lambda i: -g(data[i], '# Samples'))[:ntop]
for i in sorted(top):
    r = data[i]
    print(f"{i:5d} {g(r,'# Samples'):6d} {100*g(r,'# Samples')/max(tot,1):5.1f}% exec={g(r,'Instructions Executed'):8d} {r[col['Source']][:96]:96s} confl={r[col['L1 Conflicts Shared N-Way']]} wf={r[col['L1 Wavefronts Shared']]}")
