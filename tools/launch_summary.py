"""Per-kernel totals of ONE sampling step from an ncu launch list (`--metrics gpu__time_duration.sum --csv`): the launches
between two consecutive superpose_update_kernel launches.  Durations under ncu are cold-cache and serialised: shares, not
absolute times, are what carries over to the un-profiled step."""
import collections, csv, re, sys


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    col = {n: i for i, n in enumerate(h)}
    out = []
    for r in rows[hi + 1:]:
        if len(r) != len(h) or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        val, unit = float(r[col["Metric Value"]].replace(",", "")), r[col["Metric Unit"]]
        us = val / 1000 if unit.startswith("n") else (val if unit.startswith("u") else val * 1000)
        out.append((r[col["Kernel Name"]], us))
    return out


def main(path):
    L = load(path)
    idx = [i for i, (n, _) in enumerate(L) if "superpose_update_kernel" in n]
    step = L[idx[0] + 1: idx[1] + 1]
    tot = sum(u for _, u in step)
    agg = collections.OrderedDict()
    for n, u in step:
        k = re.sub(r"^void |sdd::", "", re.sub(r"\(.*", "", n))
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += u
    print(f"{path}: one step = {len(step)} launches, {tot:.0f} us (sum of kernel durations under ncu)")
    for k, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {u:9.1f} us {100 * u / tot:5.1f} %  x{c:3d}  {k[:100]}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
