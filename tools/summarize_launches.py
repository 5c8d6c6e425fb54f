"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of captured time)."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= v:
            continue
        name = r[k].split("(")[0].replace("void ", "").replace("wc::<unnamed>::", "")
        val = float(r[v].replace(",", ""))
        ms = val / 1e6 if r[u].startswith("n") else (val / 1e3 if r[u].startswith("u") else val)
        agg[name][0] += 1
        agg[name][1] += ms
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.3f} ms captured (cold-cache, serialised: compare SHARES)")
    print(f"{'ms':>10} {'share':>7} {'launches':>8}  kernel")
    for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ms:10.3f} {100 * ms / tot:6.1f}% {c:8d}  {n}")


if __name__ == "__main__":
    main(sys.argv[1])
