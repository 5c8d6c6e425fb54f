"""From an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log of bench.py: keep the
last N launches (one step), write a compact per-launch CSV, a per-kernel summary (time share, DRAM bytes) and the JSON that
bench.py reads for `roofline.traffic` (DRAM bytes per igemm launch).

usage: python tools/launch_traffic.py <ncu_log.csv> <launches_per_step> <out_prefix> <workload> [round]
"""
import collections
import csv
import json
import re
import sys


def clean(name):
    m = re.match(r"(?:void )?(?:.*?::)?([A-Za-z0-9_]+)(<[^(]*?>)?\(", name.replace("(bool)", "").replace("(int)", ""))
    return (m.group(1) + (m.group(2) or "")) if m else name


def main(path, per_step, prefix, workload, rnd=2):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    iid, ik, im, iv, iu = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    launch = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= iv:
            continue
        launch.setdefault(int(r[iid]), {"name": r[ik]})[r[im]] = (float(r[iv].replace(",", "")), r[iu])
    last = sorted(launch)[-per_step:]
    byte = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    ns = {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    with open(prefix + ".csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["ID", "Kernel Name", "gpu__time_duration.sum [ns]", "dram__bytes_read.sum [B]", "dram__bytes_write.sum [B]"])
        for i in last:
            d = launch[i]
            t = d["gpu__time_duration.sum"][0] * ns[d["gpu__time_duration.sum"][1]]
            rd = d["dram__bytes_read.sum"][0] * byte[d["dram__bytes_read.sum"][1]]
            wr = d["dram__bytes_write.sum"][0] * byte[d["dram__bytes_write.sum"][1]]
            w.writerow([i, clean(d["name"]), int(t), int(rd), int(wr)])
            a = agg[clean(d["name"])]
            a[0] += 1; a[1] += t / 1e6; a[2] += rd; a[3] += wr
    tot = sum(a[1] for a in agg.values())
    lines = [f"# {prefix}.csv: one {workload.upper()} step, {sum(a[0] for a in agg.values())} launches, {tot:.3f} ms captured under ncu (cold-cache, serialised: compare SHARES)",
             "# command: ncu --kernel-name-base demangled -k regex:wc:: --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
             f"--clock-control none --csv python bench.py --workload {workload} --resident-only --steps 1 --warmup 1 --no-cpu-baseline (last {per_step} launches = one step)",
             f"{'ms':>9} {'share':>7} {'launches':>8} {'DRAM rd MB':>11} {'DRAM wr MB':>11} {'GB/s':>7}  kernel"]
    for n, (c, ms, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{ms:9.3f} {100 * ms / tot:6.1f}% {c:8d} {rd / 1e6:11.1f} {wr / 1e6:11.1f} {(rd + wr) / ms / 1e6:7.0f}  {n}")
    open(prefix + "_summary.txt", "w").write("\n".join(lines) + "\n")
    ig = [a for n, a in agg.items() if n.startswith("igemm_kernel")]
    c = sum(a[0] for a in ig); rd = sum(a[2] for a in ig); wr = sum(a[3] for a in ig); ms = sum(a[1] for a in ig)
    json.dump({"workload": workload, "kernel": "igemm_kernel", "launches_per_step": c, "dram_bytes_read_per_step": rd,
               "dram_bytes_write_per_step": wr, "dram_bytes_per_launch": (rd + wr) / c, "share_of_step_under_ncu": ms / tot,
               "source": prefix + ".csv"}, open(f"profiles/r{rnd}_dram_traffic_{workload}.json", "w"), indent=1)
    print("\n".join(lines[:12]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4], int(sys.argv[5]) if len(sys.argv) > 5 else 2)
