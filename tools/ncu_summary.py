"""Text summary of an `ncu --page raw --csv` export (+ optional gzipped `--page source --csv`): the counters the roofline claims
rest on, the warp-stall mix and the instructions with the most stall samples.
usage: python tools/ncu_summary.py <raw.csv> [<source.csv.gz>] > profiles/<name>.txt"""
import collections
import csv
import gzip
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def main(raw, source=None):
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {raw}")
    print(f"Kernel: {vals[col['Kernel Name']]}")
    for w in WANT:
        if w in col:
            print(f"{w:78s} {vals[col[w]]:>18s} {units[col[w]]}")
    stalls = [(h, float(vals[i].replace(',', ''))) for h, i in col.items()
              if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and vals[i]]
    if stalls:
        print("\nwarp stall reasons (average warps stalled per issue-active cycle):")
        for h, v in sorted(stalls, key=lambda kv: -kv[1])[:8]:
            print(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:26s} {v:6.2f}")
    if source:
        rows = list(csv.reader(gzip.open(source, "rt")))
        hdr, data = rows[1], rows[2:]
        iS, iN, iP = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        tot = sum(int(r[iP]) for r in data) or 1
        print("\ninstructions with the most stall samples (share of all samples, executed count):")
        for r in sorted(data, key=lambda r: -int(r[iP]))[:12]:
            print(f"  {100 * int(r[iP]) / tot:5.1f}%  exec {r[iN]:>9s}  {r[iS].strip()[:96]}")
        h = collections.Counter()
        for r in data:
            s = r[iS].strip().split()
            if s:
                h[s[1] if s[0].startswith('@') and len(s) > 1 else s[0]] += int(r[iN])
        print("\nexecuted warp instructions by opcode (top 12): " + ", ".join(f"{k} {v / 1e6:.1f}M" for k, v in h.most_common(12)))


if __name__ == "__main__":
    main(*sys.argv[1:3])
