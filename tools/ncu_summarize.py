"""Summarise ncu output for profiles/ (run HERE, no GPU needed).

    python tools/ncu_summarize.py launches gpurun_out/launches.csv            > profiles/rNN_launches.md
    python tools/ncu_summarize.py full     gpurun_out/prof_x.ncu-rep [...]    > profiles/rNN_ncu_x.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]


def short(name):
    name = re.sub(r"vb::<unnamed>::|void |\(CUtensorMap.*", "", name)
    return name[:90]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        k = short(r[ix["Kernel Name"]])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {us:.1f} | {us / n:.2f} | {100 * us / total:.1f} % |")
    print(f"\ntotal kernel time {total / 1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches "
          "(ncu serialises launches and runs them cold-cache: compare SHARES, not absolutes)")


def full(paths):
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        print(f"## {p}\n")
        for r in rows[2:]:
            print(f"### `{short(r[hdr.index('Kernel Name')])}`\n")
            print("| metric | value | unit |\n|---|---|---|")
            for k in KEYS:
                if k in hdr:
                    print(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
            rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            print(f"\ntraffic = dram read + write = {r[rd]} {units[rd]} + {r[wr]} {units[wr]}\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])
