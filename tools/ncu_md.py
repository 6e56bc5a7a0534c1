"""Summarise the CSV pages of an `ncu --set full --import-source on` capture (tools/ncu_r02.sh) as markdown for profiles/.

    python tools/ncu_md.py gpurun_out/ncu_<name>_raw.csv [gpurun_out/ncu_<name>_source.csv] > profiles/r02_ncu_<name>.md
"""
import collections
import csv
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__cycles_elapsed.avg.per_second",
        "lts__t_sectors_srcunit_tex_op_read.sum"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "Ghz": 1e9, "Mhz": 1e6, "hz": 1.0}


def rows_of(path):
    return list(csv.reader(l for l in open(path) if not l.startswith("==")))


def raw(path):
    rows = rows_of(path)
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"## `{r[hdr.index('Kernel Name')][:110]}`\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in hdr:
                print(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        try:   # derived: L2 -> SM delivery rate (the GEMMs' wall)
            val = lambda k: float(r[hdr.index(k)]) * UNIT[units[hdr.index(k)]]
            rate = val("l1tex__m_xbar2l1tex_read_bytes.sum") / val("gpu__time_duration.sum")
            print(f"| derived: L2 -> L1 read rate | {rate / 1e12:.2f} TB/s = {rate / val('lts__cycles_elapsed.avg.per_second'):.0f} B per L2 clock | |")
        except (ValueError, KeyError):
            pass
        print()


def source(path):
    rows = rows_of(path)
    for i, r in enumerate(rows):
        if "Source" in r and "Instructions Executed" in r:
            hdr, body = r, rows[i + 1:]
            break
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    ops, opsamp, stalls, hot = collections.Counter(), collections.Counter(), collections.Counter(), []
    total_i = total_s = 0
    for n, r in enumerate(body):
        try:
            e, smp = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        except (ValueError, IndexError):
            continue
        src = r[ix["Source"]].strip()
        toks = src.split()
        op = (toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")).split(".")[0]
        ops[op] += e
        opsamp[op] += smp
        total_i += e
        total_s += smp
        hot.append((smp, e, n, src[:70]))
        for c in stall_cols:
            try:
                stalls[c] += int(r[ix[c]])
            except ValueError:
                pass
    print(f"### source page: {total_i} warp instructions executed, {total_s} stall samples\n")
    print("| opcode | warp instructions | share | stall samples |\n|---|---|---|---|")
    for k, v in ops.most_common(14):
        print(f"| {k} | {v} | {100 * v / total_i:.1f} % | {opsamp[k]} |")
    print("\n| stall reason | samples | share |\n|---|---|---|")
    for k, v in stalls.most_common(8):
        print(f"| {k} | {v} | {100 * v / max(1, total_s):.1f} % |")
    print("\n| hottest instructions (by stall samples) | samples | executed |\n|---|---|---|")
    for smp, e, n, src in sorted(hot, reverse=True)[:10]:
        print(f"| #{n} `{src}` | {smp} ({100 * smp / max(1, total_s):.1f} %) | {e} |")
    print()


if __name__ == "__main__":
    raw(sys.argv[1])
    if len(sys.argv) > 2:
        source(sys.argv[2])
