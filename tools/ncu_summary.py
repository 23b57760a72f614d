"""Condense `ncu -i X.ncu-rep --page raw --csv` into the JSON summaries kept under profiles/ (one object per captured
launch with the metrics the roofline discussion uses).   python tools/ncu_summary.py raw.csv out.json [name-filter]"""
import csv
import json
import sys

KEEP = ("Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_lsu.sum",
        "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__cycles_active.avg",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed")


def main():
    src, dst = sys.argv[1], sys.argv[2]
    flt = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(open(src, newline="")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = dict(zip(names, r))
        if flt and flt not in d["Kernel Name"]:
            continue
        o = {}
        for k in KEEP:
            if k in d:
                u = units[names.index(k)]
                o[k] = f"{d[k]} {u}".strip()
        out.append(o)
    json.dump(out, open(dst, "w"), indent=1)
    print(f"{len(out)} launches -> {dst}")


if __name__ == "__main__":
    main()
