"""Condense an `ncu --page raw --csv` export into the metrics the round summaries under profiles/ quote.

    ncu -i rep.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_summary.py raw.csv "<title line>" > profiles/<name>.txt
"""
import csv
import sys

KEEP = (
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled",
    "smsp__average_warp_latency_per_inst_issued",
)


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    print("# " + (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"## {name}")
        for h, u, v in zip(hdr, units, r):
            if any(k in h for k in KEEP) and v not in ("", "n/a"):
                if ".sum.p" in h or ".max" in h or ".min" in h or "per_second" in h and "cycles" not in h:
                    continue
                print(f"{h} [{u}] = {v}")


if __name__ == "__main__":
    main()
