#!/usr/bin/env python3
"""Turn an .ncu-rep (ncu --set full) into the compact text summary committed under profiles/.
Usage: ncu_summary.py REPORT.ncu-rep [OUT.txt]   (runs `ncu -i ... --page raw --csv`, no GPU needed)"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
]
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep = sys.argv[1]
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full --clock-control none summary of %s" % rep.split("/")[-1], file=out)
    for r in rows[2:]:
        print("\nkernel: %s" % r[idx["Kernel Name"]], file=out)
        traffic = 0.0
        for k in KEYS:
            if k in idx and r[idx[k]] != "":
                print("  %-70s %s %s" % (k, r[idx[k]], units[idx[k]]), file=out)
                if k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    traffic += float(r[idx[k]].replace(",", "")) * UNIT_SCALE.get(units[idx[k]], 1)
        print("  %-70s %d byte" % ("traffic = dram__bytes_read.sum + dram__bytes_write.sum", int(traffic)), file=out)
        print("  json: " + json.dumps({"kernel": r[idx["Kernel Name"]].split("(")[0], "traffic_bytes": int(traffic),
                                       "duration": r[idx["gpu__time_duration.sum"]] + " " + units[idx["gpu__time_duration.sum"]]}),
              file=out)


if __name__ == "__main__":
    main()
