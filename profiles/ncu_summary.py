"""Prints the judged subset of an `ncu --page raw --csv` export (one kernel per row)."""
import csv
import re
import sys

PAT = re.compile(
    r"^(gpu__time_duration.sum|dram__bytes_(read|write).sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|"
    r"sm__throughput.avg.pct_of_peak_sustained_elapsed|smsp__issue_active.avg.pct_of_peak_sustained_active|"
    r"smsp__inst_executed.sum|sm__inst_executed_pipe_(xu|alu|lsu|fma|fmaheavy|fmalite|adu|uniform|cbu|tensor)\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|"
    r"smsp__average_warps?_issue_stalled_\w+_per_issue_active.ratio|sm__warps_active.avg.pct_of_peak_sustained_active|"
    r"smsp__warps_eligible.avg.per_cycle_active|sm__cycles_elapsed.avg|lts__t_bytes.sum|lts__throughput.avg.pct_of_peak_sustained_elapsed|"
    r"l1tex__throughput.avg.pct_of_peak_sustained_elapsed|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed|"
    r"launch__(registers_per_thread|grid_size|block_size|occupancy_limit_\w+|shared_mem_per_block_dynamic)|"
    r"smsp__thread_inst_executed_per_inst_executed.ratio)$")


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(vals, units)))
        print("== kernel:", d.get("Kernel Name", ("?",))[0][:100])
        for k in sorted(d):
            if PAT.match(k):
                print(f"  {k:95s} {d[k][0]:>16s} {d[k][1]}")


if __name__ == "__main__":
    main(sys.argv[1])
