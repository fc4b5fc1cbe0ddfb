#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep (read with `ncu -i ... --page raw --csv`)."""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occ limit regs"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe % (active)"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe cycles active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % of peak (elapsed)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % of peak (active)"),
    ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "  DMMA sub-pipe % (active)"),
    ("sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "  IMMA sub-pipe % (active)"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe instructions"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue active % (elapsed)"),
    ("smsp__inst_executed.sum", "instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp latency / inst"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== %s" % d.get("Kernel Name", "?")[:110])
        for key, label in WANT:
            for h, u in zip(hdr, units):
                if h == key and d.get(h, "") != "":
                    print("   %-28s %s %s" % (label, d[h], u))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
