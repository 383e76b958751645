"""Summarise an .ncu-rep (read here, no GPU) into the handful of counters DESIGN.md cites.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "sm__cycles_active.avg",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"== {r[hdr.index('Kernel Name')]}  (id {r[hdr.index('ID')]})")
        for k in KEYS:
            if k in hdr:
                print(f"   {k:75s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
        st = [(hdr[i], r[i]) for i in range(len(hdr))
              if "smsp__average_warp" in hdr[i] and "issue_stalled" in hdr[i] and hdr[i].endswith("_per_issue_active.ratio")]
        st.sort(key=lambda x: -float(x[1].replace(",", "") or 0))
        for k, v in st[:6]:
            print(f"   stall {k.split('issue_stalled_')[1].split('_per_issue')[0]:40s} {v} warps/issue")


if __name__ == "__main__":
    main(sys.argv[1])
