"""Summarise an .ncu-rep (read on the CPU box): per launch duration, DRAM bytes, occupancy,
issue utilisation, tensor-pipe utilisation.  usage: ncu_summary.py rep.ncu-rep out.md"""
import csv, subprocess, sys, io

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__cycles_elapsed.max", "lts__t_bytes.sum", "l1tex__t_bytes.sum"]


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    with open(out, "w") as f:
        f.write(f"# ncu summary of {rep}\n\n(ncu --set full --clock-control none; per-launch values, cold cache, serialised)\n\n")
        f.write("| " + " | ".join(f"{k} [{units[i]}]" if units[i] else k for k, i in idx) + " |\n")
        f.write("|" + "---|" * len(idx) + "\n")
        for r in rows[2:]:
            f.write("| " + " | ".join(r[i] for _, i in idx) + " |\n")
    print("wrote", out, len(rows) - 2, "launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
