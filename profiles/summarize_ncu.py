"""Extract the judged metrics from an .ncu-rep (read here, no GPU needed):
   python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread ", "launch__grid_size", "launch__block_size", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma", "sm__pipe_tensor", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed.sum ", "smsp__average_warps_issue_stalled", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__cycles_active.avg.pct", "sm__cycles_elapsed.avg "]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, unit = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== kernel:", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for h, u, v in zip(hdr, unit, r):
            if any(h.startswith(w.strip()) if w.endswith(" ") else w in h for w in WANT):
                if "pcsamp" in h:
                    continue
                print(f"{h} [{u}] = {v}")


if __name__ == "__main__":
    main(sys.argv[1])
