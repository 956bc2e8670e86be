"""Turn an .ncu-rep (ncu --set full) into the small per-launch CSV kept under profiles/.

    python profiles/summarize_ncu.py gpurun_out/r1b_prof.ncu-rep profiles/r1b_ncu_full_bag_forward_summary.csv
"""
import csv
import subprocess
import sys

KEEP = ("Kernel Name", "gpu__time_duration", "dram__bytes", "dram__throughput", "gpu__dram_throughput", "warps_active",
        "registers_per_thread", "occupancy", "issue_active", "inst_executed.sum", "stalled", "lts__t_sector_hit",
        "l1tex__t_sector_hit", "sm__throughput", "launch__grid_size", "launch__block_size", "smem",
        "sm__pipe_tensor", "sm__inst_executed_pipe_tensor", "lts__t_bytes.sum", "dram__cycles_active")


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    keep = [i for i, h in enumerate(hdr) if any(k in h for k in KEEP)]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{j}" for j in range(len(data))])
        for i in keep:
            w.writerow([hdr[i], units[i]] + [r[i] for r in data])
    for name in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                 "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
                 "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active"):
        if name in hdr:
            i = hdr.index(name)
            print(name, units[i], [r[i] for r in data])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
