"""Summarise an .ncu-rep (read here, no GPU needed):  python tools/ncu_summary.py rep.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active", "sm__pipe_tensor_op",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit", "launch__waves_per_multiprocessor",
        "sm__inst_executed_pipe_tensor", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg ", "smsp__cycles_active.avg",
        "l1tex__t_bytes", "lts__t_bytes.sum", "sm__cycles_active.avg", "sm__inst_executed_pipe_uniform", "launch__grid_size",
        "smsp__warp_issue_stalled", "sm__pipe_tmem", "tmem"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("=" * 100)
        print(name[:160], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for i, h in enumerate(hdr):
            if any(w in h for w in WANT):
                v = r[i]
                if v not in ("", "0", "n/a"):
                    print(f"  {h:95s} {units[i]:14s} {v}")


if __name__ == "__main__":
    main()
