"""Summarise an .ncu-rep into a small CSV (one row per captured launch) for profiles/.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.csv"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    with open(out, "w", newline="") as fh:
        wr = csv.writer(fh)
        wr.writerow([f"{w} [{units[i]}]" if units[i] else w for w, i in idx])
        for r in rows[2:]:
            wr.writerow([r[i][:120] for _, i in idx])
    print(f"{len(rows) - 2} launches -> {out}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
