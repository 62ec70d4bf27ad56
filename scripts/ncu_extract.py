"""Trim an .ncu-rep (read with `ncu -i ... --page raw --csv`) to the metrics the roofline discussion uses.
usage: python scripts/ncu_extract.py gpurun_out/prof_gemm.ncu-rep profiles/r1_ncu_gemm.csv"""
import csv
import io
import re
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = csv.reader(io.StringIO(raw))
    hdr = next(r)
    units = next(r)
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([name + (f" [{units[i]}]" if units[i] else "") for name, i in idx])
        for row in r:
            w.writerow([re.sub(r"\(.*", "", row[i]) if name == "Kernel Name" else row[i] for name, i in idx])
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
