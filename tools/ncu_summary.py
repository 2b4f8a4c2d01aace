#!/usr/bin/env python
"""Compact per-kernel summary of an ncu report (run here, no GPU):
    python tools/ncu_summary.py gpurun_out/prof_gemm.ncu-rep > profiles/r01_gemm_ncu.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    # the tcgen05 pipe: sm__pipe_tensor_cycles_active (what SURVEY 8(d) asks for) and sm__mem_tensor_cycles_active agree to
    # 0.1 % on these kernels; the *_subpipe_hmma counters stay below 1 % because UTCHMMA is not an HMMA-subpipe instruction
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("kernel:", r[col["Kernel Name"]][:110])
        for k in KEYS:
            if k in col and r[col[k]] not in ("", "n/a"):
                print("  %-82s %s %s" % (k, r[col[k]], units[col[k]]))
        hm, cyc = col.get("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg"), col.get("sm__cycles_elapsed.avg")
        if hm is not None and cyc is not None and r[hm] not in ("", "n/a"):
            v = float(r[hm].replace(",", "")) / 4.0 / float(r[cyc].replace(",", ""))
            print("  %-82s %.1f %%" % ("tensor pipe active (hmma subpipe cycles / 4 / elapsed cycles)", 100 * v))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
