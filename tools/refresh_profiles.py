#!/usr/bin/env python
"""Turn the raw outputs of a GPU profiling call (gpurun_out/) into the tracked
summaries under profiles/:
  bench_final.json      -> r1_bench_full_100M.json
  final_launches.csv    -> r1_launches.csv, r1_launches_summary.csv   (ncu launch list with DRAM bytes)
  sorted_final.ncu-rep  -> r1_final_ncu_summary.json                   (ncu --set full of the batch kernel)
  sweep_final.txt       -> r1_final_sweep.txt
  pextra_final.txt      -> r1_extra_paths.txt
and traffic.json (DRAM bytes per line of one step, both kernels)."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")
LINES = 20_000_000

KEEP = """gpu__time_duration.sum sm__cycles_elapsed.max dram__bytes_read.sum dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum memory_l1_wavefronts_shared_ideal
l1tex__t_sector_hit_rate.pct lts__t_sector_hit_rate.pct l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
smsp__issue_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum
smsp__thread_inst_executed_per_inst_executed.ratio sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active sm__warps_active.avg.pct_of_peak_sustained_active
launch__block_size launch__grid_size launch__registers_per_thread launch__shared_mem_per_block_dynamic
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio""".split()


def main():
    for a, b in (("bench_final.json", "r1_bench_full_100M.json"), ("final_launches.csv", "r1_launches.csv"),
                 ("sweep_final.txt", "r1_final_sweep.txt"), ("pextra_final.txt", "r1_extra_paths.txt")):
        if os.path.exists(os.path.join(SRC, a)):
            shutil.copy(os.path.join(SRC, a), os.path.join(DST, b))
    with open(os.path.join(DST, "r1_launches.csv")) as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        per.setdefault((row["ID"], row["Kernel Name"].split("(")[0]), {})[row["Metric Name"]] = v
    agg = collections.OrderedDict()
    for (_, k), m in per.items():
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0)
        a[2] += m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
    with open(os.path.join(DST, "r1_launches_summary.csv"), "w") as fh:
        fh.write("kernel,launches,total_us,dram_bytes_total\n")
        for k, (c, tt, db) in agg.items():
            fh.write('"%s",%d,%.1f,%.0f\n' % (k, c, tt / 1e3, db))
    ours = {k: v for k, v in agg.items() if "red_b200" in k}
    step_bytes = sum(v[2] / v[0] for v in ours.values())
    for k, v in ours.items():
        print("%-70s %d launches, %.1f us and %.3f GB of DRAM traffic per launch" % (k[:70], v[0], v[1] / 1e3 / v[0], v[2] / v[0] / 1e9))
    rep = os.path.join(SRC, "sorted_final.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    metrics = {h: [v, u] for h, u, v in zip(hdr, units, vals) if h in KEEP}
    bench = json.load(open(os.path.join(DST, "r1_bench_full_100M.json")))
    algo = bench["roofline"]["algorithmic_bytes_per_launch"] / bench["config"]["lines_per_gpu"] * LINES
    doc = {"kernel": "batch_sorted_kernel<instant, no start, smem u16 K=1, cls u8>, tiles from the pre-sorted array (production default)",
           "command": "python bench.py --lines 20000000 --steps 3 --warmup 3 --no-e2e --no-cpu",
           "capture": "ncu --set full --clock-control none --import-source on -k regex:batch_sorted -s 3 -c 1",
           "lines_per_launch": LINES, "algorithmic_bytes_per_launch": algo,
           "dram_bytes_per_launch": step_bytes, "traffic_over_algorithmic": step_bytes / algo,
           "note": "dram_bytes_per_launch is the sum over the two kernels of a step (block_sort_kernel + batch_sorted_kernel) "
                   "from the launch list r1_launches.csv; the metrics below are the batch kernel's",
           "metrics": metrics}
    json.dump(doc, open(os.path.join(DST, "r1_final_ncu_summary.json"), "w"), indent=1)
    json.dump({"dram_bytes_per_line": step_bytes / LINES, "lines_in_capture": LINES,
               "note": "dram__bytes_read.sum + dram__bytes_write.sum of one step = block_sort_kernel + batch_sorted_kernel "
                       "(ncu, profiles/r1_launches.csv), scaled by lines"},
              open(os.path.join(DST, "traffic.json"), "w"))
    print("traffic / algorithmic: %.3f" % (step_bytes / algo))
    for k in ("gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
              "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
              "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
              "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"):
        print(k, metrics.get(k))
    print("bench: %.1f GB/s, %.3f ms/step, roofline %.4f, e2e %.1f, grep %.1f, cpu %.1f on %d cores, launches %d" % (
        bench["value"], bench["ms_per_step"], bench["roofline"]["frac"], bench["e2e"]["value"],
        bench["e2e"]["grep"]["value"], bench["cpu_baseline"]["value"], bench["cpu_baseline"]["cores"], bench["gpu_launches"]))


if __name__ == "__main__":
    sys.exit(main())
