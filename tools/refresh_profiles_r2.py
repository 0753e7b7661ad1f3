#!/usr/bin/env python
"""Turn the raw outputs of tools/profile_r2.sh (gpurun_out/) into the tracked
round-2 summaries under profiles/ and refresh traffic.json (DRAM bytes per line
of one step, read by bench.py for roofline.traffic)."""
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
l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
l1tex__t_sector_hit_rate.pct lts__t_sector_hit_rate.pct l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
smsp__issue_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum
smsp__thread_inst_executed_per_inst_executed.ratio sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__warps_active.avg.pct_of_peak_sustained_active
launch__block_size launch__grid_size launch__registers_per_thread launch__shared_mem_per_block_dynamic
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio""".split()


def launch_list(src, dst, summary):
    shutil.copy(src, dst)
    with open(dst) as fh:
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
    with open(summary, "w") as fh:
        fh.write("kernel,launches,total_us,dram_bytes_total\n")
        for k, a in agg.items():
            fh.write('"%s",%d,%.1f,%.0f\n' % (k, a[0], a[1] / 1e3, a[2]))
    return agg


def ncu_summary(rep, dst, command):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {"kernel": vals[hdr.index("Kernel Name")], "command": command, "metrics": {}}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP:
            try:
                out["metrics"][h] = {"value": float(v.replace(",", "")), "unit": u}
            except ValueError:
                pass
    # wavefronts by shared-memory opcode, from the source page
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    sh = srows[1]
    ix = {h: i for i, h in enumerate(sh)}
    ops = collections.OrderedDict()
    for r in srows[2:]:
        t = r[ix["Source"]].split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        if not any(k in op for k in ("LDS", "STS", "ATOMS", "UBLKCP", "SYNCS", "LDGSTS", "STG", "LDG")):
            continue
        a = ops.setdefault(op, [0.0, 0.0, 0.0])
        for j, key in enumerate(("Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal")):
            try:
                a[j] += float(r[ix[key]])
            except ValueError:
                pass
    out["memory_instructions"] = {k: {"executed": a[0], "shared_wavefronts": a[1], "ideal": a[2]} for k, a in ops.items()}
    with open(dst, "w") as fh:
        json.dump(out, fh, indent=1)


def main():
    for a, b in (("r2_bench.json", "r2_bench_full_100M.json"), ("r2_sweep_instant.txt", "r2_sweep_instant.txt"),
                 ("r2_sweep_last.txt", "r2_sweep_last.txt"), ("r2_sweep_full.txt", "r2_sweep_full.txt"),
                 ("r2_sweep_single.txt", "r2_sweep_single.txt"), ("r2_stream.txt", "r2_stream.txt")):
        if os.path.exists(os.path.join(SRC, a)):
            shutil.copy(os.path.join(SRC, a), os.path.join(DST, b))
    agg = launch_list(os.path.join(SRC, "r2_launches.csv"), os.path.join(DST, "r2_launches.csv"),
                      os.path.join(DST, "r2_launches_summary.csv"))
    if os.path.exists(os.path.join(SRC, "r2_stream_launches.csv")):
        launch_list(os.path.join(SRC, "r2_stream_launches.csv"), os.path.join(DST, "r2_stream_launches.csv"),
                    os.path.join(DST, "r2_stream_launches_summary.csv"))
    hot = next(v for k, v in agg.items() if "batch_hot_kernel" in k)
    per_line = hot[2] / hot[0] / LINES
    with open(os.path.join(DST, "traffic.json"), "w") as fh:
        json.dump({"dram_bytes_per_line": per_line, "lines_in_capture": LINES,
                   "note": "dram__bytes_read.sum + dram__bytes_write.sum of one step = one batch_hot_kernel launch "
                           "(ncu, profiles/r2_launches.csv), scaled by lines"}, fh)
    # the full captures: one batch_hot_kernel launch, one stream_hot_kernel launch
    ncu_summary(os.path.join(SRC, "r2_hot.ncu-rep"), os.path.join(DST, "r2_hot_ncu_summary.json"),
                "bench.py --lines 20000000 --steps 3 --warmup 3 --no-e2e --no-cpu --no-configs (4th launch of the kernel)")
    if os.path.exists(os.path.join(SRC, "r2_stream_hot.ncu-rep")):
        ncu_summary(os.path.join(SRC, "r2_stream_hot.ncu-rep"), os.path.join(DST, "r2_stream_hot_ncu_summary.json"),
                    "tools/profile_stream.py --once (union64 over 4.3 GB, style last: the main window's walk)")
    # SASS excerpt: the bulk copy, the barrier instructions, the interior loop
    sass = subprocess.run([os.path.join(ROOT, "tools", "sass_of.sh"),
                           os.path.join(ROOT, "one_b200", "_build", "kernels_hot.cu.o"),
                           "batch_hot_kernelILi1ELb1ELi132"], capture_output=True, text=True).stdout.splitlines()
    keep = set()
    for i, ln in enumerate(sass):
        if any(k in ln for k in ("UBLKCP", "SYNCS", "LDGSTS", "LDS.128")):
            keep.update(range(max(0, i - 2), min(len(sass), i + 3)))
    # the interior loop: the densest run of LDS.U8
    u8 = [i for i, ln in enumerate(sass) if "LDS.U8" in ln]
    best = max(u8, key=lambda i: sum(1 for j in u8 if i <= j < i + 80)) if u8 else 0
    keep.update(range(max(0, best - 12), min(len(sass), best + 90)))
    with open(os.path.join(DST, "r2_hot_sass_excerpt.txt"), "w") as fh:
        fh.write("# cuobjdump -sass of batch_hot_kernel<Instant, deferred, 128-column rows> (one_b200/_build/kernels_hot.cu.o):\n"
                 "# the bulk-async copy (UBLKCP), the mbarrier instructions (SYNCS), the text loads (LDS.128), the\n"
                 "# re-alignment (SEL / PRMT) and the interior loop (16 x PRMT / IMAD / LDS.U8 per chunk).\n"
                 "# counts in the whole kernel: %s\n" % ", ".join(
                     "%s %d" % (k, sum(1 for ln in sass if k in ln)) for k in ("UBLKCP", "SYNCS", "LDGSTS", "LDS.128", "LDS.U8", "LDG")))
        last = -2
        for i in sorted(keep):
            if i != last + 1:
                fh.write("...\n")
            fh.write(sass[i].rstrip()[:110] + "\n")
            last = i
    print("profiles refreshed; dram bytes per line %.2f" % per_line)


if __name__ == "__main__":
    main()
