#!/usr/bin/env python
"""bench.py - throughput of the RED batch matcher on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one pass of the hot path (red_b200_match_batch_device: one kernel
launch) over the whole synthetic corpus of this rank: the flume-style union of
64 rules (one_b200/synth.py) over --lines synthetic syslog lines, match style
instant ("contained"), per-line result id and match end written back.

ours:       value = GB/s of line bytes scanned, inputs resident in HBM.
            e2e   = same metric through red_b200_match_batch_host with pinned
                    HOST buffers (H2D of text+index and D2H of results inside
                    the timed region).
            roofline / cpu_baseline / clocks as the contract describes.
reference:  the unmodified reference (oracle/_ref/libred_ref.so: RED's own
            match() looped over the lines on all host cores), bounded sample.

Multi-GPU (torchrun): every rank matches its own shard of the corpus (weak
scaling, no collective in the data path); the per-rank match counts are
gathered once after the timed region.
"""

import argparse
import base64
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

STYLE = "instant"
VERB = "match"
POOL_LINES = 200_000
DEFAULT_LINES = 100_000_000


def load_program_blob(name="union64"):
    """Serialized DFA of the 64-rule union, as compiled by the reference itself
    (committed fixture made by tools/gen_golden.py)."""
    with open(os.path.join(ROOT, "tests", "golden", "ref_vectors.json")) as f:
        case = next(c for c in json.load(f)["cases"] if c["name"] == name)
    return zlib.decompress(base64.b64decode(case["program"]))


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_pool(seed):
    from one_b200 import synth
    return synth.make_corpus(POOL_LINES, seed=seed)


def tile_counts(n_lines):
    reps = max(1, (n_lines + POOL_LINES - 1) // POOL_LINES)
    return reps, reps * POOL_LINES


def device_corpus(pool, n_lines, device):
    """Tile the pool on the GPU up to n_lines (rounded up to whole pools)."""
    import torch
    text_p, starts_p, lens_p = pool
    reps, total = tile_counts(n_lines)
    tp = torch.from_numpy(text_p.copy()).to(device)
    text = tp.repeat(reps)
    sp = torch.from_numpy(starts_p.astype(np.int64)).to(device)
    off = torch.arange(reps, device=device, dtype=torch.int64) * tp.numel()
    starts = (off[:, None] + sp[None, :]).reshape(-1).contiguous()
    lens = torch.from_numpy(lens_p.astype(np.int32)).to(device).repeat(reps)
    return text, starts, lens, total


def host_corpus(pool, n_lines, pin):
    import torch
    text_p, starts_p, lens_p = pool
    reps, total = tile_counts(n_lines)
    text = torch.empty(reps * text_p.size, dtype=torch.uint8, pin_memory=pin)
    tn = text.numpy()
    for r in range(reps):
        tn[r * text_p.size:(r + 1) * text_p.size] = text_p
    starts = torch.empty(total, dtype=torch.int64, pin_memory=pin)
    sn = starts.numpy().view(np.uint64)
    lens = torch.empty(total, dtype=torch.int32, pin_memory=pin)
    ln = lens.numpy().view(np.uint32)
    for r in range(reps):
        sn[r * POOL_LINES:(r + 1) * POOL_LINES] = starts_p + np.uint64(r * text_p.size)
        ln[r * POOL_LINES:(r + 1) * POOL_LINES] = lens_p
    return text, starts, lens, total


def run_reference_sample(blob, pool, cores, target_s=4.0, max_lines=20_000_000):
    """RED's own matcher on all host cores over a bounded sample."""
    from oracle.loader import RefLib, Oracle, have_ref
    text_p, starts_p, lens_p = pool
    reps = max(1, min(max_lines // POOL_LINES, 4 * cores))
    text = np.tile(text_p, reps)
    starts = np.concatenate([starts_p + np.uint64(r * text_p.size) for r in range(reps)])
    lens = np.tile(lens_p, reps)
    nbytes = int(lens.astype(np.uint64).sum())
    if have_ref():
        eng, kind = RefLib(blob), "reference"
        run = lambda: eng.batch(text, starts, lens, VERB, STYLE, nthreads=cores)  # noqa: E731
    else:
        eng, kind, cores = Oracle(blob), "port", 1
        run = lambda: eng.batch(text, starts, lens, VERB, STYLE)  # noqa: E731
    run()  # warm the caches and the page tables
    passes, elapsed = 0, 0.0
    while elapsed < target_s and passes < 50:
        t0 = time.perf_counter()
        res = run()
        elapsed += time.perf_counter() - t0
        passes += 1
    gbs = nbytes * passes / elapsed / 1e9
    lps = len(starts) * passes / elapsed
    sample = "%d lines (%.2f GB) x %d passes, %s match/%s, %d threads" % (
        len(starts), nbytes / 1e9, passes, VERB, STYLE, cores)
    return {"value": gbs, "unit": "GB/s", "cores": cores, "kind": kind, "sample": sample,
            "lines_per_s": lps, "seconds": elapsed, "matched": int((res[0] > 0).sum())}, nbytes, len(starts)


def bind_to_gpu_numa_node(local):
    """Several ranks on one host: keep this rank's threads - and with them the
    pinned staging memory they first touch - on the NUMA node its GPU hangs
    off, so that the host-to-device copies of the e2e leg do not cross sockets.
    Best effort: silently does nothing where sysfs does not say."""
    try:
        import torch
        bdf = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(
            torch.cuda.get_device_properties(local), "pci_bus_id") else None
        if bdf is None:
            import ctypes
            rt = ctypes.CDLL("libcudart.so")
            buf = ctypes.create_string_buffer(32)
            if rt.cudaDeviceGetPCIBusId(buf, 32, local) != 0:
                return
            bdf = buf.value.decode()
        if isinstance(bdf, int):
            return
        path = "/sys/bus/pci/devices/%s/numa_node" % bdf.lower()
        node = int(open(path).read().strip())
        if node < 0:
            return
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:  # noqa: BLE001 - purely an optimisation
        return


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lines", type=int, default=DEFAULT_LINES, help="lines per GPU")
    ap.add_argument("--e2e-lines", type=int, default=20_000_000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--program", default="union64")
    args = ap.parse_args()

    # stdout carries exactly one JSON line.  Libraries write banners to fd 1
    # behind Python's back (NCCL prints its version there): send fd 1 to stderr
    # for the run and keep the real stdout for the result.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    blob = load_program_blob(args.program)
    workload = ("flume-style union of 64 rules (distinct result ids, fLooseStart) over synthetic "
                "syslog lines, match style instant ('contained'), result id + match end per line")

    if args.impl == "reference":
        if rank != 0:
            return
        pool = host_pool(1234)
        steps = []
        for i in range(args.warmup + args.steps):
            r, nbytes, nlines = run_reference_sample(blob, pool, cores, target_s=2.0)
            if i >= args.warmup:
                steps.append(r)
        val = statistics.mean(s["value"] for s in steps)
        ms = statistics.mean(nbytes / (s["value"] * 1e9) * 1e3 for s in steps)
        out = {"impl": "reference", "metric": "scan_throughput", "value": val, "unit": "GB/s",
               "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
               "data": "synthetic", "lines_per_s": statistics.mean(s["lines_per_s"] for s in steps),
               "config": {"workload": workload, "lines_per_step": nlines, "bytes_per_step": nbytes,
                          "program": args.program},
               "cpu_baseline": {"value": val, "unit": "GB/s", "cores": steps[0]["cores"],
                                "kind": steps[0]["kind"], "sample": steps[0]["sample"]},
               "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(out)
        return

    import torch
    import torch.distributed as dist
    from one_b200 import red

    if red.device_count() < 1:
        raise SystemExit("bench.py needs an sm_100 GPU: red_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        bind_to_gpu_numa_node(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL's own banner/debug lines go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    prog = red.Program(blob, device=local)
    info = prog.info
    pool = host_pool(1234 + rank)  # every rank its own shard of the corpus
    text, starts, lens, n = device_corpus(pool, args.lines, device)
    line_bytes = int(pool[2].astype(np.uint64).sum()) * (n // POOL_LINES)
    res = torch.empty(n, dtype=torch.int32, device=device)
    ends = torch.empty(n, dtype=torch.int32, device=device)
    out_bufs = (res, None, ends)

    def step():
        prog.match_batch_device(text, starts, lens, VERB, STYLE, out=out_bufs)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = prog.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    barrier()
    t_all0 = torch.cuda.Event(enable_timing=True)
    t_all1 = torch.cuda.Event(enable_timing=True)
    t_all0.record()
    for a, b in evs:
        a.record()
        step()
        b.record()
    t_all1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = prog.launch_count - launches0
    total_ms = t_all0.elapsed_time(t_all1)
    kern_ms = statistics.mean(a.elapsed_time(b) for a, b in evs)
    matched = int((res > 0).sum().item())

    t = torch.tensor([total_ms, kern_ms, float(line_bytes), float(n), float(matched)],
                     dtype=torch.float64, device=device)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, kern_ms = float(tmax[0]), float(tmax[1])
        all_bytes, all_lines, all_matched = float(tsum[2]), float(tsum[3]), float(tsum[4])
    else:
        all_bytes, all_lines, all_matched = float(line_bytes), float(n), float(matched)
    ms_per_step = total_ms / args.steps
    value = all_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the host-buffer C-ABI call -----------------------
    e2e = None
    if not args.no_e2e:
        import psutil
        avail = psutil.virtual_memory().available
        want = min(args.e2e_lines, args.lines)
        per_line = (line_bytes / n) + 1 + 12 + 8
        want = int(min(want, 0.25 * avail / world / per_line))
        htext, hstarts, hlens, hn = host_corpus(pool, max(want, POOL_LINES), pin=True)
        hres = torch.empty(hn, dtype=torch.int32, pin_memory=True)
        hends = torch.empty(hn, dtype=torch.int32, pin_memory=True)
        h_line_bytes = int(pool[2].astype(np.uint64).sum()) * (hn // POOL_LINES)
        outs = (hres.numpy(), None, hends.numpy().view(np.uint32))
        tn, sn, ln = htext.numpy(), hstarts.numpy().view(np.uint64), hlens.numpy().view(np.uint32)

        def e2e_step():
            prog.match_batch_host(tn, sn, ln, VERB, STYLE, out=outs)

        for _ in range(2):
            e2e_step()
        barrier()
        k = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(k):
            e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / k
        tt = torch.tensor([dt], dtype=torch.float64, device=device)
        tb = torch.tensor([float(h_line_bytes)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        assert int((hres > 0).sum()) > 0
        e2e = {"value": float(tb[0]) / float(tt[0]) / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": int(htext.numel() + hn * 12),
               "d2h_bytes_per_step": int(hn * 8), "lines_per_step": hn,
               "ms_per_step": float(tt[0]) * 1e3,
               "api": "red_b200_match_batch_host (pinned host text/index in, results+ends out)"}
        # the blob entry point (text only over PCIe, matching lines back):
        # same lines, same outcomes, reported beside e2e
        recs = np.zeros(hn // 2, dtype=red.LINE_MATCH_DTYPE)
        nm, nl = __import__("ctypes").c_uint64(), __import__("ctypes").c_uint64()

        def grep_step():
            red._check(red.lib().red_b200_grep_host(prog._h, tn.ctypes.data, tn.size, 10, red.STYLES[STYLE],
                                                    recs.ctypes.data, recs.size,
                                                    __import__("ctypes").byref(nm), __import__("ctypes").byref(nl)))

        grep_step()
        assert nl.value == hn and nm.value == int((hres > 0).sum()), (nl.value, hn, nm.value)
        t0 = time.perf_counter()
        for _ in range(k):
            grep_step()
        gdt = (time.perf_counter() - t0) / k
        tg = torch.tensor([gdt], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        e2e["grep"] = {"value": float(tb[0]) / float(tg[0]) / 1e9, "unit": "GB/s",
                       "h2d_bytes_per_step": int(htext.numel()),
                       "d2h_bytes_per_step": int(nm.value * 24), "ms_per_step": float(tg[0]) * 1e3,
                       "api": "red_b200_grep_host (log blob in, matching lines out; index, match, "
                              "compaction on the GPU)"}
        del htext, hstarts, hlens, hres, hends

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    # algorithmic bytes per line: the line, its start (8) and length (4),
    # result (4) and match end (4) written back
    algo_bytes = line_bytes + n * (8 + 4 + 4 + 4)
    achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "kernel": "block_sort_kernel + batch_sorted_kernel<instant> (both inside the timed launch duration; the batch kernel is 96% of it)", "algorithmic_bytes_per_launch": algo_bytes,
                "kernel_ms": kern_ms}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        with open(prof) as f:
            tj = json.load(f)
        roofline["traffic"] = tj["dram_bytes_per_line"] * n
        roofline["traffic_note"] = tj.get("note")

    cpu = None
    if not args.no_cpu:
        cpu, _, _ = run_reference_sample(blob, pool, cores)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    out = {"metric": "scan_throughput", "value": value, "unit": "GB/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
           "data": "synthetic", "lines_per_s": all_lines / (ms_per_step * 1e-3),
           "frac_of_hbm_peak": value / (peak * world),
           "config": {"workload": workload, "lines_per_gpu": n, "line_bytes_per_gpu": line_bytes,
                      "avg_line_bytes": line_bytes / n, "matched_lines": all_matched,
                      "l2": "inputs (%.1f GB per GPU) are far larger than the 126 MB L2" % (text.numel() / 1e9),
                      "program": {k: info[k] for k in ("state_count", "class_count", "table_mode",
                                                       "table_bytes", "format")},
                      "corpus": "pool of %d distinct lines tiled on the device" % POOL_LINES,
                      "parallelism": "lines sharded by rank, no data-path collective"},
           "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
           "roofline": roofline, "cpu_baseline": cpu}
    if rank == 0:
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
