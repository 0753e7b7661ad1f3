#!/usr/bin/env python
"""bench.py - throughput of the RED batch matcher on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one pass of the hot path (red_b200_match_batch_device: one kernel
launch) over this rank's lines of ONE synthetic corpus: the flume-style union
of 64 rules (one_b200/synth.py) over --lines x N synthetic syslog lines, match
style instant ("contained"), per-line result id and match end written back.

ours:       value = GB/s of line bytes scanned, inputs resident in HBM.
            e2e   = same metric through red_b200_match_batch_host with pinned
                    HOST buffers (H2D of text+index and D2H of results inside
                    the timed region), beside the box's measured concurrent
                    pinned-H2D bandwidth.
            roofline / cpu_baseline / clocks as the contract describes.
            configs = the other BASELINE.json configurations, one short
                    measurement each (N = 1 only): single regex, 46 k-state
                    union on the global table, 16 GB stream, the reference's
                    flume rule sets.
reference:  the unmodified reference (oracle/_ref/libred_ref.so: RED's own
            match() looped over the lines on all host cores), bounded sample.

Multi-GPU (torchrun): the corpus is cut into one contiguous, byte-balanced range
of lines per rank (one_b200.shard), every rank matches its range with its own
copy of the program - no collective in the data path - and the matched line
numbers are gathered over NCCL after the timed region and checked on rank 0.
Weak scaling is the headline (--lines per GPU); the same cut of a corpus of
--lines lines in all is timed as well and reported as strong_scaling.
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


def device_corpus_range(pool, first, end, device):
    """Lines [first, end) of the global corpus (the pool repeated for ever) on the
    GPU: whole copies of the pool's text, starts relative to the first copy."""
    import torch
    text_p, starts_p, lens_p = pool
    if end <= first:
        z = torch.zeros(0, dtype=torch.int64, device=device)
        return torch.zeros(16, dtype=torch.uint8, device=device), z, z.to(torch.int32), 0
    r0, r1 = first // POOL_LINES, (end - 1) // POOL_LINES
    reps = r1 - r0 + 1
    tp = torch.from_numpy(text_p.copy()).to(device)
    text = tp.repeat(reps)
    sp = torch.from_numpy(starts_p.astype(np.int64)).to(device)
    off = torch.arange(reps, device=device, dtype=torch.int64) * tp.numel()
    starts = (off[:, None] + sp[None, :]).reshape(-1)
    lens = torch.from_numpy(lens_p.astype(np.int32)).to(device).repeat(reps)
    lo, hi = first - r0 * POOL_LINES, end - r0 * POOL_LINES
    return text, starts[lo:hi].contiguous(), lens[lo:hi].contiguous(), end - first


def range_bytes(pool, first, end):
    """line bytes of lines [first, end) of the global corpus"""
    lens = pool[2].astype(np.uint64)
    csum = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    per_pool = int(csum[-1])

    def upto(i):
        return (i // POOL_LINES) * per_pool + int(csum[i % POOL_LINES])
    return upto(end) - upto(first)


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


def pool_expected(blob, pool, verb=VERB, style=STYLE, lines=None):
    """The checker (not the thing measured): the CPU matcher's outcomes on the
    pool's lines - the compiled reference where it is present, else the port."""
    from oracle.loader import RefLib, Oracle, have_ref
    text_p, starts_p, lens_p = pool
    if lines is not None:
        starts_p, lens_p = starts_p[:lines], lens_p[:lines]
    eng = RefLib(blob) if have_ref() else Oracle(blob)
    if have_ref():
        return eng.batch(text_p, starts_p, lens_p, verb, style, nthreads=max(1, os.cpu_count() or 1))
    return eng.batch(text_p, starts_p, lens_p, verb, style)


def bench_program(name):
    with open(os.path.join(ROOT, "tests", "golden", "bench_programs.json")) as f:
        rec = json.load(f)["programs"][name]
    return zlib.decompress(base64.b64decode(rec["program"])), rec


def other_configs(red, pool, device, peak, args):
    """BASELINE.json's other configurations, one short measurement each on one
    GPU: a few timed launches after warm-up (CUDA events), GB/s of line bytes,
    the same roofline arithmetic as the headline, and a spot check of the
    outcomes against the CPU matcher on the pool's lines."""
    import torch
    out = []
    sample = 20000  # lines of the pool the checker walks

    def lines_case(name, workload, blob, n_lines, style, verb="match", steps=5, note=None):
        prog = red.Program(blob, device=device.index)
        text, starts, lens, n = device_corpus_range(pool, 0, tile_counts(n_lines)[1], device)
        res = torch.empty(n, dtype=torch.int32, device=device)
        ends = torch.empty(n, dtype=torch.int32, device=device)
        bufs = (res, None, ends)
        for _ in range(3):
            prog.match_batch_device(text, starts, lens, verb, style, out=bufs)
        torch.cuda.synchronize()
        l0, h0 = prog.launch_count, prog.hot_launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            prog.match_batch_device(text, starts, lens, verb, style, out=bufs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        nbytes = range_bytes(pool, 0, n)
        eres, _, eme = pool_expected(blob, pool, verb, style, lines=sample)
        ok = bool(np.array_equal(res[:sample].cpu().numpy(), eres) and
                  np.array_equal(ends[:sample].cpu().numpy().view(np.uint32), eme))
        algo = nbytes + n * 20
        info = prog.info
        rec = {"name": name, "workload": workload, "lines": n, "style": style, "ms": ms,
               "value": nbytes / ms / 1e6, "unit": "GB/s", "lines_per_s": n / (ms * 1e-3),
               "roofline": {"bound": "hbm", "achieved": algo / ms / 1e6, "peak": peak, "frac": algo / ms / 1e6 / peak},
               "program": {k: info[k] for k in ("state_count", "class_count", "table_mode", "table_bytes")},
               "launches_per_step": (prog.launch_count - l0) / steps,
               "staged_kernel_launches_per_step": (prog.hot_launch_count - h0) / steps,
               "matched_lines": int((res > 0).sum().item()),
               "check": "first %d lines identical to the CPU matcher: %s" % (sample, ok)}
        if note:
            rec["note"] = note
        assert ok, name
        prog.close()
        del text, starts, lens, res, ends
        torch.cuda.empty_cache()
        return rec

    # configs[0]: the reference's own CPU-runnable case
    out.append(lines_case("single_regex", "configs[0]: 'sshd.*Failed password' (addAuto) over 1 M synthetic lines, "
                          "match style instant ('contained')", load_program_blob("single_auto"), 1_000_000, "instant"))
    have = os.path.exists(os.path.join(ROOT, "tests", "golden", "bench_programs.json"))
    if have:
        # configs[2]: a union far beyond shared memory, on the L2-resident table
        blob, rec = bench_program("keywords7000")
        for sty in ("first", "last"):
            out.append(lines_case("large_union_" + sty, "configs[2]: 7000 keywords (addExact, loose start, distinct "
                                  "result ids) = %d states, over 20 M synthetic lines, match style %s" % (
                                      rec["states"], sty), blob, 20_000_000, sty, steps=3))
        # the reference's flume configuration: one program per tailed log file,
        # smallest matching rule number = first trigger wins (tailer.cpp:86-96)
        flume = []
        for name in ("flume_auth_log", "flume_mail_log", "flume_external_log", "flume_internal_log",
                     "flume_daemon_log"):
            try:
                blob, rec = bench_program(name)
            except KeyError:
                continue
            r = lines_case(name, "flume rules of %s (%d two-part 'prog.*phrase' rules, addAuto, result = rule "
                           "number): %d states" % (rec["file"], len(rec["rules"]), rec["states"]),
                           blob, 20_000_000, "instant", steps=3)
            flume.append(r)
            out.append(r)
        if flume:
            tot_ms = sum(r["ms"] for r in flume)
            nbytes = range_bytes(pool, 0, flume[0]["lines"])
            out.append({"name": "flume_all_tailers", "workload": "the same 20 M lines through all %d flume "
                        "programs one after the other (every tailer sees every line)" % len(flume),
                        "ms": tot_ms, "value": nbytes / tot_ms / 1e6, "unit": "GB/s of lines, all programs applied"})

    # configs[3]: one 16 GB stream
    if not args.no_stream:
        from oracle.loader import Oracle
        blob = load_program_blob("union64")
        prog = red.Program(blob, device=device.index)
        tp = torch.from_numpy(pool[0].copy()).to(device)
        reps = (16 * (1 << 30)) // tp.numel() + 1
        stream_text = tp.repeat(reps)
        nbytes = int(stream_text.numel())
        for sty in ("last", "full"):
            o = prog.match_stream_device(stream_text, "match", sty)
            torch.cuda.synchronize()
            l0 = prog.launch_count
            t0 = time.perf_counter()
            k = 3
            for _ in range(k):
                o = prog.match_stream_device(stream_text, "match", sty)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) / k * 1e3
            # the checker: the stream is the pool's text over and over, and this DFA
            # forgets what lies more than a line back: the last copy alone tells
            # where the last match ends
            one = Oracle(blob).run(pool[0].tobytes(), "match", sty)
            ok = (o.result == one[0]) and ((o.end == (reps - 1) * tp.numel() + one[2]) if one[0] > 0 else o.end == 0)
            out.append({"name": "stream_16GB_" + sty, "workload": "configs[3]: one %.1f GB byte stream (the corpus "
                        "text, newlines included), union64, match style %s" % (nbytes / 1e9, sty),
                        "bytes": nbytes, "ms": ms, "value": nbytes / ms / 1e6, "unit": "GB/s",
                        "roofline": {"bound": "hbm", "achieved": nbytes / ms / 1e6, "peak": peak,
                                     "frac": nbytes / ms / 1e6 / peak},
                        "launches_per_call": (prog.launch_count - l0) / k,
                        "outcome": [int(o.result), int(o.start), int(o.end)],
                        "check": "result and end as the CPU matcher finds them in the last copy of the pool: %s" % ok})
            assert ok, ("stream", sty, o, one)
        prog.close()
        del stream_text
        torch.cuda.empty_cache()
    return out


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
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configurations")
    ap.add_argument("--no-stream", action="store_true", help="skip the 16 GB stream among them")
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
    from one_b200 import red, shard

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
    # ONE corpus: the pool repeated; cut into a byte-balanced range of lines per rank
    pool = host_pool(1234)
    reps_per_gpu, _ = tile_counts(args.lines)

    def my_range(total_reps):
        ranges = shard.byte_balanced_ranges_periodic(pool[2], total_reps, world)
        return ranges[rank]

    def timed(first, end, steps, warmup):
        """K timed passes over lines [first, end) of the global corpus, resident in
        HBM.  -> (total ms, mean per-step ms, line bytes, lines, matched, launches,
        hot launches, the device tensors)"""
        text, starts, lens, n = device_corpus_range(pool, first, end, device)
        res = torch.empty(max(n, 1), dtype=torch.int32, device=device)
        ends = torch.empty(max(n, 1), dtype=torch.int32, device=device)
        out_bufs = (res[:n], None, ends[:n])

        def step():
            prog.match_batch_device(text, starts, lens, VERB, STYLE, out=out_bufs)

        for _ in range(warmup):
            step()
        barrier()
        l0, h0 = prog.launch_count, prog.hot_launch_count
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(steps)]
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for a, b in evs:
            a.record()
            step()
            b.record()
        t1.record()
        barrier()
        total_ms = t0.elapsed_time(t1)
        kern_ms = statistics.mean(a.elapsed_time(b) for a, b in evs)
        return (total_ms, kern_ms, range_bytes(pool, first, end), n, int((res[:n] > 0).sum().item()),
                prog.launch_count - l0, prog.hot_launch_count - h0, (text, starts, lens, res[:n], ends[:n]))

    def reduce_times(total_ms, kern_ms, nbytes, nlines, matched):
        t = torch.tensor([total_ms, kern_ms, float(nbytes), float(nlines), float(matched)],
                         dtype=torch.float64, device=device)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone()
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            return float(tmax[0]), float(tmax[1]), float(tsum[2]), float(tsum[3]), float(tsum[4])
        return total_ms, kern_ms, float(nbytes), float(nlines), float(matched)

    # ---- the headline: weak scaling, --lines per GPU ---------------------------
    first, end = my_range(reps_per_gpu * world)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms, kern_ms, line_bytes, n, matched, launches, hot_launches, tensors = timed(
        first, end, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    text, starts, lens, res, ends = tensors
    total_ms, kern_ms, all_bytes, all_lines, all_matched = reduce_times(total_ms, kern_ms, line_bytes, n, matched)
    ms_per_step = total_ms / args.steps
    value = all_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- after the timed region: gather the matched line numbers ----------------
    gather = None
    if n:
        local_idx = red.compact_matches_device(res)
        torch.cuda.synchronize()
    else:
        local_idx = torch.zeros(0, dtype=torch.int64, device=device)
    if world > 1:
        all_idx = shard.gather_matched_lines(local_idx, first)
    else:
        all_idx = local_idx + first
    if rank == 0:
        # the corpus is periodic: the matched lines of its first copy, and their
        # number in every copy, are what a CPU check of the pool says they are
        head = all_idx[all_idx < POOL_LINES].cpu().numpy()
        gather = {"matched_lines_gathered": int(all_idx.numel()), "head": head,
                  "sorted": bool((all_idx[1:] > all_idx[:-1]).all().item()) if all_idx.numel() > 1 else True}
    del all_idx, local_idx

    # ---- strong scaling: a corpus of --lines lines in all, cut the same way ------
    strong = None
    if world > 1:
        sfirst, send = my_range(reps_per_gpu)
        st_ms, sk_ms, sbytes, sn, smatched, _, _, st = timed(sfirst, send, max(3, min(args.steps, 10)), 3)
        del st
        st_ms, sk_ms, sb, sl, sm = reduce_times(st_ms, sk_ms, sbytes, sn, smatched)
        sms = st_ms / max(3, min(args.steps, 10))
        strong = {"scaling": "strong", "lines_total": int(sl), "value": sb / (sms * 1e-3) / 1e9, "unit": "GB/s",
                  "ms_per_step": sms, "matched_lines": sm}

    # ---- end to end through the host-buffer C-ABI call -----------------------
    e2e = None
    if not args.no_e2e:
        import psutil
        avail = psutil.virtual_memory().available
        want = min(args.e2e_lines, args.lines)
        per_line = (line_bytes / max(n, 1)) + 1 + 12 + 8
        want = int(min(want, 0.25 * avail / world / per_line))
        htext, hstarts, hlens, hn = host_corpus(pool, max(want, POOL_LINES), pin=True)
        hres = torch.empty(hn, dtype=torch.int32, pin_memory=True)
        hends = torch.empty(hn, dtype=torch.int32, pin_memory=True)
        h_line_bytes = int(pool[2].astype(np.uint64).sum()) * (hn // POOL_LINES)
        outs = (hres.numpy(), None, hends.numpy().view(np.uint32))
        tn, sn_, ln = htext.numpy(), hstarts.numpy().view(np.uint64), hlens.numpy().view(np.uint32)

        def e2e_step():
            prog.match_batch_host(tn, sn_, ln, VERB, STYLE, out=outs)

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
        # what the host can feed: every rank copies pinned memory to its GPU at the same time
        cbytes = min(int(htext.numel()), 1 << 30)
        dbuf = torch.empty(cbytes, dtype=torch.uint8, device=device)
        dbuf.copy_(htext[:cbytes], non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            dbuf.copy_(htext[:cbytes], non_blocking=True)
        torch.cuda.synchronize()
        cdt = (time.perf_counter() - t0) / 3
        tc = torch.tensor([cdt], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        ceiling = cbytes * world / float(tc[0]) / 1e9
        moved = (e2e["h2d_bytes_per_step"] + e2e["d2h_bytes_per_step"]) * world / float(tt[0]) / 1e9
        e2e["h2d_ceiling"] = {"value": ceiling, "unit": "GB/s",
                              "what": "%d ranks copying 1 GiB of pinned host memory to their GPUs at the same "
                                      "time (cudaMemcpyAsync), aggregate" % world}
        e2e["pcie_gbs"] = moved
        e2e["frac_of_h2d_ceiling"] = moved / ceiling
        del dbuf
        # the blob entry point (text only over PCIe, matching lines back):
        # same lines, same outcomes, reported beside e2e
        import ctypes
        recs = np.zeros(hn // 2, dtype=red.LINE_MATCH_DTYPE)
        nm, nl = ctypes.c_uint64(), ctypes.c_uint64()

        def grep_step():
            red._check(red.lib().red_b200_grep_host(prog._h, tn.ctypes.data, tn.size, 10, red.STYLES[STYLE],
                                                    recs.ctypes.data, recs.size,
                                                    ctypes.byref(nm), ctypes.byref(nl)))

        grep_step()
        assert nl.value == hn and nm.value == int((hres > 0).sum()), (nl.value, hn, nm.value)
        barrier()
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
    staged = hot_launches == launches
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "kernel": ("batch_hot_kernel<instant> (kernels_hot.cu: one launch per step)" if staged else
                           "block_sort_kernel + batch_sorted_kernel<instant> (both inside the timed launch "
                           "duration)"),
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": kern_ms}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        with open(prof) as f:
            tj = json.load(f)
        roofline["traffic"] = tj["dram_bytes_per_line"] * n
        roofline["traffic_note"] = tj.get("note")

    cpu = None
    check = None
    if not args.no_cpu:
        cpu, _, _ = run_reference_sample(blob, pool, cores)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        # the checker: the pool's matched lines as the CPU matcher sees them
        exp = pool_expected(blob, pool)
        hits = np.nonzero(exp[0] > 0)[0]
        total_reps = reps_per_gpu * world
        check = {"first_copy_identical": bool(np.array_equal(gather["head"], hits)),
                 "matched_lines_expected": int(hits.size) * total_reps,
                 "matched_lines_gathered": gather["matched_lines_gathered"],
                 "in_line_order": gather["sorted"],
                 "how": "matched line numbers gathered from all ranks (shard.gather_matched_lines over NCCL) "
                        "against the CPU matcher on the pool"}
        assert check["first_copy_identical"] and check["in_line_order"], check
        assert check["matched_lines_expected"] == check["matched_lines_gathered"], check

    extra = None
    if (world == 1) and not args.no_configs:
        del text, starts, lens, res, ends, tensors
        torch.cuda.empty_cache()
        extra = other_configs(red, pool, device, peak, args)

    out = {"metric": "scan_throughput", "value": value, "unit": "GB/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
           "data": "synthetic", "lines_per_s": all_lines / (ms_per_step * 1e-3),
           "frac_of_hbm_peak": value / (peak * world),
           "config": {"workload": workload, "lines_per_gpu": n, "line_bytes_per_gpu": line_bytes,
                      "lines_total": int(all_lines),
                      "avg_line_bytes": line_bytes / max(n, 1), "matched_lines": all_matched,
                      "l2": "inputs (%.1f GB per GPU) are far larger than the 126 MB L2" % (line_bytes / 1e9),
                      "program": {k: info[k] for k in ("state_count", "class_count", "table_mode",
                                                       "table_bytes", "format")},
                      "corpus": "one corpus: a pool of %d distinct lines repeated %d times" % (
                          POOL_LINES, reps_per_gpu * world),
                      "parallelism": "byte-balanced contiguous range of lines per rank (one_b200.shard), no "
                                     "data-path collective; matched line numbers gathered after the timed region"},
           "e2e": e2e, "gpu_launches": int(launches), "hot_kernel_launches": int(hot_launches), "clocks": clocks,
           "roofline": roofline, "cpu_baseline": cpu, "gathered_check": check}
    if strong:
        out["strong_scaling"] = strong
    if extra:
        out["configs"] = extra
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
