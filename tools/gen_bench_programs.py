#!/usr/bin/env python
"""Programs bench.py runs besides the headline union64 (BASELINE.json configs),
compiled once and committed as tests/golden/bench_programs.json:

  keywords7000   configs[2]: 7000 exact keywords, loose start, distinct result
                 ids -> 46 613 states (fmt 4), far beyond shared memory.
                 Compiled by the product's own pipeline (250 s; the reference
                 needs 24 s for 300 keywords and grows faster than linearly).
  flume_*        the `prog.*phrase` rules of the reference's flume configuration
                 (quol/flume/cfg.tsv:12-60), one program per tailed log file,
                 result id = position of the rule in the file's list, added with
                 addAuto (loose at both ends): the smallest result id among the
                 rules that match is "first trigger wins"
                 (quol/flume/src/tailer.cpp:86-96).  Compiled by the product's
                 pipeline; where the reference's compiler finishes within
                 --ref-timeout seconds its program is checked to be equivalent.

    python tools/gen_bench_programs.py [--ref-timeout 300]
"""
import argparse
import base64
import json
import multiprocessing as mp
import os
import struct
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
CFG = "/root/reference/quol/flume/cfg.tsv"


def compile_product(rules):
    from one_b200 import red
    p = red.Parser()
    for pat, rid, flags, lang in rules:
        p.addAs(lang, pat, rid, flags)
    return red.compileToSerialized(p)


def _ref_worker(rules, q):
    from oracle.loader import RefLib
    q.put(RefLib.compile(rules, max_states=0))


def compile_reference(rules, timeout):
    q = mp.Queue()
    p = mp.Process(target=_ref_worker, args=(rules, q))
    p.start()
    try:
        blob = q.get(timeout=timeout)
    except Exception:  # noqa: BLE001 - timed out
        blob = None
    p.terminate()
    p.join()
    return blob


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-timeout", type=float, default=300.0)
    ap.add_argument("--skip-big", action="store_true")
    args = ap.parse_args()
    from one_b200 import synth
    import redaformat as rf
    out = {"programs": {}}

    groups = {}
    for ln in open(CFG):
        f = ln.rstrip("\n").split("\t")
        if f[0] == "scan":
            groups.setdefault(f[1], []).append(f[2])
    for path, pats in groups.items():
        name = "flume_" + os.path.basename(path).replace(".", "_")
        rules = [(p.encode(), i + 1, 0, synth.LANG_AUTO) for i, p in enumerate(pats)]
        t0 = time.time()
        blob = compile_product(rules)
        dt = time.time() - t0
        ref = compile_reference(rules, args.ref_timeout)
        checked = None if ref is None else bool(rf.equivalent(blob, ref))
        # (False: the reference's minimizer merges states of different result here, see compile/dfa.cpp;
        #  the product's default keeps the smallest-id-wins meaning, tests/test_compile.py)
        out["programs"][name] = {
            "file": path, "rules": pats, "compiler": "product", "compile_seconds": round(dt, 1),
            "equivalent_to_reference_program": checked,
            "states": struct.unpack_from("<I", blob, 16)[0],
            "program": base64.b64encode(zlib.compress(blob, 9)).decode()}
        print(name, len(pats), "rules", out["programs"][name]["states"], "states", "checked:", checked, flush=True)

    if not args.skip_big:
        rules = synth.rules_keywords(7000)
        t0 = time.time()
        blob = compile_product(rules)
        out["programs"]["keywords7000"] = {
            "rules": "one_b200.synth.rules_keywords(7000)", "compiler": "product",
            "compile_seconds": round(time.time() - t0, 1),
            "states": struct.unpack_from("<I", blob, 16)[0],
            "program": base64.b64encode(zlib.compress(blob, 9)).decode()}
        print("keywords7000", out["programs"]["keywords7000"]["states"], "states", flush=True)

    with open(os.path.join(ROOT, "tests", "golden", "bench_programs.json"), "w") as f:
        json.dump(out, f)
    print("written", os.path.getsize(os.path.join(ROOT, "tests", "golden", "bench_programs.json")), "bytes")


if __name__ == "__main__":
    main()
